// pyramid.cu -- ComputePyramid (ORBextractor.cpp:1071-1096) on the device.
//
// Level 0 is the input copied into a bordered buffer; level l is cv::resize(INTER_LINEAR) of the
// ROUNDED level l-1 (a sequential chain of 7 launches per batch), reproduced in OpenCV's 11-bit
// fixed point.  Each launch covers every frame of the batch.  The output
// domain is the bordered image: a border pixel evaluates at its reflect-101 source coordinate, so
// copyMakeBorder costs no extra pass.  Threads own 16 (copy) or 4 (resize) consecutive bytes of a
// padded row, so every store is a full aligned vector (ncu, round 1: the one-byte-per-thread
// version spent 148 lane-instructions per pixel and was ALU bound).
#include "orbx_internal.cuh"

namespace orbx {

__device__ __forceinline__ int reflect101(int i, int n)
{
    // |i| <= n + 19 and n >= 20 for every level the extractor accepts: one reflection suffices,
    // the loop only guards tiny levels.
    while (i < 0 || i >= n) {
        if (n == 1) return 0;
        i = i < 0 ? -i : 2 * (n - 1) - i;
    }
    return i;
}

// reflect-101 with at most one reflection per side in the common case (border <= 19 < level size)
__device__ __forceinline__ int reflect_fast(int i, int n)
{
    i = i < 0 ? -i : i;
    i = i >= n ? 2 * (n - 1) - i : i;
    return (unsigned)i < (unsigned)n ? i : reflect101(i, n);
}

// ---------------------------------------------------------------------------------------------
// level 0: input frames -> bordered level-0 buffers.  One thread = one aligned 16-byte chunk of a
// padded destination row (the interior starts at byte kPadX = 32 of a 64-byte aligned pitch).
// ---------------------------------------------------------------------------------------------
constexpr int kCopyRows = 8;                 // rows per thread: amortises the index math (the copy was ALU bound at 85 %)

__global__ void __launch_bounds__(128)
k_level0(const __grid_constant__ Geo g, uint8_t *__restrict__ pyr, const uint8_t *__restrict__ imgs,
         size_t in_pitch, size_t in_fstride, int src_aligned)
{
    const LevelGeom &L = g.lv[0];
    const int B = g.border_on ? kBorder : kMinBlurBorder;
    const int Y0 = (int)(blockIdx.y * 4 + threadIdx.y) * kCopyRows - B;   // first bordered row of this thread
    const int f = blockIdx.z + g.frame0;
    const int chunk = blockIdx.x * 32 + threadIdx.x;
    const int X0 = chunk * 16 - kPadX;                       // first pixel of this chunk
    if (chunk * 16 >= L.pitch || X0 + 15 < -B || X0 >= L.w + B || Y0 >= L.h + B) return;
    const uint8_t *src = imgs + (size_t)f * in_fstride;
    uint8_t *dst = pyr + L.base + (size_t)f * L.frame_stride + (size_t)(Y0 + kPadY) * L.pitch + (size_t)chunk * 16;
    const int nrows = min(kCopyRows, L.h + B - Y0);
    if (g.in_channels != 1) {
        // cvtColor(..2GRAY) fused into the level-0 pass (src/Tracking.cpp:57-70); OpenCV 8U fixed point, 15 bits
        const int C = g.in_channels, ro = g.in_rgb ? 0 : 2, bo = g.in_rgb ? 2 : 0;
        if (src_aligned && X0 >= 0 && X0 + 15 < L.w && (C == 3 || C == 4)) {
            // interior chunks of 16-byte aligned rows: the chunk's 48 / 64 source bytes as three / four 16-byte loads, a pixel's
            // three products as two DP2A (16-bit weights x the pixel's bytes).  The byte-wise form below (three byte loads and
            // two reflections per pixel) took 0.17 ms (RGB) / 0.45 ms (RGBA) per 256 VGA frames against 0.029 ms for gray input.
            const uint32_t w01 = g.in_rgb ? (9798u | (19235u << 16)) : (3735u | (19235u << 16));   // weights of bytes 0, 1
            const uint32_t w2 = g.in_rgb ? 3735u : 9798u;                                          // weight of byte 2 (byte 3: 0)
            const bool lb = B == kMinBlurBorder && X0 == 0, rb = B == kMinBlurBorder && X0 + 16 == L.w;
            for (int j = 0; j < nrows; ++j) {
                const uint4 *row = reinterpret_cast<const uint4 *>(src + (size_t)reflect_fast(Y0 + j, L.h) * in_pitch + (size_t)X0 * C);
                uint32_t px[16];
                if (C == 4) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) { const uint4 t = __ldg(row + q); px[4 * q] = t.x; px[4 * q + 1] = t.y; px[4 * q + 2] = t.z; px[4 * q + 3] = t.w; }
                } else {
                    uint32_t wd[13];
#pragma unroll
                    for (int q = 0; q < 3; ++q) { const uint4 t = __ldg(row + q); wd[4 * q] = t.x; wd[4 * q + 1] = t.y; wd[4 * q + 2] = t.z; wd[4 * q + 3] = t.w; }
                    wd[12] = 0u;
#pragma unroll
                    for (int k = 0; k < 16; ++k) px[k] = __funnelshift_r(wd[(3 * k) >> 2], wd[((3 * k) >> 2) + 1], ((3 * k) & 3) * 8);
                }
                uint32_t w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint32_t v = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t gray = __dp2a_hi(w2, px[4 * q + k], __dp2a_lo(w01, px[4 * q + k], 16384u)) >> 15;
                        v |= gray << (8 * k);
                    }
                    w[q] = v;
                }
                uint8_t *d = dst + (size_t)j * L.pitch;
                *reinterpret_cast<uint4 *>(d) = make_uint4(w[0], w[1], w[2], w[3]);
                if (lb) *reinterpret_cast<uint32_t *>(d - 4) = __byte_perm(w[0], w[1], 0x1234);
                if (rb) *reinterpret_cast<uint32_t *>(d + 16) = __byte_perm(w[2], w[3], 0x3456);
            }
            return;
        }
        // chunks that hold nothing but the 4-px border are served by their neighbours above
        if (src_aligned && (C == 3 || C == 4) && B == kMinBlurBorder && L.w >= 16 && (X0 == -16 || (X0 == L.w && (L.w & 15) == 0))) return;
        for (int j = 0; j < nrows; ++j) {
            const uint8_t *row = src + (size_t)reflect_fast(Y0 + j, L.h) * in_pitch;
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t v = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint8_t *p = row + (size_t)reflect_fast(min(max(X0 + 4 * q + k, -B), L.w + B - 1), L.w) * C;
                    const int gray = (__ldg(p + ro) * 9798 + __ldg(p + 1) * 19235 + __ldg(p + bo) * 3735 + 16384) >> 15;
                    v |= (uint32_t)gray << (8 * k);
                }
                w[q] = v;
            }
            *reinterpret_cast<uint4 *>(dst + (size_t)j * L.pitch) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return;
    }
    // rows that lie inside the image need no reflection (ncu: the reflect loop was 20 % of this kernel's instructions)
    const bool rows_inside = Y0 >= 0 && Y0 + kCopyRows <= L.h;
    if (X0 >= 0 && X0 + 15 < L.w && src_aligned) {
        uint4 v[kCopyRows];
        if (rows_inside) {
            const uint8_t *p = src + (size_t)Y0 * in_pitch + X0;
#pragma unroll
            for (int j = 0; j < kCopyRows; ++j, p += in_pitch) v[j] = __ldg(reinterpret_cast<const uint4 *>(p));
        } else {
#pragma unroll
            for (int j = 0; j < kCopyRows; ++j)
                if (j < nrows) v[j] = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)reflect_fast(Y0 + j, L.h) * in_pitch + X0));
        }
        // The first and the last chunk of a row also write the four reflected pixels beside it (one PRMT of the chunk's own
        // bytes: pixel -k is pixel k, pixel w-1+k is pixel w-1-k).  ncu, round 2: the byte-gather path below ran for ONE lane in
        // every warp of the launch -- the chunk that holds only border pixels -- and was more than half of this kernel's
        // instructions (25 per row against 2 for an interior chunk).
        const bool lb = B == kMinBlurBorder && X0 == 0, rb = B == kMinBlurBorder && X0 + 16 == L.w;
#pragma unroll
        for (int j = 0; j < kCopyRows; ++j)
            if (j < nrows) {
                uint8_t *d = dst + (size_t)j * L.pitch;
                *reinterpret_cast<uint4 *>(d) = v[j];
                if (lb) *reinterpret_cast<uint32_t *>(d - 4) = __byte_perm(v[j].x, v[j].y, 0x1234);
                if (rb) *reinterpret_cast<uint32_t *>(d + 16) = __byte_perm(v[j].z, v[j].w, 0x3456);
            }
        return;
    }
    if (X0 >= 0 && X0 + 19 < L.w) {
        // rows that are not 16-byte aligned (a 1241-px KITTI frame): five aligned words and four funnel shifts per row instead of
        // sixteen byte loads (the 20-byte window stays inside the row, so nothing is read past the caller's buffer)
        const bool lb = B == kMinBlurBorder && X0 == 0;
        for (int j = 0; j < nrows; ++j) {
            const uint8_t *a = src + (size_t)(rows_inside ? Y0 + j : reflect_fast(Y0 + j, L.h)) * in_pitch + X0;
            const uint32_t sh = (uint32_t)((uintptr_t)a & 3) * 8u;
            const uint32_t *a4 = reinterpret_cast<const uint32_t *>(a - ((uintptr_t)a & 3));
            const uint32_t w0 = __ldg(a4), w1 = __ldg(a4 + 1), w2 = __ldg(a4 + 2), w3 = __ldg(a4 + 3), w4 = __ldg(a4 + 4);
            const uint4 v = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
            uint8_t *d = dst + (size_t)j * L.pitch;
            *reinterpret_cast<uint4 *>(d) = v;
            if (lb) *reinterpret_cast<uint32_t *>(d - 4) = __byte_perm(v.x, v.y, 0x1234);
        }
        return;
    }
    // chunks that hold nothing but the 4-px border are served by their neighbours above
    if (B == kMinBlurBorder && L.w >= 20 && (X0 == -16 || (src_aligned && X0 == L.w && (L.w & 15) == 0))) return;
    // edge chunks (and unaligned sources): byte gathers at the reflected columns -- only for the 4-byte words that hold
    // at least one pixel of the bordered level (with the default 4-px border that is one word of an edge chunk, not four;
    // the other bytes of the chunk are padding nobody reads)
    int xs[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) xs[k] = reflect_fast(min(max(X0 + k, -B), L.w + B - 1), L.w);
    bool need[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) need[q] = X0 + 4 * q + 3 >= -B && X0 + 4 * q < L.w + B;
    for (int j = 0; j < nrows; ++j) {
        const uint8_t *row = src + (size_t)reflect_fast(Y0 + j, L.h) * in_pitch;
        uint32_t w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            w[q] = !need[q] ? 0u
                            : (uint32_t)__ldg(row + xs[4 * q]) | ((uint32_t)__ldg(row + xs[4 * q + 1]) << 8) |
                              ((uint32_t)__ldg(row + xs[4 * q + 2]) << 16) | ((uint32_t)__ldg(row + xs[4 * q + 3]) << 24);
        *reinterpret_cast<uint4 *>(dst + (size_t)j * L.pitch) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// level l >= 1 from level l-1.  One thread = 4 consecutive pixels x kGatherRows rows of the padded
// destination.  The horizontal parameters (source offsets, 11-bit coefficients) are unpacked once
// and reused for every row; the row index depends only on blockIdx, so source-row pointers are
// warp-uniform; and, like OpenCV's row cache, the horizontally interpolated lower source row is
// reused when the next destination row starts on it (scale 1.2: ~5 of 6 rows).
// ---------------------------------------------------------------------------------------------
constexpr int kGatherRows = 8;

__global__ void __launch_bounds__(128, 10)
k_resize_gather(const __grid_constant__ Geo g, const uint8_t *__restrict__ pyr_src, uint8_t *__restrict__ pyr_dst,
         const int2 *__restrict__ tables, int level)
{
    const LevelGeom &D = g.lv[level];
    const LevelGeom &S = g.lv[level - 1];
    const int B = g.border_on ? kBorder : kMinBlurBorder;
    const int f = blockIdx.z + g.frame0;
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int X0 = chunk * 4 - kPadX;
    if (chunk * 4 >= D.pitch || X0 + 3 < -B || X0 >= D.w + B) return;
    const int2 *tabx = tables + D.tabx;
    int o0[4], c0[4], c1[4];
    if (X0 >= 0 && X0 + 3 < D.w) {
        // interior chunk: the four table entries are one aligned 32-byte run
        const int4 t01 = __ldg(reinterpret_cast<const int4 *>(tabx + X0));
        const int4 t23 = __ldg(reinterpret_cast<const int4 *>(tabx + X0) + 1);
        o0[0] = t01.x; c0[0] = t01.y; o0[1] = t01.z; c0[1] = t01.w; o0[2] = t23.x; c0[2] = t23.y; o0[3] = t23.z; c0[3] = t23.w;
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int2 t = __ldg(tabx + reflect101(min(max(X0 + k, -B), D.w + B - 1), D.w));
            o0[k] = t.x; c0[k] = t.y;
        }
    }
    // source and destination levels live in the same allocation but never overlap: separate restrict
    // pointers let the compiler hoist the next row's loads above this row's store
    const uint8_t *__restrict__ src = pyr_src + S.base + (size_t)f * S.frame_stride + (size_t)kPadY * S.pitch + kPadX;
    uint8_t *__restrict__ dst = pyr_dst + D.base + (size_t)f * D.frame_stride + (size_t)chunk * 4;
    // per-pixel column pointers.  The second tap reads the byte after the first: where OpenCV clamps the
    // offset at the right edge the table has fraction 0 (c1 == 0), so the byte behind the row (padding
    // inside the pitch) is multiplied by zero instead of being clamped away.
    const uint8_t *col[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        col[k] = src + o0[k];
        c1[k] = c0[k] >> 16; c0[k] = (short)(c0[k] & 0xffff);
    }
    const int2 *taby = tables + D.taby;
    const int Ybase = (int)blockIdx.y * kGatherRows - B;
    const int spitch = S.pitch;
    // the y-table entry of the NEXT row is fetched one iteration ahead: otherwise every row pays two dependent
    // global-load latencies (table entry -> source row pointer -> pixels)
    int2 ty_next = __ldg(taby + reflect_fast(min(Ybase, D.h + B - 1), D.h));
    int prev_sy1 = -0x7fffffff;
    int hp[4] = { 0, 0, 0, 0 };                           // horizontally interpolated lower source row, already >> 4
    uint8_t *drow = dst + (size_t)(Ybase + kPadY) * D.pitch;
#pragma unroll 2
    for (int r = 0; r < kGatherRows; ++r, drow += D.pitch) {
        const int Y = Ybase + r;
        if (Y >= D.h + B) break;
        const int2 ty = ty_next;
        ty_next = __ldg(taby + reflect_fast(min(Y + 1, D.h + B - 1), D.h));
        const int sy0 = min(max(ty.x, 0), S.h - 1), sy1 = min(max(ty.x + 1, 0), S.h - 1);
        const int cy0 = (short)(ty.y & 0xffff), cy1 = ty.y >> 16;
        int h0[4], h1[4];
        if (sy0 == prev_sy1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) h0[k] = hp[k];
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { const uint8_t *p = col[k] + (size_t)(unsigned)(sy0 * spitch); h0[k] = (p[0] * c0[k] + p[1] * c1[k]) >> 4; }
        }
        if (sy1 == sy0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) h1[k] = h0[k];
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) { const uint8_t *p = col[k] + (size_t)(unsigned)(sy1 * spitch); h1[k] = (p[0] * c0[k] + p[1] * c1[k]) >> 4; }
        }
        // ((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2 never leaves [0, 255]: H <= 255 * 2049 and b0 + b1 <= 2049
        // give at most 1020 + 2 before the last shift, so OpenCV's saturate_cast is the identity here
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t v = (uint32_t)(((cy0 * h0[k]) >> 16) + ((cy1 * h1[k]) >> 16) + 2) >> 2;
            out |= v << (8 * k);
            hp[k] = h1[k];
        }
        prev_sy1 = sy1;
        *reinterpret_cast<uint32_t *>(drow) = out;
    }
}

// ---------------------------------------------------------------------------------------------
// Staged variant (the one that runs for every scale factor <= 3).  ncu, round 1: the gather kernel above
// spent ~47 lane-instructions per pixel, most of them 64-bit address arithmetic and byte packing around
// eight LDG.U8 per four pixels.  Here a block first stages the source rows it needs in shared memory with
// 16-byte loads; a thread then gets the two taps of TWO neighbouring pixels with two aligned LDS.32 and one
// PRMT (the four bytes lie inside an 8-byte window because neighbouring source offsets are <= 3 apart --
// checked on the host), and IDP.2A does both multiplies of a pixel: ~14 instructions per interpolated row of
// four pixels instead of ~50.  Border pixels need no special case: the padded tables hold their reflected
// source coordinates and PRMT does not care about the order of the bytes it picks.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}

__global__ void __launch_bounds__(128, 8)
k_resize(const __grid_constant__ Geo g, const uint8_t *__restrict__ pyr_src, uint8_t *__restrict__ pyr_dst,
         const int2 *__restrict__ tables, int level)
{
    extern __shared__ __align__(16) unsigned char rs_tile[];
    __shared__ int4 s_ty[kResizeRows];                     // per destination row: tile byte offsets of the two source rows, cy0, cy1
    const LevelGeom &D = g.lv[level];
    const LevelGeom &S = g.lv[level - 1];
    const int var = g.border_on ? 1 : 0;
    const int B = var ? kBorder : kMinBlurBorder;
    const int f = blockIdx.z + g.frame0;
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const int X0 = chunk * 4 - kPadX;
    const bool active = chunk * 4 < D.pitch && X0 + 3 >= -B && X0 < D.w + B;
    const int kRows = D.rs_rows;
    const int Ybase = (int)blockIdx.y * kRows - B;
    const int TP = D.rs_tile_w;
    // source ranges of this block, precomputed on the host: columns [xr.x, xr.y] (+1 for the second tap), rows [yr.x, yr.y]
    const int2 xr = __ldg(tables + D.rs_xr + var * D.rs_nbx + blockIdx.x);
    const int2 yr = __ldg(tables + D.rs_yr + var * D.rs_nby0 + blockIdx.y);
    if (xr.y < 0) return;                                  // no pixel of this block lies inside the bordered level (uniform)
    const int tx0 = xr.x & ~15, smin = yr.x;
    if (threadIdx.x < kRows) {
        const int2 t = __ldg(tables + D.tabyp + kBorder + min(Ybase + (int)threadIdx.x, D.h + B - 1));
        s_ty[threadIdx.x] = make_int4(((t.x & 0xffff) - smin) * TP, ((t.x >> 16) - smin) * TP, (short)(t.y & 0xffff), t.y >> 16);
    }
    {
        const uint8_t *__restrict__ src = pyr_src + S.base + (size_t)f * S.frame_stride + (size_t)(kPadY + smin) * S.pitch + kPadX + tx0;
        const int nvec = ((xr.y + 1 - tx0) >> 4) + 1, total = nvec * (yr.y - smin + 1);
        // row = i / nvec without an integer division: inv >= 2^32 / nvec with an excess below 2^11 (float rounding + margin),
        // exact as long as i * nvec < 2^21 (a tile has < 2^16 vectors: the host caps it at 46 KB)
        const uint32_t inv = __float2uint_rz(__fdividef(4294967296.f, (float)nvec)) + 1024u;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int row = nvec == 1 ? i : (int)__umulhi((uint32_t)i, inv), v = i - row * nvec;
            cp_async16(rs_tile + row * TP + v * 16, reinterpret_cast<const uint4 *>(src + (size_t)row * S.pitch) + v);
        }
    }
    // horizontal parameters of the thread's four pixels (padded table: no reflect / clamp here)
    int pbase[2]; uint32_t psel[2], cf[4];
    if (active) {
        const int4 *tx = reinterpret_cast<const int4 *>(tables + D.tabxp) + chunk * 2;
        const int4 t01 = __ldg(tx), t23 = __ldg(tx + 1);
        cf[0] = (uint32_t)t01.y; cf[1] = (uint32_t)t01.w; cf[2] = (uint32_t)t23.y; cf[3] = (uint32_t)t23.w;
        const int o0[4] = { t01.x - tx0, t01.z - tx0, t23.x - tx0, t23.z - tx0 };
        // per pixel pair: 8-byte window base inside a tile row and the PRMT selector {p0_a, p1_a, p0_b, p1_b}
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int a = o0[2 * j], b = o0[2 * j + 1];
            pbase[j] = min(a, b) & ~3;
            const uint32_t da = (uint32_t)(a - pbase[j]), db = (uint32_t)(b - pbase[j]);
            psel[j] = da | ((da + 1) << 4) | (db << 8) | ((db + 1) << 12);
        }
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();
    if (!active) return;

    uint8_t *drow = pyr_dst + D.base + (size_t)f * D.frame_stride + (size_t)chunk * 4 + (size_t)(Ybase + kPadY) * D.pitch;
    const int nrows = min(kRows, D.h + B - Ybase);
    int prev_off1 = -1;
    int hp[4] = { 0, 0, 0, 0 };                           // horizontally interpolated lower source row, already >> 4
    // shared-window addresses of the two 8-byte windows, formed once: through the generic pointer the compiler rebuilt the window
    // base (S2UR + UMOV + UIADD3 + ULEA) in front of every interpolated row
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(rs_tile);
    const uint32_t tb[2] = { tile_s + (uint32_t)pbase[0], tile_s + (uint32_t)pbase[1] };
    auto hrow = [&](int off, int (&h)[4]) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            uint32_t w0, w1;
            asm volatile("ld.shared.u32 %0, [%2];\n\tld.shared.u32 %1, [%2+4];" : "=r"(w0), "=r"(w1) : "r"(tb[j] + (uint32_t)off));
            const uint32_t P = __byte_perm(w0, w1, psel[j]);
            h[2 * j] = (int)(__dp2a_lo(cf[2 * j], P, 0u) >> 4);
            h[2 * j + 1] = (int)(__dp2a_hi(cf[2 * j + 1], P, 0u) >> 4);
        }
    };
#pragma unroll 2
    for (int r = 0; r < nrows; ++r, drow += D.pitch) {
        const int4 ty = s_ty[r];                           // {offset of source row 0, of source row 1, cy0, cy1}
        int h0[4], h1[4];
        if (ty.x == prev_off1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) h0[k] = hp[k];
        } else {
            hrow(ty.x, h0);
        }
        if (ty.y == ty.x) {
#pragma unroll
            for (int k = 0; k < 4; ++k) h1[k] = h0[k];
        } else {
            hrow(ty.y, h1);
        }
        // ((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2 never leaves [0, 255] (see k_resize_gather).  Two pixels per
        // register from here on: one PRMT takes the upper halves of two products (the two >> 16), the sums stay below 1024 per
        // 16-bit half, (sum + 2) << 6 puts (sum + 2) >> 2 into the second byte of each half, and a last PRMT collects the four bytes:
        // 8 IMAD + 5 PRMT + 2 IADD + 2 IMAD per four pixels instead of 27 scalar instructions (ncu: 26 % of the kernel).
        uint32_t s2[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const uint32_t a = __byte_perm((uint32_t)(ty.z * h0[2 * j]), (uint32_t)(ty.z * h0[2 * j + 1]), 0x7632u);
            const uint32_t b = __byte_perm((uint32_t)(ty.w * h1[2 * j]), (uint32_t)(ty.w * h1[2 * j + 1]), 0x7632u);
            s2[j] = (a + b) * 64u + 0x00800080u;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) hp[k] = h1[k];
        prev_off1 = ty.y;
        *reinterpret_cast<uint32_t *>(drow) = __byte_perm(s2[0], s2[1], 0x7531u);
    }
}

// copyMakeBorder(REFLECT_101) of one level of one frame, on demand (orbx_download_level with border = 19)
__global__ void k_fill_border(const __grid_constant__ Geo g, uint8_t *__restrict__ pyr, int level, int frame)
{
    const LevelGeom &L = g.lv[level];
    const int X = blockIdx.x * blockDim.x + threadIdx.x - kBorder, Y = blockIdx.y * blockDim.y + threadIdx.y - kBorder;
    if (X >= L.w + kBorder || Y >= L.h + kBorder) return;
    if (X >= 0 && X < L.w && Y >= 0 && Y < L.h) return;                  // interior stays as it is
    uint8_t *img = pyr + L.base + (size_t)frame * L.frame_stride + (size_t)kPadY * L.pitch + kPadX;
    img[(ptrdiff_t)Y * L.pitch + X] = img[(size_t)reflect101(Y, L.h) * L.pitch + reflect101(X, L.w)];
}

void launch_fill_border(const Geo &g, const DevBuffers &b, int level, int frame, cudaStream_t s)
{
    dim3 blk(32, 8), grd((g.lv[level].w + 2 * kBorder + 31) / 32, (g.lv[level].h + 2 * kBorder + 7) / 8);
    k_fill_border<<<grd, blk, 0, s>>>(g, b.pyr, level, frame);
}

void launch_level0(const Geo &g, const DevBuffers &b, const uint8_t *d_imgs, size_t pitch, size_t fstride, int nframes, cudaStream_t s)
{
    const int B = g.border_on ? kBorder : kMinBlurBorder;
    const int chunks = g.lv[0].pitch / 16;
    const int aligned = ((((uintptr_t)d_imgs) | pitch | fstride) & 15) == 0;
    dim3 grd((chunks + 31) / 32, (g.lv[0].h + 2 * B + 4 * kCopyRows - 1) / (4 * kCopyRows), nframes);
    k_level0<<<grd, dim3(32, 4), 0, s>>>(g, b.pyr, d_imgs, pitch, fstride, aligned);
}

void launch_resize(const Geo &g, const DevBuffers &b, int level, int nframes, cudaStream_t s)
{
    const int B = g.border_on ? kBorder : kMinBlurBorder;
    const LevelGeom &L = g.lv[level];
    const int chunks = L.pitch / 4;
    const int bw = L.rs_bw;
    const int rows = L.rs_staged ? L.rs_rows : kGatherRows;
    dim3 grd((chunks + bw - 1) / bw, (L.h + 2 * B + rows - 1) / rows, nframes);
    if (L.rs_staged)
        k_resize<<<grd, bw, (size_t)L.rs_tile_w * L.rs_tile_h, s>>>(g, b.pyr, b.pyr, b.tables, level);
    else
        k_resize_gather<<<grd, bw, 0, s>>>(g, b.pyr, b.pyr, b.tables, level);
}

} // namespace orbx

// describe.cu -- GaussianBlur (ORBextractor.cpp:1045-1046), computeOrientation / IC_Angle
// (:27-54, :422-429), computeOrbDescriptor (:57-97) and the operator() epilogue (:1036-1064).
#include "orbx_internal.cuh"

#include <cuda_fp16.h>
#include <cstdlib>
#include <mutex>

namespace orbx {

__device__ __forceinline__ int reflect101d(int i, int n)
{
    while (i < 0 || i >= n) {
        if (n == 1) return 0;
        i = i < 0 ? -i : 2 * (n - 1) - i;
    }
    return i;
}

// ---------------------------------------------------------------------------------------------
// 7x7 sigma-2 Gaussian in OpenCV's Q8.8 fixed point: taps 18,34,48,56,48,34,18,
// dst = (sum_ij k_i k_j p + 32768) >> 16.  OpenCV runs the horizontal pass first; with exact
// integer intermediates the order is irrelevant, so this kernel runs the VERTICAL pass first:
//   * a lane owns one 4-pixel word column and slides down the rows, keeping the last 7 source
//     words split into even/odd byte lanes (2 x 16 bit per register); the vertical 7-tap sum is
//     SIMD-within-a-register: every 16-bit lane stays <= 65280, so no carries cross lanes
//   * the horizontal pass reads the neighbouring word columns with warp shuffles and folds pairs
//     of 16-bit values with DP2A (2 MACs per instruction) into 32-bit sums
// No shared memory, no barriers; the reflect-101 border comes for free from the bordered pyramid
// buffer (>= 4 px are always written).  ncu, round 1: the shared-memory tile version issued 44
// lane-instructions per pixel; this one ~13.
// A warp covers 30 output word columns (120 px; lanes 0 and 31 are halo) x kBlurRows rows.
// ---------------------------------------------------------------------------------------------
constexpr int kBlurRows = 35;                 // a multiple of the 7-row unrolled window: no rows computed and thrown away
constexpr int kBlurWarps = 4;
#ifndef ORBX_BLUR_PREFETCH
#define ORBX_BLUR_PREFETCH 4
#endif
constexpr int kBlurPrefetch = ORBX_BLUR_PREFETCH;   // source rows in flight per lane ahead of the one being consumed
struct BlurRows { int first[ORBX_MAX_LEVELS + 1]; };   // first blockIdx.y of every level
// DP2A weight words (low byte x low 16-bit lane, next byte x high lane).  Passed as a kernel parameter so that they
// are constant-bank operands; as literals the compiler re-materialised all of them in registers in every row iteration.
struct BlurWeights { uint32_t k0k2, z_k1, k3k1, k2k0, z_k0, k2k2, k0_z, k1k3, k1_z, half; };

__device__ __forceinline__ uint32_t dp2(uint32_t a, uint32_t wts, uint32_t c) { return __dp2a_lo(a, wts, c); }

__global__ void __launch_bounds__(32 * kBlurWarps, 10)
k_blur(const __grid_constant__ Geo g, const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur, const __grid_constant__ BlurRows rows,
       const __grid_constant__ BlurWeights W)
{
    // one launch covers every level: blockIdx.y walks the concatenated row groups of all levels.  Fully unrolled over
    // constant-bank operands (entries past nlevels hold the total, so they never match): as a rolled loop over a by-value
    // struct this lookup was 19 % of the kernel's instructions (ncu source page: 41 instructions per iteration).
    int level = 0;
#pragma unroll
    for (int l = 1; l < ORBX_MAX_LEVELS; ++l) if ((int)blockIdx.y >= rows.first[l]) level = l;
    const LevelGeom &L = g.lv[level];
    const int lane = threadIdx.x, f = blockIdx.z + g.frame0;
    const int y0 = (((int)blockIdx.y - rows.first[level]) * kBlurWarps + threadIdx.y) * kBlurRows;
    if (y0 >= L.h || (int)blockIdx.x * 120 >= L.w) return;
    const int y1 = min(y0 + kBlurRows, L.h);
    const int wc_raw = (int)blockIdx.x * 30 + lane - 1;                 // word column (4 px) in the level, -1 = left halo
    const int wc_max = ((L.pitch - kPadX) >> 2) - 1;
    const int wc = min(wc_raw, wc_max);
    const uint8_t *src = pyr + L.base + (size_t)f * L.frame_stride + (size_t)kPadY * L.pitch + kPadX + 4 * wc;
    uint8_t *dst = blur + L.blur_base + (size_t)f * L.blur_frame_stride + 4 * wc;
    const bool writer = lane >= 1 && lane <= 30 && 4 * wc_raw < L.w;

    constexpr uint32_t K0 = 18, K1 = 34, K2 = 48, K3 = 56;             // k4 = k2, k5 = k1, k6 = k0
    uint32_t lo[7], hi[7];                                             // sliding window of split source rows
    const uint8_t *rp = src + (ptrdiff_t)(y0 - 3) * L.pitch;           // walks down the source rows
#pragma unroll
    for (int i = 0; i < 6; ++i, rp += L.pitch) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(rp));
        lo[i] = w & 0x00ff00ffu; hi[i] = (w >> 8) & 0x00ff00ffu;
    }
    // two-deep software prefetch of the next source rows (ncu: 53 % of the stall samples sat on this load).
    // Rows up to h+3 are read; the buffer has 19 rows below the level, so the reads stay inside it.
    uint32_t wn[7];
    static_assert(kBlurPrefetch <= 7, "the prefetch ring has seven slots");
#pragma unroll
    for (int i = 0; i < kBlurPrefetch; ++i) { wn[i] = __ldg(reinterpret_cast<const uint32_t *>(rp)); rp += L.pitch; }
    uint8_t *dp = dst + (size_t)y0 * L.blur_pitch;
    for (int yb = y0; yb < y1; yb += 7) {
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int y = yb + j;
            // no branch around the body: rows past y1 are computed from valid (allocated) memory and simply not
            // stored, which keeps the shuffles convergent (the compiler bracketed them with WARPSYNC otherwise)
            {
                // window slot (j + 6) % 7 receives source row y + 3; rows y-3 .. y+3 are slots j .. j+6 (mod 7)
                // the rows in flight sit in a ring of seven (= the unroll): every index is static, no register moves
                const uint32_t w = wn[j];
                wn[(j + kBlurPrefetch) % 7] = __ldg(reinterpret_cast<const uint32_t *>(rp)); rp += L.pitch;
                lo[(j + 6) % 7] = w & 0x00ff00ffu; hi[(j + 6) % 7] = (w >> 8) & 0x00ff00ffu;
                // vertical 7-tap, two 16-bit lanes per register: V_lo = (V[4c], V[4c+2]), V_hi = (V[4c+1], V[4c+3])
                const uint32_t vlo = K0 * (lo[j % 7] + lo[(j + 6) % 7]) + K1 * (lo[(j + 1) % 7] + lo[(j + 5) % 7]) +
                                     K2 * (lo[(j + 2) % 7] + lo[(j + 4) % 7]) + K3 * lo[(j + 3) % 7];
                const uint32_t vhi = K0 * (hi[j % 7] + hi[(j + 6) % 7]) + K1 * (hi[(j + 1) % 7] + hi[(j + 5) % 7]) +
                                     K2 * (hi[(j + 2) % 7] + hi[(j + 4) % 7]) + K3 * hi[(j + 3) % 7];
                const uint32_t plo = __shfl_up_sync(0xffffffffu, vlo, 1), phi = __shfl_up_sync(0xffffffffu, vhi, 1);     // column c-1
                const uint32_t nlo = __shfl_down_sync(0xffffffffu, vlo, 1), nhi = __shfl_down_sync(0xffffffffu, vhi, 1); // column c+1
                // horizontal 7-tap on 16-bit values with 32-bit sums (+ rounding), SURVEY.md A2
                uint32_t o0 = dp2(phi, W.k0k2, W.half); o0 = dp2(plo, W.z_k1, o0); o0 = dp2(vlo, W.k3k1, o0); o0 = dp2(vhi, W.k2k0, o0);
                uint32_t o1 = dp2(plo, W.z_k0, W.half); o1 = dp2(phi, W.z_k1, o1); o1 = dp2(vlo, W.k2k2, o1); o1 = dp2(vhi, W.k3k1, o1); o1 = dp2(nlo, W.k0_z, o1);
                uint32_t o2 = dp2(phi, W.z_k0, W.half); o2 = dp2(vlo, W.k1k3, o2); o2 = dp2(vhi, W.k2k2, o2); o2 = dp2(nlo, W.k1_z, o2); o2 = dp2(nhi, W.k0_z, o2);
                uint32_t o3 = dp2(vlo, W.k0k2, W.half); o3 = dp2(vhi, W.k1k3, o3); o3 = dp2(nlo, W.k2k0, o3); o3 = dp2(nhi, W.k1_z, o3);
                const uint32_t out = __byte_perm(__byte_perm(o0, o1, 0x0062), __byte_perm(o2, o3, 0x0062), 0x5410);
                if (writer && y < y1) *reinterpret_cast<uint32_t *>(dp) = out;
                dp += L.blur_pitch;
            }
        }
    }
}

void launch_blur(const Geo &g, const DevBuffers &b, int nframes, cudaStream_t s)
{
    BlurRows rows;
    int total = 0, strips = 1;
    for (int l = 0; l < g.nlevels; ++l) {
        rows.first[l] = total;
        total += (g.lv[l].h + kBlurRows * kBlurWarps - 1) / (kBlurRows * kBlurWarps);
        const int words = (g.lv[l].w + 3) / 4;
        strips = strips > (words + 29) / 30 ? strips : (words + 29) / 30;
    }
    for (int l = g.nlevels; l <= ORBX_MAX_LEVELS; ++l) rows.first[l] = total;
    dim3 grd(strips, total, nframes);
    auto W2 = [](uint32_t a, uint32_t b) { return a | (b << 8); };
    const uint32_t K0 = 18, K1 = 34, K2 = 48, K3 = 56;
    const BlurWeights W = { W2(K0, K2), W2(0, K1), W2(K3, K1), W2(K2, K0), W2(0, K0), W2(K2, K2), W2(K0, 0), W2(K1, K3), W2(K1, 0), 32768u };
    k_blur<<<grd, dim3(32, kBlurWarps), 0, s>>>(g, b.pyr, b.blur, rows, W);
}

// ---------------------------------------------------------------------------------------------
// orientation + descriptor + epilogue: one warp per output keypoint slot.
// ---------------------------------------------------------------------------------------------
// cv::fastAtan2; separate roundings (no FMA) to match the scalar C++ build.
__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale);
    const float p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

constexpr int kDescWarps = 8;
constexpr int kDescSlots = 8;             // keypoints handled by one warp (amortises the pattern load and the sincos)
constexpr int kPatchR = 18;               // largest |rotated pattern offset| (18 px)
constexpr int kPatchRows = 2 * kPatchR + 1;               // 37
constexpr int kPatchWords = (2 * kPatchR + 1 + 3 + 3) / 4; // 37 px + up to 3 px of alignment slack = 11 words
constexpr int kStagePasses = (kPatchRows + 1) / 2;        // two patch rows (22 lanes) per cp.async pass
constexpr int kPatchAlloc = 2 * kStagePasses * kPatchWords; // words per patch buffer (one spare row for the last pass)
// Patch staging variants (ORBX_DESC_STAGE): 0 = 4-byte cp.async.ca, rows of 11 words from a 4-byte aligned column;
// 1 = 16-byte cp.async.cg (L1 bypass), 2 = one cp.async.bulk (TMA engine, mbarrier completion) per row -- both with rows of
// 64 bytes from a 16-byte aligned column (x - 18 rounded down to 16: the 37 columns end at byte 51 at most).
// 3 = 8-byte cp.async.ca, rows of 48 bytes from an 8-byte aligned column, five rows (30 lanes) per pass: 8 LDGSTS.64 per patch
// instead of 19 LDGSTS.32, and a 12-word pitch on which eight consecutive rows of 16 bytes fall into 32 different banks.
__host__ __device__ constexpr int patch_pitch(int stage) { return stage == 0 || stage == 4 ? kPatchWords * 4 : stage == 3 ? 48 : 64; }           // bytes per staged patch row
__host__ __device__ constexpr int patch_buf_words(int stage) { return stage == 0 || stage == 4 ? kPatchAlloc : stage == 3 ? kPatchRows * 12 : kPatchRows * 16; }  // words per patch buffer
__host__ __device__ constexpr int patch_align_mask(int stage) { return stage == 0 || stage == 4 ? ~3 : stage == 3 ? ~7 : ~15; }
// IC_Angle as dot products: the 31 x 31 window is read as 9 aligned words per row, three rows per warp pass
// (lane = row-in-pass * 9 + word), and a host-built table holds for every (alignment, pass, lane) the four
// column weights u and the four row weights v as signed bytes, zero outside the circular patch (umax) and
// outside the window -- m10 += dp4a(pixels, u-weights), m01 += dp4a(pixels, v-weights).
constexpr int kMomentPasses = 11;         // 33 rows >= 31
constexpr int kMomentTabWords = 4 * kMomentPasses * 32 * 2;

__device__ __forceinline__ int dp4a_u8s8(uint32_t pix, uint32_t wts, int acc)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(pix), "r"(wts), "r"(acc));
    return d;
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async16_cg(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8_ca(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// blurred 37 x 37 patch (word aligned, 11 words per row) -> shared memory: lanes 0..21 copy two rows per pass
template <int kStage>
__device__ __forceinline__ void stage_patch(uint32_t *patch, const uint8_t *base, int bp, int lane, uint64_t *bar)
{
    // base: row y - 18 of the blurred level at the variant's aligned first column (x - 18 rounded down); bp: bytes per blurred row
    if (kStage == 1) {
        // 16 bytes per lane, four lanes per 64-byte row, eight rows per pass: 5 LDGSTS.128 per patch, L1 bypassed
        const uint8_t *src = base + (size_t)(lane >> 2) * bp + 16 * (lane & 3);
        uint8_t *dst = reinterpret_cast<uint8_t *>(patch) + lane * 16;
#pragma unroll
        for (int it = 0; it < (kPatchRows + 7) / 8; ++it, src += 8 * bp, dst += 8 * 64)
            if (it * 8 + (lane >> 2) < kPatchRows) cp_async16_cg(dst, src);
        cp_async_commit();
        return;
    }
    if (kStage == 3) {
        const int r = lane / 6, wx = lane - r * 6;
        if (lane < 30) {
            const uint8_t *src = base + (size_t)r * bp + 8 * wx;
            uint8_t *dst = reinterpret_cast<uint8_t *>(patch) + r * 48 + wx * 8;
#pragma unroll
            for (int it = 0; it < 8; ++it, src += 5 * bp)
                if (it < 7 || r < 2) cp_async8_ca(dst + it * 5 * 48, src);      // rows 35, 36 in the last pass
        }
        cp_async_commit();
        return;
    }
    if (kStage == 2) {
        // the TMA engine copies the rows: one 64-byte bulk copy per row, issued by one lane, completion on the buffer's mbarrier
        if (lane == 0) {
            mbar_expect_tx(bar, kPatchRows * 64);
            const uint8_t *src = base;
            uint8_t *dst = reinterpret_cast<uint8_t *>(patch);
#pragma unroll 1
            for (int r = 0; r < kPatchRows; ++r, src += bp, dst += 64) bulk_g2s(dst, src, 64, bar);
        }
        return;
    }
    if (lane < 2 * kPatchWords) {
        const int r = lane >= kPatchWords ? 1 : 0, wx = lane - r * kPatchWords;
        const uint8_t *src = base + (size_t)r * bp + 4 * wx;
        uint32_t *dst = patch + lane;
#pragma unroll
        for (int it = 0; it < kStagePasses; ++it, src += 2 * bp)
            if (it < kStagePasses - 1 || r == 0) cp_async4(dst + it * 2 * kPatchWords, src);
    }
    cp_async_commit();
}

// One warp per kSlotsPerWarp (eight) keypoints.  ncu, round 1: the direct-gather version needed ~20 L1 wavefronts for
// each of the 16 scattered descriptor loads per lane; the 37x37 blurred patch is therefore staged in shared memory
// (cp.async, double buffered: the next keypoint's patch lands while this one's descriptor is computed) and gathered
// from there (<= 4-way bank conflicts).  Three phases per warp: (1) the moments of its keypoints, (2) fastAtan2 + the
// double-precision sincos ONCE, lane s working on keypoint s, (3) per keypoint the rotated BRIEF tests and the output
// row.  The kernel runs at 86 % of the L1 data-pipe wavefront peak, not at the issue limit.
template <int MINB, int kSlotsPerWarp, int kStage>
__global__ void __launch_bounds__(kDescWarps * 32, MINB)
k_describe(const __grid_constant__ Geo g, const uint8_t *__restrict__ pyr, const uint8_t *__restrict__ blur,
           const uint32_t *__restrict__ kept, const int *__restrict__ nkept,
           orbx_keypoint *__restrict__ out_kps, uint8_t *__restrict__ out_desc, int *__restrict__ out_counts,
           const uint32_t *__restrict__ pattern_words, const uint2 *__restrict__ moment_tab)
{
    constexpr int kPatchPitch = patch_pitch(kStage);
    extern __shared__ __align__(128) uint32_t patch_dyn[];      // [kDescWarps][2][patch_buf_words]: dynamic, the 64-byte-row variants pass 48 KB of static data
    uint32_t (*patch_all)[2][patch_buf_words(kStage)] = reinterpret_cast<uint32_t (*)[2][patch_buf_words(kStage)]>(patch_dyn);
    __shared__ __align__(8) uint64_t patch_bar[kDescWarps][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y + g.frame0;
    // lane i owns descriptor byte i = pattern points 16i .. 16i+15.  The pattern lives in shared memory as floats, one
    // float4 (x0, y0, x1, y1 of a test pair) per (pair k, lane): consecutive lanes read consecutive 16 bytes.  Held in
    // registers (8 packed words + the floats the compiler hoisted out of the slot loop) it cost 96 bytes of spills.
    __shared__ uint2 pat[8][32];                            // two half2 per entry: 8 bytes per lane, half the L1 wavefronts of float4
    __shared__ uint2 mtab[4 * kMomentPasses * 32];          // IC_Angle weights, see kMomentPasses
    if (threadIdx.x < 256) {
        const int k = threadIdx.x >> 5, l = threadIdx.x & 31;
        const uint32_t w = __ldg(pattern_words + l * 8 + k);
        const __half2 p0 = __floats2half2_rn((float)(int)(signed char)(w), (float)(int)(signed char)(w >> 8));
        const __half2 p1 = __floats2half2_rn((float)(int)(signed char)(w >> 16), (float)(int)(signed char)(w >> 24));
        pat[k][l] = make_uint2(*reinterpret_cast<const uint32_t *>(&p0), *reinterpret_cast<const uint32_t *>(&p1));
    }
    for (int i = threadIdx.x; i < 4 * kMomentPasses * 32; i += kDescWarps * 32) mtab[i] = __ldg(moment_tab + i);
    if (kStage == 2 && threadIdx.x < 2 * kDescWarps) {
        mbar_init(&patch_bar[0][0] + threadIdx.x, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // variant 4: the lane's eight pattern entries live in registers (16 of the 64 a 4-block launch bound allows; shared memory
    // holds 4 blocks per SM anyway): eight LDS.64 = 16 of ~200 L1 wavefronts per keypoint less
    constexpr bool kRegPat = kStage == 4;
    uint2 rpat[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) rpat[k] = kRegPat ? pat[k][lane] : make_uint2(0u, 0u);
    // per-level keypoint counts -> inclusive prefix in lanes 0..nlevels-1 (level-major concatenation, :1036-1063)
    const int myc = lane < g.nlevels ? nkept[f * g.nlevels + lane] : 0;
    int incl = myc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (blockIdx.x == 0 && threadIdx.x == 0) out_counts[f] = total;
    const int slot0 = (blockIdx.x * kDescWarps + warp) * kSlotsPerWarp;   // first output row of this warp inside the frame
    const int nslot = min(kSlotsPerWarp, total - slot0);
    if (nslot <= 0) return;

    // lane s < nslot looks up keypoint s: its level from the prefix sums, then its packed key (one load latency for all four)
    uint32_t my_key = 0; int my_level = 0;
    {
        const int slot = slot0 + min(lane, nslot - 1);
        int first = 0;
        for (int l = 0; l < g.nlevels; ++l) {
            const int e = __shfl_sync(0xffffffffu, incl, l);
            if (slot >= e) { my_level = l + 1; first = e; }
        }
        my_key = __ldg(kept + (size_t)f * g.kept_total + g.lv[my_level].kept_base + (slot - first));
    }

    // Per-slot records, written once by lane s for keypoint s and read back by the whole warp with ONE broadcast load each: the
    // window / patch addresses used to be rebuilt by every lane for every keypoint from two shuffles (key, level) and three
    // constant-bank loads of the level's geometry -- all of them slots on the pipe the kernel is bound by (ncu: L1 wavefronts).
    __shared__ uint4 rec_win[kDescWarps][kDescSlots], rec_patch[kDescWarps][kDescSlots];   // {address lo, hi, pitch, alignment / x - xa}
    __shared__ float2 rec_ab[kDescWarps][kDescSlots];
    if (lane < nslot) {
        const LevelGeom &Lk = g.lv[my_level];
        const int x = cand_x(my_key) + kMinBorder, y = cand_y(my_key) + kMinBorder;   // :801-802
        const int al = (x - kHalfPatch + kPadX) & 3;       // bytes between the aligned first word and column x-15
        const unsigned long long w = (unsigned long long)(pyr + Lk.base + (size_t)f * Lk.frame_stride + (size_t)(kPadY + y - kHalfPatch) * Lk.pitch
                                                          + (kPadX + x - kHalfPatch - al));
        rec_win[warp][lane] = make_uint4((uint32_t)w, (uint32_t)(w >> 32), (uint32_t)Lk.pitch, (uint32_t)al);
        const int xa = (x - kPatchR) & patch_align_mask(kStage);   // first staged column (keypoints sit >= 19 px inside)
        const unsigned long long pb = (unsigned long long)(blur + Lk.blur_base + (size_t)f * Lk.blur_frame_stride + (size_t)(y - kPatchR) * Lk.blur_pitch + xa);
        rec_patch[warp][lane] = make_uint4((uint32_t)pb, (uint32_t)(pb >> 32), (uint32_t)Lk.blur_pitch, (uint32_t)(x - xa));
    }
    __syncwarp();
    auto rec_ptr = [](const uint4 &r) { return reinterpret_cast<const uint8_t *>((unsigned long long)r.x | ((unsigned long long)r.y << 32)); };

    // ---- phase 1: IC_Angle moments on the un-blurred level (:27-54) for the warp's keypoints; lane s keeps keypoint s.
    //      The eleven pixel words of keypoint s+1 are loaded before the dot products of keypoint s. ----
    const int mrow = lane / 9, mword = lane - mrow * 9;    // lanes 27..31 read a fourth row with zero weights
    float my_m01 = 0.f, my_m10 = 0.f;
    uint32_t px[kMomentPasses], nx[kMomentPasses];
    int al = 0, nal = 0;
    auto load_window = [&](int si, uint32_t (&dst)[kMomentPasses], int &al_out, bool stage) {
        if (stage) { const uint4 rp = rec_patch[warp][0]; stage_patch<kStage>(patch_all[warp][0], rec_ptr(rp), (int)rp.z, lane, &patch_bar[warp][0]); }
        const uint4 rw = rec_win[warp][si];
        const int pitch = (int)rw.z;
        al_out = (int)rw.w;
        const uint8_t *src = rec_ptr(rw) + (size_t)mrow * pitch + 4 * mword;
        // only the 31 rows of the window are read (lanes 27..31 and the last pass's rows 31, 32 carry zero weights): every row a
        // load instruction touches is one more L1 wavefront, and the kernel sits on the L1 wavefront ceiling
#pragma unroll
        for (int it = 0; it < kMomentPasses; ++it, src += 3 * pitch)
            dst[it] = (lane < 27 && 3 * it + mrow < 2 * kHalfPatch + 1) ? __ldg(reinterpret_cast<const uint32_t *>(src)) : 0u;
    };
    load_window(0, px, al, true);
#pragma unroll 1
    for (int si = 0; si < nslot; ++si) {
        if (si + 1 < nslot) load_window(si + 1, nx, nal, false);
        const uint2 *tab = mtab + al * (kMomentPasses * 32) + lane;
        int m10 = 0, m01 = 0;
#pragma unroll
        for (int it = 0; it < kMomentPasses; ++it) {
            const uint2 wt = tab[it * 32];
            m10 = dp4a_u8s8(px[it], wt.x, m10);
            m01 = dp4a_u8s8(px[it], wt.y, m01);
        }
        // warp sums in one instruction each (REDUX.SUM) instead of two five-step shuffle ladders (ncu: 7.7 % of the kernel's stall samples)
        m10 = __reduce_add_sync(0xffffffffu, m10); m01 = __reduce_add_sync(0xffffffffu, m01);
        if (lane == si) { my_m01 = (float)m01; my_m10 = (float)m10; }
#pragma unroll
        for (int it = 0; it < kMomentPasses; ++it) px[it] = nx[it];
        al = nal;
    }

    // ---- phase 2: angle and rotation of keypoint s in lane s ----
    const float my_angle = fast_atan2_deg(my_m01, my_m10);
    float my_a, my_b;
    {
        const float factorPI = (float)(3.14159265358979323846 / (double)180.f);
        const float ang = __fmul_rn(my_angle, factorPI);
        double sd, cd;
        sincos((double)ang, &sd, &cd);                       // one range reduction for both
        my_a = (float)cd; my_b = (float)sd;
    }

    // the keypoint records of the warp's slots, lane s writing slot s: seven stores per warp instead of seven per keypoint
    if (lane < nslot) {
        rec_ab[warp][lane] = make_float2(my_a, my_b);
        const LevelGeom &Lk = g.lv[my_level];
        orbx_keypoint kp;
        kp.x = (float)(cand_x(my_key) + kMinBorder); kp.y = (float)(cand_y(my_key) + kMinBorder);               // :801-802
        if (my_level != 0) { kp.x = __fmul_rn(kp.x, Lk.scale); kp.y = __fmul_rn(kp.y, Lk.scale); }   // :1055-1061
        kp.size = (float)Lk.patch_size;
        kp.angle = my_angle;
        kp.response = (float)cand_score(my_key);
        kp.octave = my_level;
        kp.class_id = -1;
        out_kps[(size_t)f * g.capacity + slot0 + lane] = kp;
    }

    // ---- phase 3: rotated BRIEF, lane i produces descriptor byte i from its 16 pattern points ----
    __syncwarp();
    uint4 nrec = rec_patch[warp][0];
#pragma unroll 1
    for (int si = 0; si < nslot; ++si) {
        const int slot = slot0 + si;
        const float2 ab = rec_ab[warp][si];
        const float a = ab.x, b = ab.y;
        const int dx = (int)nrec.w;                          // x - xa of this keypoint (its patch was staged one iteration ago)
        if (si + 1 < nslot) {                                // next keypoint's patch into the other buffer
            nrec = rec_patch[warp][si + 1];
            stage_patch<kStage>(patch_all[warp][(si + 1) & 1], rec_ptr(nrec), (int)nrec.z, lane, &patch_bar[warp][(si + 1) & 1]);
            if (kStage != 2) cp_async_wait<1>();
        } else {
            if (kStage != 2) cp_async_wait<0>();
        }
        if (kStage == 2) mbar_wait(&patch_bar[warp][si & 1], (si >> 1) & 1);   // buffer b is filled for keypoints b, b+2, ...: parity of its use count
        __syncwarp();
        const uint8_t *pc = reinterpret_cast<const uint8_t *>(patch_all[warp][si & 1]) + kPatchR * kPatchPitch + dx;   // patch centre
        int val = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint2 pp = kRegPat ? rpat[k] : pat[k][lane];     // pattern offsets are integers in [-13, 13]: exact in half
            if (kRegPat) asm volatile("" : "+r"(pp.x), "+r"(pp.y));   // keeps the half -> float conversions inside the loop (hoisted they cost 32 registers)
            const float2 q0 = __half22float2(*reinterpret_cast<const __half2 *>(&pp.x)), q1 = __half22float2(*reinterpret_cast<const __half2 *>(&pp.y));
            const float x0 = q0.x, y0 = q0.y, x1 = q1.x, y1 = q1.y;
            const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
            const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
            const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
            const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
            const int t0 = pc[r0 * kPatchPitch + c0], t1 = pc[r1 * kPatchPitch + c1];
            val |= (t0 < t1) << k;
        }
        out_desc[((size_t)f * g.capacity + slot) * 32 + lane] = (uint8_t)val;

        __syncwarp();                                        // this buffer is the staging target of the next iteration
    }
}

struct DescribeTables { uint32_t *pattern; uint2 *moments; };
static DescribeTables g_desc_tab[64] = {};   // per device: the pattern as packed words, the IC_Angle weight table

void launch_describe(const Geo &g, const DevBuffers &b, int nframes, orbx_keypoint *d_kps, uint8_t *d_desc, int *d_counts, cudaStream_t s)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64) dev = 63;
    static std::mutex pattern_mutex;                      // handles are per-thread objects; this table is process-wide
    std::lock_guard<std::mutex> lock(pattern_mutex);
    if (!g_desc_tab[dev].pattern) {
        static const signed char h_pattern[1024] = {
#include "orb_pattern_31.inc"
        };
        uint32_t *p = nullptr;
        cudaMalloc(&p, 1024);
        cudaMemcpy(p, h_pattern, 1024, cudaMemcpyHostToDevice);
        // IC_Angle weights: entry [al][pass][lane] = {u weights, v weights} of the word at row 3*pass + lane/9,
        // word lane%9 of the window whose first byte is column -15 - al (umax is fixed by HALF_PATCH_SIZE = 15)
        static uint32_t h_tab[kMomentTabWords];
        for (int al = 0; al < 4; ++al)
            for (int it = 0; it < kMomentPasses; ++it)
                for (int l = 0; l < 32; ++l) {
                    const int r = 3 * it + l / 9, w = l % 9, v = r - kHalfPatch;
                    uint32_t wu = 0, wv = 0;
                    for (int j = 0; j < 4; ++j) {
                        const int u = 4 * w + j - al - kHalfPatch;
                        const int av = v < 0 ? -v : v, au = u < 0 ? -u : u;
                        if (l < 27 && av <= kHalfPatch && au <= g.umax[av]) {
                            wu |= (uint32_t)(uint8_t)(signed char)u << (8 * j);
                            wv |= (uint32_t)(uint8_t)(signed char)v << (8 * j);
                        }
                    }
                    h_tab[((al * kMomentPasses + it) * 32 + l) * 2] = wu;
                    h_tab[((al * kMomentPasses + it) * 32 + l) * 2 + 1] = wv;
                }
        uint2 *m = nullptr;
        cudaMalloc(&m, sizeof(h_tab));
        cudaMemcpy(m, h_tab, sizeof(h_tab), cudaMemcpyHostToDevice);
        g_desc_tab[dev].pattern = p; g_desc_tab[dev].moments = m;
    }
    // small batches (the low-latency path): two keypoints per warp instead of eight -- four times the warps, a quarter of
    // the serial depth; the amortisation of the pattern load and the sincos only matters when the GPU is full anyway
    const bool small = nframes <= 4;
    const int per_block = kDescWarps * (small ? 2 : kDescSlots);
    dim3 grd((g.capacity + per_block - 1) / per_block, nframes);
    // measured on a B200 (256 VGA frames): the kernel is bound by the L1 data pipe, and a small L1 (the default carve-out
    // maximises shared memory: 28 KB of L1 left) costs 30 %.  Round 2, after the shuffle ladders were gone: 141 KB of shared
    // memory = 3 resident blocks and ~85 KB of L1 beat 164 KB = 4 blocks (0.2134 against 0.2195 ms; 114 KB = 2 blocks 0.2176),
    // and with three blocks the 64 registers of variant 4 (pattern entries in registers) cost nothing: 0.2108 ms.
    static bool configured[64] = {};
    static int stage = 4, carve = 62;
    if (!configured[dev]) {
        if (const char *e = std::getenv("ORBX_DESC_STAGE")) stage = std::atoi(e);
        if (const char *e = std::getenv("ORBX_DESC_CARVEOUT")) carve = std::atoi(e);
        cudaFuncSetAttribute(k_describe<5, kDescSlots, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_describe<5, kDescSlots, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_describe<5, kDescSlots, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_describe<5, kDescSlots, 3>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_describe<5, 2, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 72);
        cudaFuncSetAttribute(k_describe<4, kDescSlots, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
        cudaFuncSetAttribute(k_describe<5, kDescSlots, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDescWarps * 2 * patch_buf_words(1) * 4);
        cudaFuncSetAttribute(k_describe<5, kDescSlots, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDescWarps * 2 * patch_buf_words(2) * 4);
        configured[dev] = true;
    }
#define DESC_ARGS g, b.pyr, b.blur, b.kept, b.nkept, d_kps, d_desc, d_counts, g_desc_tab[dev].pattern, g_desc_tab[dev].moments
    const int dyn0 = kDescWarps * 2 * patch_buf_words(0) * 4, dyn1 = kDescWarps * 2 * patch_buf_words(1) * 4;
    if (small) k_describe<5, 2, 0><<<grd, kDescWarps * 32, dyn0, s>>>(DESC_ARGS);
    else if (stage == 1) k_describe<5, kDescSlots, 1><<<grd, kDescWarps * 32, dyn1, s>>>(DESC_ARGS);
    else if (stage == 2) k_describe<5, kDescSlots, 2><<<grd, kDescWarps * 32, dyn1, s>>>(DESC_ARGS);
    else if (stage == 3) k_describe<5, kDescSlots, 3><<<grd, kDescWarps * 32, kDescWarps * 2 * patch_buf_words(3) * 4, s>>>(DESC_ARGS);
    else if (stage == 4) k_describe<4, kDescSlots, 4><<<grd, kDescWarps * 32, dyn0, s>>>(DESC_ARGS);
    else k_describe<5, kDescSlots, 0><<<grd, kDescWarps * 32, dyn0, s>>>(DESC_ARGS);
#undef DESC_ARGS
}

// ---------------------------------------------------------------------------------------------
// Frame::UndistortedKeyPoints (src/Frame.cpp:80-109) = cv::undistortPoints with R = I, P = K: five fixed
// iterations of the inverse Brown model in double precision (OpenCV 4.13.0 cvUndistortPointsInternal).
// Explicit round-to-nearest intrinsics keep the operation order of the scalar C++ code (no FMA contraction).
// ---------------------------------------------------------------------------------------------
__global__ void k_undistort(const orbx_keypoint *__restrict__ in, orbx_keypoint *__restrict__ out, int n,
                            double fx, double fy, double cx, double cy, double k1, double k2, double p1, double p2, double k3,
                            int passthrough, int literal_bug)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    orbx_keypoint k = in[i];
    if (!passthrough) {
        double x = __ddiv_rn(__dsub_rn((double)k.x, cx), fx), y = __ddiv_rn(__dsub_rn((double)k.y, cy), fy);
        const double x0 = x, y0 = y;
        for (int it = 0; it < 5; ++it) {
            const double r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
            const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k3, r2), k2), r2), k1), r2));
            const double icdist = __ddiv_rn(1.0, den);
            if (icdist < 0) { x = x0; y = y0; break; }
            const double dX = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, p1), x), y), __dmul_rn(p2, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, x), x))));
            const double dY = __dadd_rn(__dmul_rn(p1, __dadd_rn(r2, __dmul_rn(__dmul_rn(2.0, y), y))), __dmul_rn(__dmul_rn(__dmul_rn(2.0, p2), x), y));
            x = __dmul_rn(__dsub_rn(x0, dX), icdist); y = __dmul_rn(__dsub_rn(y0, dY), icdist);
        }
        const float ux = (float)__dadd_rn(__dmul_rn(x, fx), cx), uy = (float)__dadd_rn(__dmul_rn(y, fy), cy);
        k.x = ux; k.y = literal_bug ? ux : uy;
    }
    out[i] = k;
}

void launch_undistort(const orbx_keypoint *in, orbx_keypoint *out, int n, const float *cam, const float *dist, int literal_bug, cudaStream_t s)
{
    if (n <= 0) return;
    k_undistort<<<(n + 255) / 256, 256, 0, s>>>(in, out, n, cam[0], cam[1], cam[2], cam[3], dist[0], dist[1], dist[2], dist[3], dist[4],
                                               dist[0] == 0.0f, literal_bug);
}

} // namespace orbx

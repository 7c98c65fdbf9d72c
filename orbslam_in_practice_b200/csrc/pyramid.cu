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

// ---------------------------------------------------------------------------------------------
// level 0: input frames -> bordered level-0 buffers.  One thread = one aligned 16-byte chunk of a
// padded destination row (the interior starts at byte kPadX = 32 of a 64-byte aligned pitch).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_level0(const __grid_constant__ Geo g, uint8_t *__restrict__ pyr, const uint8_t *__restrict__ imgs,
         size_t in_pitch, size_t in_fstride, int src_aligned)
{
    const LevelGeom &L = g.lv[0];
    const int B = g.border_on ? kBorder : kMinBlurBorder;
    const int Y = (int)(blockIdx.y * 4 + threadIdx.y) - B;   // bordered row
    const int f = blockIdx.z;
    const int chunk = blockIdx.x * 32 + threadIdx.x;
    const int X0 = chunk * 16 - kPadX;                       // first pixel of this chunk
    if (chunk * 16 >= L.pitch || X0 + 15 < -B || X0 >= L.w + B || Y >= L.h + B) return;
    const int y = reflect101(Y, L.h);
    const uint8_t *src = imgs + (size_t)f * in_fstride + (size_t)y * in_pitch;
    uint8_t *dst = pyr + L.base + (size_t)f * L.frame_stride + (size_t)(Y + kPadY) * L.pitch + (size_t)chunk * 16;
    if (X0 >= 0 && X0 + 15 < L.w && src_aligned) {
        *reinterpret_cast<uint4 *>(dst) = __ldg(reinterpret_cast<const uint4 *>(src + X0));
        return;
    }
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int X = X0 + 4 * j + k;
            const int x = reflect101(min(max(X, -B), L.w + B - 1), L.w);
            v |= (uint32_t)__ldg(src + x) << (8 * k);
        }
        w[j] = v;
    }
    *reinterpret_cast<uint4 *>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------------------------------------
// level l >= 1 from level l-1.  One thread = 4 consecutive pixels of a padded destination row.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t resize_px(const uint8_t *__restrict__ r0, const uint8_t *__restrict__ r1,
                                              int2 tx, int sw, int cy0, int cy1)
{
    const int sx0 = tx.x, sx1 = min(sx0 + 1, sw - 1);
    const int cx0 = (short)(tx.y & 0xffff), cx1 = tx.y >> 16;
    const int h0 = r0[sx0] * cx0 + r0[sx1] * cx1;
    const int h1 = r1[sx0] * cx0 + r1[sx1] * cx1;
    const int v = (((cy0 * (h0 >> 4)) >> 16) + ((cy1 * (h1 >> 4)) >> 16) + 2) >> 2;
    return (uint32_t)min(max(v, 0), 255);
}

__global__ void __launch_bounds__(128)
k_resize(const __grid_constant__ Geo g, uint8_t *__restrict__ pyr, const int2 *__restrict__ tables, int level)
{
    const LevelGeom &D = g.lv[level];
    const LevelGeom &S = g.lv[level - 1];
    const int B = g.border_on ? kBorder : kMinBlurBorder;
    const int Y = (int)(blockIdx.y * 4 + threadIdx.y) - B;
    const int f = blockIdx.z;
    const int chunk = blockIdx.x * 32 + threadIdx.x;
    const int X0 = chunk * 4 - kPadX;
    if (chunk * 4 >= D.pitch || X0 + 3 < -B || X0 >= D.w + B || Y >= D.h + B) return;
    const int y = reflect101(Y, D.h);
    const int2 ty = __ldg(tables + D.taby + y);
    const int sy0 = min(max(ty.x, 0), S.h - 1), sy1 = min(max(ty.x + 1, 0), S.h - 1);
    const int cy0 = (short)(ty.y & 0xffff), cy1 = ty.y >> 16;
    const uint8_t *src = pyr + S.base + (size_t)f * S.frame_stride + (size_t)kPadY * S.pitch + kPadX;
    const uint8_t *r0 = src + (size_t)sy0 * S.pitch, *r1 = src + (size_t)sy1 * S.pitch;
    const int2 *tabx = tables + D.tabx;
    uint32_t out = 0;
    if (X0 >= 0 && X0 + 3 < D.w) {
        // interior chunk: the four table entries are one aligned 32-byte run
        const int4 t01 = __ldg(reinterpret_cast<const int4 *>(tabx + X0));
        const int4 t23 = __ldg(reinterpret_cast<const int4 *>(tabx + X0) + 1);
        out = resize_px(r0, r1, make_int2(t01.x, t01.y), S.w, cy0, cy1) |
              (resize_px(r0, r1, make_int2(t01.z, t01.w), S.w, cy0, cy1) << 8) |
              (resize_px(r0, r1, make_int2(t23.x, t23.y), S.w, cy0, cy1) << 16) |
              (resize_px(r0, r1, make_int2(t23.z, t23.w), S.w, cy0, cy1) << 24);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int X = min(max(X0 + k, -B), D.w + B - 1);
            const int x = reflect101(X, D.w);
            out |= resize_px(r0, r1, __ldg(tabx + x), S.w, cy0, cy1) << (8 * k);
        }
    }
    *reinterpret_cast<uint32_t *>(pyr + D.base + (size_t)f * D.frame_stride + (size_t)(Y + kPadY) * D.pitch + (size_t)chunk * 4) = out;
}

void launch_level0(const Geo &g, const DevBuffers &b, const uint8_t *d_imgs, size_t pitch, size_t fstride, int nframes, cudaStream_t s)
{
    const int B = g.border_on ? kBorder : kMinBlurBorder;
    const int chunks = g.lv[0].pitch / 16;
    const int aligned = ((((uintptr_t)d_imgs) | pitch | fstride) & 15) == 0;
    dim3 grd((chunks + 31) / 32, (g.lv[0].h + 2 * B + 3) / 4, nframes);
    k_level0<<<grd, dim3(32, 4), 0, s>>>(g, b.pyr, d_imgs, pitch, fstride, aligned);
}

void launch_resize(const Geo &g, const DevBuffers &b, int level, int nframes, cudaStream_t s)
{
    const int B = g.border_on ? kBorder : kMinBlurBorder;
    const int chunks = g.lv[level].pitch / 4;
    dim3 grd((chunks + 31) / 32, (g.lv[level].h + 2 * B + 3) / 4, nframes);
    k_resize<<<grd, dim3(32, 4), 0, s>>>(g, b.pyr, b.tables, level);
}

} // namespace orbx

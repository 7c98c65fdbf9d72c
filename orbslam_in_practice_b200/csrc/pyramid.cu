// pyramid.cu -- ComputePyramid (ORBextractor.cpp:1071-1096) on the device.
//
// Level 0 is the input copied into a bordered buffer; level l is cv::resize(INTER_LINEAR) of the
// ROUNDED level l-1 (a sequential chain of 7 launches per batch), reproduced in OpenCV's 11-bit
// fixed point.  Each launch covers every frame of the batch.  The output
// domain is the bordered image: a border pixel evaluates the resize at its reflect-101 source
// coordinate, so copyMakeBorder costs no extra pass and no divergent fix-up.
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
// level 0: input frames -> bordered level-0 buffers.  4 pixels per thread in the interior rows
// would need aligned input pitches we do not control; a byte-wise copy is HBM-trivial here
// (0.66 MB per VGA frame) and is kept simple.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_level0(const __grid_constant__ Geo g, uint8_t *__restrict__ pyr, const uint8_t *__restrict__ imgs,
         size_t in_pitch, size_t in_fstride)
{
    const LevelGeom &L = g.lv[0];
    const int B = g.border_on ? kBorder : 0;
    const int X = blockIdx.x * 32 + threadIdx.x - B;
    const int Y = blockIdx.y * 8 + threadIdx.y - B;
    const int f = blockIdx.z;
    if (X >= L.w + B || Y >= L.h + B) return;
    const int x = reflect101(X, L.w), y = reflect101(Y, L.h);
    const uint8_t v = __ldg(imgs + (size_t)f * in_fstride + (size_t)y * in_pitch + x);
    pyr[L.base + (size_t)f * L.frame_stride + (size_t)(Y + kPadY) * L.pitch + kPadX + X] = v;
}

// ---------------------------------------------------------------------------------------------
// level l >= 1 from level l-1
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_resize(const __grid_constant__ Geo g, uint8_t *__restrict__ pyr, const int2 *__restrict__ tables, int level)
{
    const LevelGeom &D = g.lv[level];
    const LevelGeom &S = g.lv[level - 1];
    const int B = g.border_on ? kBorder : 0;
    const int X = blockIdx.x * 32 + threadIdx.x - B;
    const int Y = blockIdx.y * 8 + threadIdx.y - B;
    const int f = blockIdx.z;
    if (X >= D.w + B || Y >= D.h + B) return;
    const int x = reflect101(X, D.w), y = reflect101(Y, D.h);
    const int2 tx = __ldg(tables + D.tabx + x);
    const int2 ty = __ldg(tables + D.taby + y);
    const int sx0 = tx.x, sx1 = min(sx0 + 1, S.w - 1);
    const int cx0 = (short)(tx.y & 0xffff), cx1 = (short)(tx.y >> 16);
    const int sy0 = min(max(ty.x, 0), S.h - 1), sy1 = min(max(ty.x + 1, 0), S.h - 1);
    const int cy0 = (short)(ty.y & 0xffff), cy1 = (short)(ty.y >> 16);
    const uint8_t *src = pyr + S.base + (size_t)f * S.frame_stride + (size_t)kPadY * S.pitch + kPadX;
    const uint8_t *r0 = src + (size_t)sy0 * S.pitch, *r1 = src + (size_t)sy1 * S.pitch;
    const int h0 = r0[sx0] * cx0 + r0[sx1] * cx1;
    const int h1 = r1[sx0] * cx0 + r1[sx1] * cx1;
    int v = (((cy0 * (h0 >> 4)) >> 16) + ((cy1 * (h1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
    pyr[D.base + (size_t)f * D.frame_stride + (size_t)(Y + kPadY) * D.pitch + kPadX + X] = (uint8_t)v;
}

void launch_level0(const Geo &g, const DevBuffers &b, const uint8_t *d_imgs, size_t pitch, size_t fstride, int nframes, cudaStream_t s)
{
    const int B = g.border_on ? kBorder : 0;
    dim3 blk(32, 8), grd((g.lv[0].w + 2 * B + 31) / 32, (g.lv[0].h + 2 * B + 7) / 8, nframes);
    k_level0<<<grd, blk, 0, s>>>(g, b.pyr, d_imgs, pitch, fstride);
}

void launch_resize(const Geo &g, const DevBuffers &b, int level, int nframes, cudaStream_t s)
{
    const int B = g.border_on ? kBorder : 0;
    dim3 blk(32, 8), grd((g.lv[level].w + 2 * B + 31) / 32, (g.lv[level].h + 2 * B + 7) / 8, nframes);
    k_resize<<<grd, blk, 0, s>>>(g, b.pyr, b.tables, level);
}

} // namespace orbx

// describe.cu -- GaussianBlur (ORBextractor.cpp:1045-1046), computeOrientation / IC_Angle
// (:27-54, :422-429), computeOrbDescriptor (:57-97) and the operator() epilogue (:1036-1064).
#include "orbx_internal.cuh"

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
// 7x7 sigma-2 Gaussian in OpenCV's Q8.8 fixed point: taps 18,34,48,56,48,34,18;
// horizontal pass exact in u16, vertical pass u32, dst = (v + 32768) >> 16.  Border: reflect-101
// of the level itself (the blur runs on a border-less clone in the reference).
// One block = 64 x 16 output tile, all levels and frames in one launch (tile table by level).
// ---------------------------------------------------------------------------------------------
constexpr int kBlurTW = 64, kBlurTH = 16;

__global__ void __launch_bounds__(256)
k_blur(const __grid_constant__ Geo g, const uint8_t *__restrict__ pyr, uint8_t *__restrict__ blur, int level)
{
    __shared__ uint8_t src[(kBlurTH + 6)][kBlurTW + 8];
    __shared__ uint16_t hor[(kBlurTH + 6)][kBlurTW];
    const LevelGeom &L = g.lv[level];
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * kBlurTW, y0 = blockIdx.y * kBlurTH;
    const uint8_t *img = pyr + L.base + (size_t)f * L.frame_stride + (size_t)kPadY * L.pitch + kPadX;
    const int tid = threadIdx.x;
    for (int i = tid; i < (kBlurTH + 6) * (kBlurTW + 6); i += 256) {
        const int r = i / (kBlurTW + 6), c = i - r * (kBlurTW + 6);
        const int y = reflect101d(y0 + r - 3, L.h), x = reflect101d(x0 + c - 3, L.w);
        src[r][c] = img[(size_t)y * L.pitch + x];
    }
    __syncthreads();
    for (int i = tid; i < (kBlurTH + 6) * kBlurTW; i += 256) {
        const int r = i / kBlurTW, c = i - r * kBlurTW;
        const uint8_t *p = &src[r][c];
        hor[r][c] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    uint8_t *dst = blur + L.blur_base + (size_t)f * L.blur_frame_stride;
    for (int i = tid; i < kBlurTH * kBlurTW; i += 256) {
        const int r = i / kBlurTW, c = i - r * kBlurTW;
        const int x = x0 + c, y = y0 + r;
        if (x >= L.w || y >= L.h) continue;
        const uint32_t v = 18u * (hor[r][c] + hor[r + 6][c]) + 34u * (hor[r + 1][c] + hor[r + 5][c]) +
                           48u * (hor[r + 2][c] + hor[r + 4][c]) + 56u * hor[r + 3][c];
        dst[(size_t)y * L.blur_pitch + x] = (uint8_t)((v + 32768u) >> 16);
    }
}

void launch_blur(const Geo &g, const DevBuffers &b, int nframes, cudaStream_t s)
{
    for (int l = 0; l < g.nlevels; ++l) {
        dim3 grd((g.lv[l].w + kBlurTW - 1) / kBlurTW, (g.lv[l].h + kBlurTH - 1) / kBlurTH, nframes);
        k_blur<<<grd, 256, 0, s>>>(g, b.pyr, b.blur, l);
    }
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
constexpr int kSlotsPerWarp = 4;          // keypoints handled by one warp (amortises the pattern load)

__global__ void __launch_bounds__(kDescWarps * 32)
k_describe(const __grid_constant__ Geo g, const uint8_t *__restrict__ pyr, const uint8_t *__restrict__ blur,
           const uint32_t *__restrict__ kept, const int *__restrict__ nkept,
           orbx_keypoint *__restrict__ out_kps, uint8_t *__restrict__ out_desc, int *__restrict__ out_counts,
           const uint32_t *__restrict__ pattern_words)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.y;
    // lane i owns descriptor byte i = pattern points 16i .. 16i+15 = 32 signed bytes = 8 words
    uint32_t pw[8];
    {
        const uint4 p0 = __ldg(reinterpret_cast<const uint4 *>(pattern_words) + lane * 2);
        const uint4 p1 = __ldg(reinterpret_cast<const uint4 *>(pattern_words) + lane * 2 + 1);
        pw[0] = p0.x; pw[1] = p0.y; pw[2] = p0.z; pw[3] = p0.w; pw[4] = p1.x; pw[5] = p1.y; pw[6] = p1.z; pw[7] = p1.w;
    }
    const int *nk = nkept + f * g.nlevels;
    int total = 0;
#pragma unroll 1
    for (int l = 0; l < g.nlevels; ++l) total += nk[l];
    if (blockIdx.x == 0 && threadIdx.x == 0) out_counts[f] = total;

    const int slot0 = (blockIdx.x * kDescWarps + warp) * kSlotsPerWarp;
#pragma unroll 1
    for (int si = 0; si < kSlotsPerWarp; ++si) {
        const int slot = slot0 + si;                          // output row inside the frame
        if (slot >= total) return;
        // level-major concatenation (:1036-1063): find the level this slot falls in
        int level = 0, first = 0;
#pragma unroll 1
        for (int l = 0, acc = 0; l < g.nlevels; ++l) {
            const int c = nk[l];
            if (slot >= acc && slot < acc + c) { level = l; first = acc; }
            acc += c;
        }
        const LevelGeom &L = g.lv[level];
        const uint32_t key = kept[(size_t)f * g.kept_total + L.kept_base + (slot - first)];
        const int x = cand_x(key) + kMinBorder, y = cand_y(key) + kMinBorder;   // :801-802

        // ---- IC_Angle on the un-blurred level: lane = column u, all 31 row loads in flight ----
        const int pitch = L.pitch;
        const uint8_t *img = pyr + L.base + (size_t)f * L.frame_stride + (size_t)(kPadY + y) * pitch + kPadX + x;
        int m10 = 0, m01 = 0;
        const int u = lane - kHalfPatch;
        const int au = u < 0 ? -u : u;
        if (lane < 2 * kHalfPatch + 1) {
            int colsum = 0;
#pragma unroll
            for (int v = -kHalfPatch; v <= kHalfPatch; ++v) {
                if (au <= g.umax[v < 0 ? -v : v]) {
                    const int I = img[v * pitch + u];
                    colsum += I; m01 += v * I;
                }
            }
            m10 = u * colsum;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
        const float angle = fast_atan2_deg((float)m01, (float)m10);

        // ---- rotated BRIEF: lane i produces descriptor byte i from its 16 pattern points ----
        const float factorPI = (float)(3.14159265358979323846 / (double)180.f);
        const float ang = __fmul_rn(angle, factorPI);
        const float a = (float)cos((double)ang), b = (float)sin((double)ang);
        const int bp = L.blur_pitch;
        const uint8_t *center = blur + L.blur_base + (size_t)f * L.blur_frame_stride + (size_t)y * bp + x;
        int t[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t w = pw[k >> 1];
            const float px = (float)(int)(signed char)(w >> ((k & 1) * 16));
            const float py = (float)(int)(signed char)(w >> ((k & 1) * 16 + 8));
            const int r = __float2int_rn(__fadd_rn(__fmul_rn(px, b), __fmul_rn(py, a)));
            const int c = __float2int_rn(__fsub_rn(__fmul_rn(px, a), __fmul_rn(py, b)));
            t[k] = center[r * bp + c];
        }
        int val = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) val |= (t[2 * k] < t[2 * k + 1]) << k;
        out_desc[((size_t)f * g.capacity + slot) * 32 + lane] = (uint8_t)val;

        if (lane == 0) {
            orbx_keypoint kp;
            kp.x = (float)x; kp.y = (float)y;
            if (level != 0) { kp.x = __fmul_rn(kp.x, L.scale); kp.y = __fmul_rn(kp.y, L.scale); }   // :1055-1061
            kp.size = (float)L.patch_size;
            kp.angle = angle;
            kp.response = (float)cand_score(key);
            kp.octave = level;
            kp.class_id = -1;
            out_kps[(size_t)f * g.capacity + slot] = kp;
        }
    }
}

static uint32_t *g_pattern_dev[64] = { nullptr };   // per device copy of the pattern as packed words

void launch_describe(const Geo &g, const DevBuffers &b, int nframes, orbx_keypoint *d_kps, uint8_t *d_desc, int *d_counts, cudaStream_t s)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !g_pattern_dev[dev]) {
        static const signed char h_pattern[1024] = {
#include "orb_pattern_31.inc"
        };
        uint32_t *p = nullptr;
        cudaMalloc(&p, 1024);
        cudaMemcpy(p, h_pattern, 1024, cudaMemcpyHostToDevice);
        g_pattern_dev[dev] = p;
    }
    const int per_block = kDescWarps * kSlotsPerWarp;
    dim3 grd((g.capacity + per_block - 1) / per_block, nframes);
    k_describe<<<grd, kDescWarps * 32, 0, s>>>(g, b.pyr, b.blur, b.kept, b.nkept, d_kps, d_desc, d_counts, g_pattern_dev[dev]);
}

} // namespace orbx

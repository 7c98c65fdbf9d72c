// fast.cu -- the per-cell FAST-9 loop of ComputeKeyPointsOctTree (ORBextractor.cpp:745-786).
//
// One thread block per (cell, frame).  The block stages the cell's (wCell+6) x (hCell+6) u8 tile
// (3-px halo) in shared memory, scores every interior pixel once with threshold 0 semantics
// (SURVEY.md A3 identity: corner(t) <=> score0 >= t and score_t == score0 for corners), applies
// the strict 3x3 non-max suppression *inside the cell* for iniThFAST, and -- only if the cell is
// empty after NMS, exactly like the reference's vKeysCell.empty() retry -- for minThFAST.
// Survivors are compacted in row-major order into the cell's slot array; the octree kernel
// concatenates cells in row-major cell order, reproducing the order of vToDistributeKeys.
#include "orbx_internal.cuh"

namespace orbx {

constexpr int kFastThreads = 256;
constexpr int kMaxCellDim = 64;                 // wCell, hCell < 60 (width/30 cells of ceil size)
constexpr int kTileDim = kMaxCellDim + 6;

// circular 16-bit mask: any run of >= 9 set bits?
__device__ __forceinline__ bool has_arc9(uint32_t m)
{
    m |= m << 16;                       // unroll the circle
    uint32_t r = m & (m >> 1);          // runs >= 2
    r &= r >> 2;                        // >= 4
    r &= r >> 4;                        // >= 8
    r &= m >> 8;                        // >= 9
    return (r & 0xffffu) != 0;
}

// max over the 16 arcs of 9 contiguous ring pixels of min(d) -- cornerScore<16>'s 'a0' part.
__device__ __forceinline__ int arc9_maxmin(const int (&d)[16])
{
    int m2[16], m4[16], m8[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m2[k] = min(d[k], d[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m4[k] = min(m2[k], m2[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m8[k] = min(m4[k], m4[(k + 4) & 15]);
    int best = -512;
#pragma unroll
    for (int k = 0; k < 16; ++k) best = max(best, min(m8[k], d[(k + 8) & 15]));
    return best;
}

__global__ void __launch_bounds__(kFastThreads)
k_fast_cells(const __grid_constant__ Geo g, const uint8_t *__restrict__ pyr,
             int *__restrict__ cell_count, uint32_t *__restrict__ cell_slots)
{
    __shared__ uint8_t tile[kTileDim * kTileDim];
    __shared__ uint8_t score[(kMaxCellDim + 2) * (kMaxCellDim + 2)];
    __shared__ int warp_tot[kFastThreads / 32];

    const int cell = blockIdx.x, f = blockIdx.y;
    int level = 0;
#pragma unroll 1
    for (int l = 1; l < g.nlevels; ++l) if (cell >= g.lv[l].cell_base) level = l;
    const LevelGeom &L = g.lv[level];
    const int c = cell - L.cell_base;
    const int ci = c / L.nCols, cj = c - ci * L.nCols;
    int *count_out = cell_count + (size_t)f * g.total_cells + cell;

    // cell rectangle, ORBextractor.cpp:745-762 (all values are integers held in floats there)
    const int iniY = kMinBorder + ci * L.hCell, iniX = kMinBorder + cj * L.wCell;
    int maxY = iniY + L.hCell + 6, maxX = iniX + L.wCell + 6;
    if (iniY >= L.maxBorderY - 3 || iniX >= L.maxBorderX - 6) { if (threadIdx.x == 0) *count_out = 0; return; }
    maxY = min(maxY, L.maxBorderY); maxX = min(maxX, L.maxBorderX);
    const int cw = maxX - iniX, ch = maxY - iniY;      // cell image size handed to cv::FAST
    const int iw = cw - 6, ih = ch - 6;                // pixels FAST actually tests
    if (iw <= 0 || ih <= 0) { if (threadIdx.x == 0) *count_out = 0; return; }

    const uint8_t *img = pyr + L.base + (size_t)f * L.frame_stride + (size_t)(kPadY + iniY) * L.pitch + kPadX + iniX;
    for (int i = threadIdx.x; i < cw * ch; i += kFastThreads) {
        const int y = i / cw, x = i - y * cw;
        tile[y * kTileDim + x] = img[(size_t)y * L.pitch + x];
    }
    const int sw = iw + 2;                              // score map with a zero frame
    for (int i = threadIdx.x; i < sw * (ih + 2); i += kFastThreads) score[i] = 0;
    __syncthreads();

    const int minTh = g.min_th, iniTh = g.ini_th;
    const int npix = iw * ih;
    for (int i = threadIdx.x; i < npix; i += kFastThreads) {
        const int y = i / iw, x = i - y * iw;
        const uint8_t *p = tile + (y + 3) * kTileDim + (x + 3);
        const int v = p[0];
        int d[16];
        d[0] = v - p[3 * kTileDim];      d[1] = v - p[3 * kTileDim + 1];  d[2] = v - p[2 * kTileDim + 2];  d[3] = v - p[kTileDim + 3];
        d[4] = v - p[3];                 d[5] = v - p[-kTileDim + 3];     d[6] = v - p[-2 * kTileDim + 2]; d[7] = v - p[-3 * kTileDim + 1];
        d[8] = v - p[-3 * kTileDim];     d[9] = v - p[-3 * kTileDim - 1]; d[10] = v - p[-2 * kTileDim - 2]; d[11] = v - p[-kTileDim - 3];
        d[12] = v - p[-3];               d[13] = v - p[kTileDim - 3];     d[14] = v - p[2 * kTileDim - 2];  d[15] = v - p[3 * kTileDim - 1];
        uint32_t dark = 0, bright = 0;                  // ring darker / brighter than centre by > minTh
#pragma unroll
        for (int k = 0; k < 16; ++k) { dark |= (uint32_t)(d[k] > minTh) << k; bright |= (uint32_t)(d[k] < -minTh) << k; }
        int s = 0;
        if (has_arc9(dark) || has_arc9(bright)) {
            const int a = arc9_maxmin(d);
            int nd[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) nd[k] = -d[k];
            const int b = arc9_maxmin(nd);
            s = max(a, b) - 1;                          // = cornerScore (>= minTh for corners at minTh)
        }
        score[(y + 1) * sw + (x + 1)] = (uint8_t)s;
    }
    __syncthreads();

    // NMS at both thresholds; neighbours below the threshold (non-corners) and outside the cell count as 0
    uint32_t keep_ini = 0, keep_min = 0;               // bit t = this thread's t-th pixel
    int it = 0;
    for (int i = threadIdx.x; i < npix; i += kFastThreads, ++it) {
        const int y = i / iw, x = i - y * iw;
        const uint8_t *q = score + (y + 1) * sw + (x + 1);
        const int s = q[0];
        if (s < minTh) continue;
        int nmax_min = 0, nmax_ini = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (dx == 0 && dy == 0) continue;
                const int n = q[dy * sw + dx];
                nmax_min = max(nmax_min, n >= minTh ? n : 0);
                nmax_ini = max(nmax_ini, n >= iniTh ? n : 0);
            }
        if (s > nmax_min) keep_min |= 1u << it;
        if (s >= iniTh && s > nmax_ini) keep_ini |= 1u << it;
    }
    const int any_ini = __syncthreads_or(keep_ini != 0);
    const uint32_t keep = any_ini ? keep_ini : keep_min;

    // ordered compaction: pixel index i = it * 256 + tid is row-major, so process 'it' in order
    uint32_t *slots = cell_slots + (size_t)f * g.slots_per_frame + L.slot_base + (size_t)c * L.cell_cap;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int niter = (npix + kFastThreads - 1) / kFastThreads;
    int base = 0;
    for (int t = 0; t < niter; ++t) {
        const bool k = (keep >> t) & 1u;
        const uint32_t bal = __ballot_sync(0xffffffffu, k);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < kFastThreads / 32; ++w) { const int n = warp_tot[w]; if (w < warp) woff += n; tot += n; }
        if (k) {
            const int i = t * kFastThreads + threadIdx.x;
            const int y = i / iw, x = i - y * iw;
            const int pos = base + woff + __popc(bal & ((1u << lane) - 1u));
            // keypoint in level coordinates relative to (16,16): cell pixel (x+3, y+3) + (j*wCell, i*hCell)
            slots[pos] = pack_cand(x + 3 + cj * L.wCell, y + 3 + ci * L.hCell, score[(y + 1) * sw + (x + 1)]);
        }
        base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_out = base;
}

void launch_fast(const Geo &g, const DevBuffers &b, int nframes, cudaStream_t s)
{
    dim3 grd(g.total_cells, nframes);
    k_fast_cells<<<grd, kFastThreads, 0, s>>>(g, b.pyr, b.cell_count, b.cell_slots);
}

} // namespace orbx

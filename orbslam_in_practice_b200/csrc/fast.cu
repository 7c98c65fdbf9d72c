// fast.cu -- the per-cell FAST-9 loop of ComputeKeyPointsOctTree (ORBextractor.cpp:745-786).
//
// One thread block per (cell, frame).  The block stages the cell's (wCell+6) x (hCell+6) u8 tile
// (3-px halo) in shared memory and then works in four dense phases, because the expensive parts of
// FAST touch only a few percent of the pixels (ncu, round 1: the naive one-thread-per-pixel kernel
// spent 406 lane-instructions per pixel and was ALU-pipe bound at 79 %):
//   A  every interior pixel: 4-point compass test (every 9-arc of the 16-ring contains one pixel
//      of each opposite pair, so (N|S)&(E|W) must hold for one polarity); survivors (~9 %) are
//      pushed to a shared-memory queue with warp-aggregated atomics
//   B  queue, all lanes busy: full 16-ring arc test at min(iniTh, minTh) and, for corners (~5 %),
//      the exact cornerScore via a sliding-window min (SURVEY.md A3 identity: corner(t) <=>
//      score0 >= t and score_t == score0, so one score serves both thresholds)
//   C  corners only: strict 3x3 non-max suppression *inside the cell* at iniTh and at minTh
//      (neighbours below the threshold or outside the cell read 0) -> two bitmaps
//   D  the reference's retry: use the iniTh bitmap unless it is empty (vKeysCell.empty()), then
//      emit set bits in row-major order into the cell's slot array.
// The octree kernel concatenates cells in row-major cell order = the order of vToDistributeKeys.
#include "orbx_internal.cuh"

namespace orbx {

constexpr int kFastThreads = 128;
constexpr int kFastWarps = kFastThreads / 32;
// Shared memory is sized at launch from the largest cell of the current geometry (wCell, hCell
// <= 64 is checked in build_geometry; VGA levels need ~7 KB per block).
struct FastSmem { int tp, sp, npix_max; };      // tile pitch, score pitch, max interior pixels

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

// warp-aggregated push of (value) for lanes with pred set; returns nothing, order irrelevant
__device__ __forceinline__ void queue_push(bool pred, uint16_t value, uint16_t *queue, int *count)
{
    const uint32_t bal = __ballot_sync(0xffffffffu, pred);
    if (bal == 0) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(count, __popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (pred) queue[base + __popc(bal & ((1u << lane) - 1u))] = value;
}

__global__ void __launch_bounds__(kFastThreads)
k_fast_cells(const __grid_constant__ Geo g, const uint8_t *__restrict__ pyr,
             int *__restrict__ cell_count, uint32_t *__restrict__ cell_slots, const FastSmem sm, const int tile_rows)
{
    extern __shared__ __align__(16) unsigned char fast_smem[];
    __shared__ int q_count, c_count;
    const int kTP = sm.tp, kSP = sm.sp;
    uint8_t *tile = fast_smem;                                         // [tile_rows][tp]
    uint8_t *score = tile + tile_rows * kTP;                           // [tile_rows - 4][sp]
    uint16_t *queue = reinterpret_cast<uint16_t *>(score + ((tile_rows - 4) * kSP + 15) / 16 * 16);   // phase A survivors
    uint16_t *corners = queue + sm.npix_max;                           // phase B corners
    uint32_t *bm_ini = reinterpret_cast<uint32_t *>(corners + sm.npix_max);
    uint32_t *bm_min = bm_ini + (sm.npix_max + 31) / 32;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
    if (iniY >= L.maxBorderY - 3 || iniX >= L.maxBorderX - 6) { if (tid == 0) *count_out = 0; return; }
    maxY = min(maxY, L.maxBorderY); maxX = min(maxX, L.maxBorderX);
    const int cw = maxX - iniX, ch = maxY - iniY;      // cell image size handed to cv::FAST
    const int iw = cw - 6, ih = ch - 6;                // pixels FAST actually tests
    if (iw <= 0 || ih <= 0) { if (tid == 0) *count_out = 0; return; }

    // ---- phase 0: stage the tile, clear score map / bitmaps ----
    // The tile is filled with aligned 32-bit loads: its column 0 is the 4-byte aligned pixel at or
    // left of iniX (rows are 64-byte aligned and the interior starts at byte 32), so cell column c
    // lives at tile column c + xoff.  (Byte-wise staging was 24 % of this kernel's instructions.)
    const int xoff = iniX & 3;
    {
        const uint8_t *img = pyr + L.base + (size_t)f * L.frame_stride + (size_t)(kPadY + iniY) * L.pitch + kPadX + (iniX - xoff);
        const int wpr = (cw + xoff + 3) >> 2;               // words per tile row
        const int nw = wpr * ch;
        uint32_t *t32 = reinterpret_cast<uint32_t *>(tile);
        const int tpw = kTP >> 2;
        for (int i = tid; i < nw; i += kFastThreads) {
            const int r = i / wpr, wx = i - r * wpr;
            t32[r * tpw + wx] = __ldg(reinterpret_cast<const uint32_t *>(img + (size_t)r * L.pitch) + wx);
        }
    }
    {
        uint32_t *s32 = reinterpret_cast<uint32_t *>(score);
        const int nwords = ((ih + 2) * kSP + 3) >> 2;
        for (int i = tid; i < nwords; i += kFastThreads) s32[i] = 0;
        const int nbm = (iw * ih + 31) >> 5;
        for (int i = tid; i < nbm; i += kFastThreads) { bm_ini[i] = 0; bm_min[i] = 0; }
        if (tid == 0) { q_count = 0; c_count = 0; }
    }
    __syncthreads();

    const int minTh = g.min_th, iniTh = g.ini_th;
    const int lowTh = min(minTh, iniTh);

    // ---- phase A: compass pre-test on every interior pixel ----
    for (int y = warp; y < ih; y += kFastWarps) {
        const uint8_t *row = tile + (y + 3) * kTP + 3 + xoff;
        for (int x0 = 0; x0 < iw; x0 += 32) {
            const int x = x0 + lane;
            bool pass = false;
            if (x < iw) {
                const uint8_t *p = row + x;
                const int v = p[0];
                const int hi = v + lowTh, lo = v - lowTh;
                const int n = p[3 * kTP], s = p[-3 * kTP], e = p[3], w = p[-3];
                const bool bright = ((n > hi) | (s > hi)) & ((e > hi) | (w > hi));
                const bool dark = ((n < lo) | (s < lo)) & ((e < lo) | (w < lo));
                pass = bright | dark;
            }
            queue_push(pass, (uint16_t)((y << 8) | x), queue, &q_count);
        }
    }
    __syncthreads();

    // ---- phase B: full ring test + score on the survivors ----
    const int nq = q_count;
    for (int i0 = 0; i0 < nq; i0 += kFastThreads) {
        const int i = i0 + tid;
        bool corner = false;
        int x = 0, y = 0;
        if (i < nq) {
            const uint32_t e = queue[i];
            x = e & 0xff; y = e >> 8;
            const uint8_t *p = tile + (y + 3) * kTP + (x + 3 + xoff);
            const int v = p[0];
            int d[16];
            d[0] = v - p[3 * kTP];      d[1] = v - p[3 * kTP + 1];  d[2] = v - p[2 * kTP + 2];  d[3] = v - p[kTP + 3];
            d[4] = v - p[3];            d[5] = v - p[-kTP + 3];     d[6] = v - p[-2 * kTP + 2]; d[7] = v - p[-3 * kTP + 1];
            d[8] = v - p[-3 * kTP];     d[9] = v - p[-3 * kTP - 1]; d[10] = v - p[-2 * kTP - 2]; d[11] = v - p[-kTP - 3];
            d[12] = v - p[-3];          d[13] = v - p[kTP - 3];     d[14] = v - p[2 * kTP - 2];  d[15] = v - p[3 * kTP - 1];
            uint32_t dark = 0, bright = 0;          // ring darker / brighter than the centre by > lowTh
#pragma unroll
            for (int k = 0; k < 16; ++k) { dark |= (uint32_t)(d[k] > lowTh) << k; bright |= (uint32_t)(d[k] < -lowTh) << k; }
            if (has_arc9(dark) || has_arc9(bright)) {
                corner = true;
                const int a = arc9_maxmin(d);
                int nd[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) nd[k] = -d[k];
                const int b = arc9_maxmin(nd);
                score[(y + 1) * kSP + (x + 1)] = (uint8_t)(max(a, b) - 1);   // = cornerScore, >= lowTh
            }
        }
        queue_push(corner, (uint16_t)((y << 8) | x), corners, &c_count);
    }
    __syncthreads();

    // ---- phase C: NMS at both thresholds, corners only ----
    const int nc = c_count;
    for (int i = tid; i < nc; i += kFastThreads) {
        const uint32_t e = corners[i];
        const int x = e & 0xff, y = e >> 8;
        const uint8_t *q = score + (y + 1) * kSP + (x + 1);
        const int s = q[0];
        int nmax_min = 0, nmax_ini = 0;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (dx == 0 && dy == 0) continue;
                const int n = q[dy * kSP + dx];
                nmax_min = max(nmax_min, n >= minTh ? n : 0);
                nmax_ini = max(nmax_ini, n >= iniTh ? n : 0);
            }
        const int idx = y * iw + x;
        if (s >= minTh && s > nmax_min) atomicOr(&bm_min[idx >> 5], 1u << (idx & 31));
        if (s >= iniTh && s > nmax_ini) atomicOr(&bm_ini[idx >> 5], 1u << (idx & 31));
    }
    __syncthreads();

    // ---- phase D: retry rule + ordered emission (row-major = ascending bit index) ----
    const int nbm = (iw * ih + 31) >> 5;
    int any = 0;
    for (int i = tid; i < nbm; i += kFastThreads) any |= (bm_ini[i] != 0);
    const int any_ini = __syncthreads_or(any);
    if (warp != 0) return;
    const uint32_t *bm = any_ini ? bm_ini : bm_min;
    uint32_t *slots = cell_slots + (size_t)f * g.slots_per_frame + L.slot_base + (size_t)c * L.cell_cap;
    int base = 0;
    for (int w0 = 0; w0 < nbm; w0 += 32) {
        const int wi = w0 + lane;
        uint32_t bits = wi < nbm ? bm[wi] : 0u;
        const int cnt = __popc(bits);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        int pos = base + inc - cnt;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int idx = wi * 32 + b;
            const int y = idx / iw, x = idx - y * iw;
            // keypoint in level coordinates relative to (16,16): cell pixel (x+3, y+3) + (j*wCell, i*hCell)
            slots[pos++] = pack_cand(x + 3 + cj * L.wCell, y + 3 + ci * L.hCell, score[(y + 1) * kSP + (x + 1)]);
        }
        base += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) *count_out = base;
}

void launch_fast(const Geo &g, const DevBuffers &b, int nframes, cudaStream_t s)
{
    int mw = 1, mh = 1;
    for (int l = 0; l < g.nlevels; ++l) if (g.lv[l].nCols > 0) { mw = mw > g.lv[l].wCell ? mw : g.lv[l].wCell; mh = mh > g.lv[l].hCell ? mh : g.lv[l].hCell; }
    FastSmem sm;
    sm.tp = (mw + 6 + 3 + 3) / 4 * 4; sm.sp = (mw + 2 + 3) / 4 * 4; sm.npix_max = (mw * mh + 1) / 2 * 2;
    const int tile_rows = mh + 6;
    const size_t bytes = (size_t)tile_rows * sm.tp + ((size_t)(tile_rows - 4) * sm.sp + 15) / 16 * 16 +
                         2 * (size_t)sm.npix_max * sizeof(uint16_t) + 2 * (size_t)((sm.npix_max + 31) / 32) * sizeof(uint32_t) + 16;
    dim3 grd(g.total_cells, nframes);
    k_fast_cells<<<grd, kFastThreads, bytes, s>>>(g, b.pyr, b.cell_count, b.cell_slots, sm, tile_rows);
}

} // namespace orbx

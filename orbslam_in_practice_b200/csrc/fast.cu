// fast.cu -- the per-cell FAST-9 loop of ComputeKeyPointsOctTree (ORBextractor.cpp:745-786).
//
// One WARP per (cell, frame); eight cells per block, each warp with a private shared-memory region, so
// there are no block barriers and no idle warps.  ncu history, round 1 (lane-instructions per pixel):
// thread-per-pixel 406 (ALU pipe 79 %), block-per-cell with work queues 224, warp-per-cell ~100, this ~50
// (3.76e8 warp instructions for 256 VGA frames = 243 M tested pixels).
// The warp stages the cell's (wCell+6) x (hCell+6) u8 tile with aligned 8-byte cp.async into a tile whose pitch is
// an ODD multiple of 8 bytes (56, 72, ...: 14, 18, ... words), so that 16 consecutive rows of one column fall into 16
// different banks.  ncu, round 2: with the 64-byte pitch of round 1 a column's rows alternate between TWO banks, the
// candidates of a phase-B chunk come column by column, and every ring load cost 4.8 wavefronts -- the kernel sat at
// 82 % of the shared-memory wavefront peak (46 % of the wavefronts in phase B), not at the issue limit.  It then runs the
// reference's two attempts literally -- cv::FAST(cell, iniThFAST) and, only if that returned nothing,
// cv::FAST(cell, minThFAST) (:766-773).  One attempt at threshold t:
//   A  every interior pixel, lane = column, verdicts collected in per-lane bitmasks: compass pre-test.
//      Every 9-arc of the 16-ring holds one pixel of each opposite pair, so a corner needs
//      min(max(N,S),max(E,W)) > v+t (bright) or max(min(N,S),min(E,W)) < v-t (dark).  Survivors of either
//      polarity are compacted into one queue.
//   B  queue, all lanes busy: cornerScore as a sliding-window min over the circular ring with 3-input
//      min/max, BOTH polarities in one pass: a ring pixel e is held as the 16-bit pair (e, 255 - e), so
//      one VIMNMX3.U16x2 takes the arc minimum of e (bright) in the low half and 255 - (arc maximum of e)
//      (dark) in the high half.  score = max(max_arcs(min_arc e) - v, v - min_arcs(max_arc e)) - 1, and the
//      pixel is a corner iff that maximum > t.  ncu, round 2: the two scalar passes (one per polarity, ~150
//      instructions per 32 candidates, nearly every chunk mixed) were 24 % of the kernel's instructions.
//   C  corners only: strict 3x3 non-max suppression *inside the cell* (non-corners and pixels outside the
//      cell read 0) -> bitmap
//   D  emit set bits in row-major order into the cell's slot array.
// The octree kernel concatenates cells in row-major cell order = the order of vToDistributeKeys.
#include "orbx_internal.cuh"

namespace orbx {

constexpr int kFastWarps = 4;
constexpr int kFastThreads = kFastWarps * 32;
constexpr int kFastLeftover = 4;             // cell columns past a multiple of 32 that run transposed (lane = row) in phase A
// per-warp shared-memory layout, sized at launch from the largest cell of the geometry
struct FastSmem { int tp, sp, tile_rows, npix_max, off_score, off_queue, per_warp, zero_vecs; };

__device__ __forceinline__ int min3(int a, int b, int c) { return __vimin3_s32(a, b, c); }
__device__ __forceinline__ int max3(int a, int b, int c) { return __vimax3_s32(a, b, c); }

// ring pixel e (0..255) as the 16-bit pair (e, 255 - e): e * (1 - 2^16) + (255 << 16), one IMAD
__device__ __forceinline__ uint32_t pack_pm(int e) { return (uint32_t)e * 0xffff0001u + 0x00ff0000u; }
__device__ __forceinline__ uint32_t min3x2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t max3x2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t min2x2(uint32_t a, uint32_t b) { uint32_t r; asm("min.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// cornerScore<16>'s core for both polarities at once on packed ring values (see pack_pm):
//   low half:  max over the 16 arcs of 9 contiguous ring values of their minimum           (bright arcs)
//   high half: max over the arcs of min(255 - e) = 255 - (min over the arcs of their maximum) (dark arcs)
__device__ __forceinline__ uint32_t arc9_both(const uint32_t (&e)[16])
{
    uint32_t m3[16], m9[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = min3x2(e[k], e[(k + 1) & 15], e[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m9[k] = min3x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
    const uint32_t b0 = max3x2(m9[0], m9[1], m9[2]), b1 = max3x2(m9[3], m9[4], m9[5]), b2 = max3x2(m9[6], m9[7], m9[8]);
    const uint32_t b3 = max3x2(m9[9], m9[10], m9[11]), b4 = max3x2(m9[12], m9[13], m9[14]);
    return max3x2(max3x2(b0, b1, b2), b3, max3x2(b4, m9[15], m9[15]));
}

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// TP: tile pitch as a compile-time constant (56 covers every cell up to 56 - 7 - 6 = 43 px wide incl. the alignment
// slack, i.e. all VGA / KITTI / 4K geometries), 0 = the runtime pitch sm.tp.
// kRange: the launch covers cells [g.fast_cell_lo, g.fast_cell_hi) (one level, low-latency path) instead of all of them.  A template
// parameter on purpose: with the range test in the one kernel, ptxas allocated 39 registers instead of 48 (8 bytes of spills) and
// the batch path lost 3.5 % (0.387 -> 0.401 ms per 256 VGA frames); the full-range instantiation is the round-2 kernel unchanged.
template <int TP, bool kRange>
__global__ void __launch_bounds__(kFastThreads, 8)
k_fast_cells(const __grid_constant__ Geo g, const uint8_t *__restrict__ pyr,
             int *__restrict__ cell_count, uint32_t *__restrict__ cell_slots, const int4 *__restrict__ cell_tab, const FastSmem sm)
{
    extern __shared__ __align__(16) unsigned char fast_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cell = (kRange ? g.fast_cell_lo : 0) + blockIdx.x * kFastWarps + warp, f = blockIdx.y + g.frame0;
    if (cell >= (kRange ? g.fast_cell_hi : g.total_cells)) return;
    unsigned char *mine = fast_smem + warp * sm.per_warp;
    uint8_t *tile = mine;                                                     // [tile_rows][tp]
    uint8_t *score = mine + sm.off_score;                                     // [tile_rows - 4][sp]
    uint16_t *queue = reinterpret_cast<uint16_t *>(mine + sm.off_queue);      // phase A survivors (x | y << 6), either polarity
    const int kTP = TP ? TP : sm.tp, kSP = sm.sp;

    // cell rectangle, ORBextractor.cpp:745-762, precomputed on the host (build_cell_table): one 16-byte load
    const int4 ct = __ldg(cell_tab + cell);
    const LevelGeom &L = g.lv[ct.x];
    int *count_out = cell_count + (size_t)f * g.total_cells + cell;
    const int iniX = ct.y & 0xffff, iniY = ct.y >> 16;
    const int cw = ct.z & 0xffff, ch = ct.z >> 16;     // cell image size handed to cv::FAST (0: the reference skips the cell)
    if (cw == 0) { if (lane == 0) *count_out = 0; return; }
    const int iw = cw - 6, ih = ch - 6;                // pixels FAST actually tests

    // ---- phase 0: stage the tile with 8-byte cp.async (no registers, every row of the lane in flight at once):
    //      tile column 0 is the 8-byte aligned pixel at or left of iniX (rows are 64-byte aligned, the interior
    //      starts at byte 32), so cell column c lives at tile column c + xoff.  4 rows per warp pass (2 for wide cells). ----
    const int xoff = iniX & 7;
    {
        const uint8_t *img = pyr + L.base + (size_t)f * L.frame_stride + (size_t)(kPadY + iniY) * L.pitch + kPadX + (iniX - xoff);
        const int vpr = (cw + xoff + 7) >> 3;               // 8-byte vectors per tile row (<= kTP / 8)
        const int lsh = vpr <= 8 ? 3 : 4;                   // lanes per row: 8 (4 rows per pass) or 16 (2 rows per pass)
        const int sub = lane & ((1 << lsh) - 1), rr = lane >> lsh, rstep = 32 >> lsh;
        if (sub < vpr) {
            const uint8_t *src = img + (size_t)rr * L.pitch + sub * 8;
            uint8_t *dst = tile + rr * kTP + sub * 8;
            const size_t sstep = (size_t)rstep * L.pitch;
            for (int r = rr; r < ch; r += rstep, src += sstep, dst += rstep * kTP) cp_async8(dst, src);
        }
    }
    const int minTh = g.min_th, iniTh = g.ini_th;
    uint32_t *slots = cell_slots + (size_t)f * g.slots_per_frame + (size_t)(unsigned)ct.w;
    const uint32_t lt_mask = (1u << lane) - 1u;

    // phase A for 32 columns x0 .. x0+31 and rows yb .. yb+rows-1: compass pre-test, lane = column, walking down the rows with
    // the column's last six pixels in registers (N of row y is the centre of row y+3 and S of row y+6): three loads per pixel.
    // The window holds the pixels PACKED as (e, 255 - e) (see pack_pm), so one 16x2 min/max works on both polarities: with
    // T = pack(v) + (t, t), max(min(max(N,S),max(E,W)), T) = min(max3(N,S,T), max3(E,W,T)) differs from T iff
    // min(max(N,S),max(E,W)) > v + t (low half, bright) or 255 - max(min(N,S),min(E,W)) > 255 - v + t (high half, dark).
    // Every half of Y is >= T's, so the 32-bit difference T - Y is negative iff either half differs: its sign is the verdict,
    // shifted into the lane's mask with one funnel shift.  Per pixel: 3 LDS.U8, 3 packs + T + difference on the FMA pipe (IMAD),
    // 2 VIMNMX3 + VIMNMX + SHF on the ALU pipe.  The scalar form (6 VIMNMX, 2 IADD3, LOP3, SHF = 10 ALU-pipe instructions of 13,
    // two cycles each) was ALU-pipe bound.  Returns the block TRANSPOSED: lane = row yb + lane, bit = column x0 + bit.
    auto column_block = [&](int x0, int yb, int rows, uint32_t tt) -> uint32_t {
        const int x = x0 + lane;
        const bool inx = x < iw;
        const int nr = (rows + 7) & ~7;                    // rows walked (the extra ones are masked off below)
        const uint8_t *p = tile + (yb + 3) * kTP + 3 + xoff + (inx ? x : 0);
        uint32_t c0 = pack_pm(p[-3 * kTP]), c1 = pack_pm(p[-2 * kTP]), c2 = pack_pm(p[-kTP]), c3 = pack_pm(p[0]), c4 = pack_pm(p[kTP]),
                 c5 = pack_pm(p[2 * kTP]);
        uint32_t m = 0;
#pragma unroll 1
        for (int y0 = 0; y0 < nr; y0 += 8, p += 8 * kTP) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t n = pack_pm(p[(k + 3) * kTP]), e = pack_pm(p[k * kTP + 3]), w = pack_pm(p[k * kTP - 3]);
                const uint32_t T = c3 + tt;
                const uint32_t Y = min2x2(max3x2(n, c0, T), max3x2(e, w, T));
                m = __funnelshift_l(T - Y, m, 1);
                c0 = c1; c1 = c2; c2 = c3; c3 = c4; c4 = c5; c5 = n;
            }
        }
        // row y of the block sits at bit nr-1-y: reverse to bit y, drop the rows past the cell and the columns past it
        m = __brev(m) >> (32 - nr);
        m = inx ? m & (0xffffffffu >> (32 - rows)) : 0u;
        // 32 x 32 bit transpose across the warp (lane c, bit r) -> (lane r, bit c): five butterfly stages, each swapping the
        // off-diagonal blocks of 2j x 2j tiles; the byte-granular stages are one PRMT, the others a rotate and a bitwise select
        {
            const uint32_t y16 = __shfl_xor_sync(0xffffffffu, m, 16);
            m = __byte_perm(m, y16, (lane & 16) ? 0x3276u : 0x5410u);
            const uint32_t y8 = __shfl_xor_sync(0xffffffffu, m, 8);
            m = __byte_perm(m, y8, (lane & 8) ? 0x3715u : 0x6240u);
#pragma unroll
            for (int j = 4; j >= 1; j >>= 1) {
                const uint32_t km = j == 4 ? 0x0f0f0f0fu : j == 2 ? 0x33333333u : 0x55555555u;
                const bool up = (lane & j) != 0;
                const uint32_t y = __shfl_xor_sync(0xffffffffu, m, j);
                const uint32_t ry = __funnelshift_l(y, y, up ? 32 - j : j);     // partner's block moved onto the positions this lane takes over
                const uint32_t keep = up ? ~km : km;
                m = (m & keep) | (ry & ~keep);
            }
        }
        return m;
    };

    // The reference calls cv::FAST(cell, iniThFAST) and, only if that returns nothing, cv::FAST(cell, minThFAST)
    // (:766-773).  Same here: the whole detect-score-NMS pipeline runs at iniThFAST first (fewer compass
    // survivors and corners than at minThFAST) and is repeated at minThFAST only for cells left empty.
    int base = 0;                                          // keypoints emitted
#pragma unroll 1
    for (int attempt = 0; attempt < 2 && base == 0; ++attempt) {
        const int th = attempt == 0 ? iniTh : minTh;
        const uint32_t tt = (uint32_t)th * 0x10001u;
        {
            uint4 *z = reinterpret_cast<uint4 *>(score);
#pragma unroll 1                                           // three trips; unrolled by eight with a remainder ladder it was 65 instructions
            for (int i = lane; i < sm.zero_vecs; i += 32) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        cp_async_wait_all();                               // the tile (first attempt; nothing pending on the second)
        __syncwarp();

        // ---- phase A: compass pre-test; the survivors of 32 rows go into the queue in ROW-MAJOR order (lane = row after the
        //      transpose, ascending columns inside a row), so the queue order is the order of the output ----
        int nq = 0;                                        // queue fill
        for (int yb = 0; yb < ih; yb += 32) {
            const int rows = min(32, ih - yb);
            uint32_t rlo = 0u, rhi = 0u;
            const bool leftover = iw > 32 && iw - 32 <= kFastLeftover;
            const int xfull = leftover ? 32 : iw;          // columns that run as lane-per-column blocks
#pragma unroll 1
            for (int x0 = 0; x0 < xfull; x0 += 32) {       // one body for both blocks (inlined twice it cost 16 registers)
                const uint32_t m = column_block(x0, yb, rows, tt);
                if (x0 == 0) rlo = m; else rhi = m;
            }
            if (leftover) {
                // a few columns past a multiple of 32 (VGA levels 2 and 5 have 33-px cells): a lane-per-column pass would walk
                // every row for one or two active lanes, so these columns run lane = row, five loads per pixel
                const int y = yb + lane;
                const bool iny = lane < rows;
                const uint8_t *p = tile + ((iny ? y : 0) + 3) * kTP + 3 + xoff + 32;
                for (int x = 32; x < iw; ++x, ++p) {
                    const int v = p[0], n = p[-3 * kTP], s2 = p[3 * kTP], e = p[3], w = p[-3];
                    const int hi = min(max(n, s2), max(e, w)), lo = max(min(n, s2), min(e, w));
                    if (iny && (hi > v + th || lo < v - th)) rhi |= 1u << (x - 32);
                }
            }
            const int cnt = __popc(rlo) + __popc(rhi);
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            // queue entry: x | y << 6   (x, y < 64)
            const int ebase = (yb + lane) << 6;
            uint16_t *q = queue + nq + (inc - cnt);
            while (rlo) {
                const int b = __ffs(rlo) - 1;
                rlo &= rlo - 1;
                *q++ = (uint16_t)(ebase + b);
            }
            while (rhi) {
                const int b = __ffs(rhi) - 1;
                rhi &= rhi - 1;
                *q++ = (uint16_t)(ebase + 32 + b);
            }
            nq += __shfl_sync(0xffffffffu, inc, 31);
        }
        __syncwarp();

        // ---- phase B: exact score of both polarities in one packed pass (see arc9_both): bright = max over arcs of
        //      the arc's minimum - v, dark = v - min over arcs of the arc's maximum (cornerScore's two halves). ----
        for (int i0 = 0; i0 < nq; i0 += 32) {
            const int i = i0 + lane;
            if (i < nq) {
                const uint32_t ent = queue[i];
                const int x = ent & 63, y = ent >> 6;
                const uint8_t *p = tile + (y + 3) * kTP + (x + 3 + xoff);
                const int v = p[0];
                uint32_t e[16];
                e[0] = pack_pm(p[3 * kTP]);      e[1] = pack_pm(p[3 * kTP + 1]);   e[2] = pack_pm(p[2 * kTP + 2]);    e[3] = pack_pm(p[kTP + 3]);
                e[4] = pack_pm(p[3]);            e[5] = pack_pm(p[-kTP + 3]);      e[6] = pack_pm(p[-2 * kTP + 2]);   e[7] = pack_pm(p[-3 * kTP + 1]);
                e[8] = pack_pm(p[-3 * kTP]);     e[9] = pack_pm(p[-3 * kTP - 1]);  e[10] = pack_pm(p[-2 * kTP - 2]);  e[11] = pack_pm(p[-kTP - 3]);
                e[12] = pack_pm(p[-3]);          e[13] = pack_pm(p[kTP - 3]);      e[14] = pack_pm(p[2 * kTP - 2]);   e[15] = pack_pm(p[3 * kTP - 1]);
                const uint32_t both = arc9_both(e);
                // low half: max_arcs(min_arc e);  high half: 255 - min_arcs(max_arc e)
                const int best = max((int)(both & 0xffffu) - v, (int)(both >> 16) + v - 255);
                if (best > th)                            // corner at th; cornerScore = best - 1 >= th
                    score[(y + 1) * kSP + (x + 1)] = (uint8_t)(best - 1);
            }
        }
        __syncwarp();

        // ---- phase C: strict 3x3 NMS inside the cell (non-corners and pixels outside the cell read 0) over the same queue, and the
        //      emission in one: the queue is row-major, so a ballot-ranked store writes the survivors in output order ----
        for (int i0 = 0; i0 < nq; i0 += 32) {
            const int i = i0 + lane;
            bool keep = false;
            uint32_t cand = 0;
            if (i < nq) {
                const uint32_t ent = queue[i];
                const int x = ent & 63, y = ent >> 6;
                const uint8_t *q = score + (y + 1) * kSP + (x + 1);
                const int s = q[0];
                if (s != 0) {                             // every stored score is >= th >= 1
                    const int nmax = max3(max3((int)q[-kSP - 1], (int)q[-kSP], (int)q[-kSP + 1]), max3((int)q[-1], (int)q[1], (int)q[kSP - 1]),
                                          max((int)q[kSP], (int)q[kSP + 1]));
                    keep = s > nmax;
                    // keypoint in level coordinates relative to (16,16): cell pixel (x+3, y+3) + (j*wCell, i*hCell)
                    cand = pack_cand(x + 3 + iniX - kMinBorder, y + 3 + iniY - kMinBorder, s);
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, keep);
            if (keep) slots[base + __popc(bal & lt_mask)] = cand;
            base += __popc(bal);
        }
    }
    if (lane == 0) *count_out = base;
}

void launch_fast(const Geo &g, const DevBuffers &b, int nframes, cudaStream_t s, int level_lo, int level_hi)
{
    // cells of levels [level_lo, level_hi); the default covers the whole pyramid in one launch
    if (level_hi > g.nlevels) level_hi = g.nlevels;
    const int cell_lo = g.lv[level_lo].cell_base;
    const int cell_hi = level_hi >= g.nlevels ? g.total_cells : g.lv[level_hi].cell_base;
    if (cell_hi <= cell_lo) return;
    int mw = 1, mh = 1;
    for (int l = 0; l < g.nlevels; ++l) if (g.lv[l].nCols > 0) { mw = mw > g.lv[l].wCell ? mw : g.lv[l].wCell; mh = mh > g.lv[l].hCell ? mh : g.lv[l].hCell; }
    auto up16 = [](int v) { return (v + 15) / 16 * 16; };
    FastSmem sm;
    sm.tp = (mw + 6 + 7 + 7) / 8 * 8;                     // cell + 7 bytes of alignment slack, in 8-byte vectors ...
    if ((sm.tp & 15) == 0) sm.tp += 8;                   // ... and an odd number of them (bank spread, see the file header)
    if (sm.tp < 56) sm.tp = 56;
    sm.sp = (mw + 2 + 3) / 4 * 4; sm.tile_rows = mh + 6;
    sm.npix_max = (mw * mh + 1) / 2 * 2;
    sm.off_score = up16(sm.tile_rows * sm.tp);
    sm.off_queue = sm.off_score + up16((mh + 2) * sm.sp);
    sm.per_warp = sm.off_queue + up16(sm.npix_max * 2);
    sm.zero_vecs = (sm.off_queue - sm.off_score) / 16;
    // phase A walks its rows in groups of eight: the rows past the cell are masked off, but they are read, so the
    // warp's region must reach (mh rounded up to 8) + 6 rows plus the column slack of one more row
    const int walk = ((mh + 7) / 8 * 8 + 7) * sm.tp;
    if (sm.per_warp < walk) sm.per_warp = up16(walk);
    const size_t bytes = (size_t)sm.per_warp * kFastWarps;
    dim3 grd((cell_hi - cell_lo + kFastWarps - 1) / kFastWarps, nframes);
    Geo gl = g;
    gl.fast_cell_lo = cell_lo; gl.fast_cell_hi = cell_hi;
    // per-device function attribute; a handful of nanoseconds, so no process-wide caching (one handle per device each)
    const bool range = cell_lo != 0 || cell_hi != g.total_cells;
#define FAST_LAUNCH(TPV, RV) do { \
        if (bytes > 48 * 1024) cudaFuncSetAttribute(k_fast_cells<TPV, RV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); \
        k_fast_cells<TPV, RV><<<grd, kFastThreads, bytes, s>>>(gl, b.pyr, b.cell_count, b.cell_slots, b.cell_tab, sm); } while (0)
    if (sm.tp == 56) { if (range) FAST_LAUNCH(56, true); else FAST_LAUNCH(56, false); }
    else { if (range) FAST_LAUNCH(0, true); else FAST_LAUNCH(0, false); }
#undef FAST_LAUNCH
}

} // namespace orbx

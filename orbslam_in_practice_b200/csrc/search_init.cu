// search_init.cu -- ORBmatcher::SearchForInitialization (src/ORBmatcher.cpp:9-126) with the Frame grid it
// queries (Frame::AssignFeaturesToGrid / GetGridId src/Frame.cpp:144-168, GetFeaturesInArea :219-271).
//
// One thread block per frame pair (F1 = reference frame, F2 = current frame):
//   1  grid of F2: keys (cell << 16 | keypoint index) of the octave-0 keypoints, bitonic-sorted in shared memory.
//      Cell id = ix * 48 + iy, so the reference's enumeration order (ix outer, iy inner, push_back order inside a
//      cell) is exactly ascending key order, and for a fixed ix the cells iy0..iy1 are ONE contiguous range.
//   2a parallel, warp per octave-0 query: window query -> candidate list (index, Hamming distance) in reference
//      scan order, written to the caller's workspace.  This is the data-parallel part (all the POPC work).
//   2b sequential, one warp: the reference's order-dependent part -- the gate `vMatchedDistance[i2] <= dist`
//      (:49-50) makes every query depend on the acceptances before it -- replayed over the compact lists:
//      best-2 with first-minimum-wins, TH_LOW / ratio acceptance, one-to-one displacement (:65-77), rotation
//      histogram pushes (:79-90).
//   3  ComputeThreeMaxima (:147-188) and the histogram filter (:93-118), then the prev-matched update (:121-124).
#include "orbx_internal.cuh"

#include <climits>

namespace orbx {

constexpr int kGridRows = 48, kGridCols = 64;   // Frame.h:11-12
constexpr int kHistoLength = 30;                // ORBmatcher.cpp:6
constexpr int kSiThreads = 512;
constexpr uint32_t kInfKey = 0xffffffffu;

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4 *__restrict__ b)
{
    const uint4 b0 = __ldg(b), b1 = __ldg(b + 1);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// block-wide bitonic sort of n (a power of two) keys in shared memory; ends with a barrier
__device__ __forceinline__ void block_bitonic_sort(uint32_t *keys, int n, int tid)
{
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n; i += kSiThreads) {
                const int p = i ^ j;
                if (p > i) {
                    const uint32_t a = keys[i], b = keys[p];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[p] = a; }
                }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(kSiThreads)
k_search_init(const __grid_constant__ SearchInitArgs A)
{
    extern __shared__ __align__(16) unsigned char si_smem[];
    uint32_t *keys = reinterpret_cast<uint32_t *>(si_smem);                         // [sort_n] sorted grid keys of F2
    unsigned short *cellstart = reinterpret_cast<unsigned short *>(keys + A.sort_n);// [64 * 49 + 1] first key of (ix, iy)
    int *matchedDist = reinterpret_cast<int *>(cellstart + ((kGridCols * (kGridRows + 1) + 2) & ~1));   // [cap]
    int *m21 = matchedDist + A.cap;                                                  // [cap]
    unsigned short *qcnt = reinterpret_cast<unsigned short *>(m21 + A.cap);          // [cap] candidates per query of F1
    unsigned short *active = qcnt + A.cap;                                           // [cap] queries with candidates, ascending
    float *ang2 = reinterpret_cast<float *>(active + A.cap);                         // [cap] angles of F2's keypoints (read by the replay)
    __shared__ int hist[kHistoLength];
    __shared__ int s_nvalid, s_nmatches, s_keep[3];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pair = blockIdx.x;
    const int fa = A.pair_a[pair], fb = A.pair_b[pair];
    const int n1 = min(max(A.counts[fa], 0), A.cap), n2 = min(max(A.counts[fb], 0), A.cap);   // never trust a count beyond the row capacity
    const orbx_keypoint *kp1 = A.kps + (size_t)fa * A.cap, *kp2 = A.kps + (size_t)fb * A.cap;
    const uint4 *d1 = reinterpret_cast<const uint4 *>(A.desc + (size_t)fa * A.cap * 32);
    const uint4 *d2 = reinterpret_cast<const uint4 *>(A.desc + (size_t)fb * A.cap * 32);
    float *prev = A.prev_matched + (size_t)pair * A.cap * 2;
    int *m12 = A.matches12 + (size_t)pair * A.cap;
    uint32_t *ws = A.workspace + (size_t)pair * A.ws_words_per_pair;
    // group mode (SearchByBoW-style): candidates of a query are the F2 keypoints of the same group, no window
    const unsigned short *g1 = A.groups ? A.groups + (size_t)fa * A.cap : nullptr, *g2 = A.groups ? A.groups + (size_t)fb * A.cap : nullptr;

    // image bounds for zero distortion (Frame.cpp:113-118) and grid cell sizes (:59-60)
    const orbm_window_params &W = A.w;
    const float minX = W.use_bounds ? W.min_x : 0.f, maxX = W.use_bounds ? W.max_x : (float)W.width;
    const float minY = W.use_bounds ? W.min_y : 0.f, maxY = W.use_bounds ? W.max_y : (float)W.height;
    const float wInv = (float)kGridCols / (maxX - minX), hInv = (float)kGridRows / (maxY - minY);

    // ---- 1: grid keys of F2; only the octaves some query can ask for (SearchForInitialization: octave 0 alone,
    //      it queries minLevel = maxLevel = 0) ----
    // Keys are appended in any order (they are unique, so the sorted result does not depend on it), then only the next
    // power of two above their number is sorted: SearchForInitialization keeps ~1/5 of the keypoints (octave 0).
    if (tid == 0) { s_nvalid = 0; s_nmatches = 0; }
    if (tid < kHistoLength) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n2; i += kSiThreads) {
        const orbx_keypoint k = kp2[i];
        ang2[i] = k.angle;
        // GetGridId takes doubles (Frame.cpp:161-168); std::round = half away from zero
        const int ix = (int)round(((double)k.x - (double)minX) * (double)wInv);
        const int iy = (int)round(((double)k.y - (double)(W.literal_gridid_bug ? maxY : minY)) * (double)hInv);
        if (g2) {
            const uint32_t g = g2[i];
            if (g != 0xffffu) keys[atomicAdd(&s_nvalid, 1)] = (g << 16) | (uint32_t)i;
        } else if (k.octave >= A.grid_level_min && k.octave <= A.grid_level_max && ix >= 0 && ix < kGridCols && iy >= 0 && iy < kGridRows)
            keys[atomicAdd(&s_nvalid, 1)] = ((uint32_t)(ix * kGridRows + iy) << 16) | (uint32_t)i;
    }
    for (int i = tid; i < A.cap; i += kSiThreads) { matchedDist[i] = INT_MAX; m21[i] = -1; m12[i] = -1; }
    __syncthreads();
    int sort_m = 32;
    while (sort_m < s_nvalid) sort_m <<= 1;                                          // <= A.sort_n
    for (int i = s_nvalid + tid; i < sort_m; i += kSiThreads) keys[i] = kInfKey;
    __syncthreads();
    block_bitonic_sort(keys, sort_m, tid);
    // first key position of every (ix, iy) and the end sentinel: lower_bound by binary search
    for (int c = tid; c <= kGridCols * kGridRows && !g2; c += kSiThreads) {
        const uint32_t target = (uint32_t)c << 16;
        int lo = 0, hi = sort_m;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < target) lo = mid + 1; else hi = mid; }
        cellstart[c] = (unsigned short)lo;                                           // the sentinel entry equals s_nvalid
    }
    __syncthreads();
    const int nvalid = s_nvalid;                        // keypoints of F2 that landed in the grid
    // workspace rows: one candidate list per query of F1, stride = nvalid entries (+1 word for the count)
    const unsigned long long stride = (unsigned long long)nvalid + 1ull;
    const bool ws_ok = stride * (unsigned long long)n1 <= A.ws_words_per_pair;
    if (!ws_ok) { if (tid == 0) A.nmatches[pair] = -1; return; }          // caller's workspace too small

    // ---- 2a: candidate lists, warp per query ----
    for (int q = warp; q < n1; q += kSiThreads / 32) {
        uint32_t *list = ws + (unsigned long long)q * stride;
        const orbx_keypoint kq = kp1[q];
        int cnt = 0;
        if (g1) {
            const uint32_t g = g1[q];
            if (g != 0xffffu) {
                // the group's keys are one contiguous range of the sorted array, in F2 index order (= the node's feature list)
                int lo = 0, hi = sort_m;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (keys[mid] < (g << 16)) lo = mid + 1; else hi = mid; }
                int e = lo; hi = sort_m;
                while (e < hi) { const int mid = (e + hi) >> 1; if (keys[mid] < ((g + 1u) << 16)) e = mid + 1; else hi = mid; }
                const uint4 a0 = __ldg(d1 + 2 * q), a1 = __ldg(d1 + 2 * q + 1);
                for (int p = lo + lane; p < e; p += 32) {
                    const int i2 = (int)(keys[p] & 0xffffu);
                    list[1 + p - lo] = (uint32_t)i2 | ((uint32_t)hamming256(a0, a1, d2 + 2 * i2) << 16);
                }
                cnt = e - lo;
            }
            if (lane == 0) { list[0] = (uint32_t)cnt; qcnt[q] = (unsigned short)cnt; }
            continue;
        }
        const float x = prev[2 * q], y = prev[2 * q + 1];
        // :25-27 level1 > 0 -> continue (query_level range 0..0 there); a NaN centre marks a query without a projection
        if (kq.octave >= W.query_level_min && kq.octave <= W.query_level_max && x == x) {
            const float r = __fmul_rn(W.radius, W.level_scale[kq.octave & 15]);
            // GetFeaturesInArea's level filter, Frame.cpp:245-258: octave < minLevel or (maxLevel >= 0 and octave > maxLevel) -> skip
            const int minLevel = W.level_below < 0 ? 0 : kq.octave - W.level_below;
            const int maxLevel = W.level_above < 0 ? -1 : kq.octave + W.level_above;
            // GetFeaturesInArea cell window, Frame.cpp:225-239 (float math)
            const int cx0 = max(0, (int)floorf((x - minX - r) * wInv));
            const int cx1 = min(kGridCols - 1, (int)ceilf((x - minX + r) * wInv));
            const int cy0 = max(0, (int)floorf((y - minY - r) * hInv));
            const int cy1 = min(kGridRows - 1, (int)ceilf((y - minY + r) * hInv));
            if (cx0 < kGridCols && cx1 >= 0 && cy0 < kGridRows && cy1 >= 0) {
                const uint4 a0 = __ldg(d1 + 2 * q), a1 = __ldg(d1 + 2 * q + 1);
                // one 32-key chunk of the sorted keys [b0, e); keys are visited in ascending order = the reference's
                // enumeration (ix outer, iy inner, insertion order inside a cell); cells outside rows cy0..cy1 are skipped
                auto chunk = [&](int b0, int e) {
                    const int p = b0 + lane;
                    bool ok = false; uint32_t ent = 0;
                    if (p < e) {
                        const uint32_t key = keys[p];
                        const int i2 = (int)(key & 0xffffu), cell = (int)(key >> 16);
                        const int iy = cell - (cell / kGridRows) * kGridRows;
                        if (iy >= cy0 && iy <= cy1) {
                            const orbx_keypoint k2 = kp2[i2];
                            const float dx = k2.x - x, dy = k2.y - y;                  // Frame.cpp:263-266
                            const bool lvl = k2.octave >= minLevel && (maxLevel < 0 || k2.octave <= maxLevel);
                            if (lvl && fabsf(dx) < r && fabsf(dy) < r) {
                                ok = true;
                                ent = (uint32_t)i2 | ((uint32_t)hamming256(a0, a1, d2 + 2 * i2) << 16);
                            }
                        }
                    }
                    const uint32_t bal = __ballot_sync(0xffffffffu, ok);
                    if (ok) list[1 + cnt + __popc(bal & ((1u << lane) - 1u))] = ent;
                    cnt += __popc(bal);
                };
                // Columns cx0..cx1 are one contiguous key range.  Walking it whole (and skipping the rows outside the
                // window) needs fewer 32-key chunks than one pass per column whenever the band holds <= 32 keys per
                // column on average -- the common case (SearchForInitialization: ~3 keys per column pass otherwise).
                const int ks = cellstart[cx0 * kGridRows], ke = cellstart[cx1 * kGridRows + kGridRows];
                if (ke - ks <= 32 * (cx1 - cx0 + 1)) {
                    for (int b0 = ks; b0 < ke; b0 += 32) chunk(b0, ke);
                } else {
                    for (int ix = cx0; ix <= cx1; ++ix) {
                        const int s_ = cellstart[ix * kGridRows + cy0], e_ = cellstart[ix * kGridRows + cy1 + 1];
                        for (int b0 = s_; b0 < e_; b0 += 32) chunk(b0, e_);
                    }
                }
            }
        }
        if (lane == 0) { list[0] = (uint32_t)cnt; qcnt[q] = (unsigned short)cnt; }
    }
    __syncthreads();

    // group mode replays the queries node by node (ascending group, then ascending index -- the order in which upstream's
    // SearchByBoW walks the two feature vectors): sort (group << 16 | q) of the queries that have candidates
    if (g1) {
        if (tid == 0) s_nvalid = 0;
        __syncthreads();
        for (int q = tid; q < n1; q += kSiThreads)
            if (qcnt[q] != 0) keys[atomicAdd(&s_nvalid, 1)] = ((uint32_t)g1[q] << 16) | (uint32_t)q;
        __syncthreads();
        int m = 32;
        while (m < s_nvalid) m <<= 1;
        for (int i = s_nvalid + tid; i < m; i += kSiThreads) keys[i] = kInfKey;
        __syncthreads();
        block_bitonic_sort(keys, m, tid);
        for (int i = tid; i < s_nvalid; i += kSiThreads) active[i] = (unsigned short)(keys[i] & 0xffffu);
        __syncthreads();
    }

    // ---- 2b: the order-dependent replay, one warp ----
    if (warp == 0) {
        const float factor = kHistoLength / 360.0f;                          // :17
        int nmatches = 0;
        // Only queries with candidates take part (:32-33 skips the others, which also covers wrong octaves); their
        // indices are compacted in ascending order first.  The lists live in global memory: a query's first 64 entries
        // are fetched one iteration ahead, so the serial chain never waits on a load (ncu: the replay used to be ~70 %
        // of the kernel's time, spent on dependent global loads by a single warp).
        int na = g1 ? s_nvalid : 0;
        for (int b0 = 0; b0 < n1 && !g1; b0 += 32) {
            const int q = b0 + lane;
            const bool on = q < n1 && qcnt[q] != 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, on);
            if (on) active[na + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)q;
            na += __popc(bal);
        }
        __syncwarp();
        int qn = 0, cn = 0;
        const uint32_t *ln = ws;
        uint32_t n0 = 0, n1e = 0;
        float an = 0.f;                                                        // angle of the next query's keypoint
        if (na > 0) {
            qn = active[0]; cn = qcnt[qn]; ln = ws + (unsigned long long)qn * stride;
            n0 = lane < cn ? ln[1 + lane] : 0u; n1e = 32 + lane < cn ? ln[33 + lane] : 0u;
            an = kp1[qn].angle;
        }
        for (int ai = 0; ai < na; ++ai) {
            const int q = qn, cnt = cn;
            const uint32_t *list = ln;
            const uint32_t e0 = n0, e1 = n1e;
            const float angle1 = an;
            if (ai + 1 < na) {
                qn = active[ai + 1]; cn = qcnt[qn]; ln = ws + (unsigned long long)qn * stride;
                n0 = lane < cn ? ln[1 + lane] : 0u; n1e = 32 + lane < cn ? ln[33 + lane] : 0u;
                an = kp1[qn].angle;
            }
            uint32_t k1 = kInfKey, k2 = kInfKey;                              // two smallest (dist << 16 | scan position)
            for (int b0 = 0; b0 < cnt; b0 += 32) {
                const int p = b0 + lane;
                uint32_t key = kInfKey;
                if (p < cnt) {
                    const uint32_t ent = b0 == 0 ? e0 : (b0 == 32 ? e1 : list[1 + p]);
                    const int i2 = (int)(ent & 0xffffu), dist = (int)(ent >> 16);
                    // gate 0: :49-50 (matched at a distance <= dist);  gate 1: already taken by an earlier query
                    const bool skip = W.gate == 0 ? matchedDist[i2] <= dist : m21[i2] >= 0;
                    if (!skip) key = ((uint32_t)dist << 16) | (uint32_t)p;
                }
                const uint32_t hi = max(k1, key);
                k1 = min(k1, key); k2 = min(k2, hi);
            }
            // merge the lanes' sorted pairs with two warp reductions (REDUX): the smallest key, then the smallest of what
            // is left (keys are unique -- they contain the scan position -- except for the "none" value)
            {
                const uint32_t best = __reduce_min_sync(0xffffffffu, k1);
                k2 = __reduce_min_sync(0xffffffffu, k1 == best ? k2 : k1);
                k1 = best;
            }
            // the winner's list entry: from the prefetched registers when it is among the first 64
            const int wp = (int)(k1 & 0xffffu);
            const uint32_t w0 = __shfl_sync(0xffffffffu, e0, wp & 31), w1 = __shfl_sync(0xffffffffu, e1, wp & 31);
            if (lane == 0 && k1 != kInfKey) {
                const int bestDist = (int)(k1 >> 16);
                const int bestDist2 = k2 == kInfKey ? A.second_init : (int)(k2 >> 16);
                const int bestIdx2 = (int)((wp < 32 ? w0 : (wp < 64 ? w1 : list[1 + wp])) & 0xffffu);
                const bool ratio_ok = W.nnratio <= 0.f || (float)bestDist < __fmul_rn((float)bestDist2, W.nnratio);
                if (bestDist <= W.th_dist && ratio_ok) {                                // :65-67
                    if (m21[bestIdx2] >= 0) { m12[m21[bestIdx2]] = -1; nmatches--; }
                    m12[q] = bestIdx2; m21[bestIdx2] = q; matchedDist[bestIdx2] = bestDist; nmatches++;
                    if (W.check_orientation) {                                // :79-90
                        float rot = __fsub_rn(angle1, ang2[bestIdx2]);
                        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                        int bin = (int)roundf(__fmul_rn(rot, factor));
                        if (bin == kHistoLength) bin = 0;
                        hist[bin]++;
                        ws[(unsigned long long)q * stride] = 0x80000000u | (uint32_t)bin;   // remember the bin (count no longer needed)
                    }
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            s_nmatches = nmatches;
            // ComputeThreeMaxima :147-188
            int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
            for (int i = 0; i < kHistoLength; ++i) {
                const int s = hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < 0.1f * (float)max1) { ind3 = -1; }
            s_keep[0] = ind1; s_keep[1] = ind2; s_keep[2] = ind3;
        }
    }
    __syncthreads();

    // ---- 3: histogram filter (:93-118) and prev-matched update (:121-124) ----
    int removed = 0;
    for (int q = tid; q < n1; q += kSiThreads) {
        int m = m12[q];
        if (m >= 0 && W.check_orientation) {
            const uint32_t tag = ws[(unsigned long long)q * stride];
            const int bin = (int)(tag & 0xffu);
            if ((tag & 0x80000000u) && bin != s_keep[0] && bin != s_keep[1] && bin != s_keep[2]) { m12[q] = -1; m = -1; ++removed; }
        }
        if (m >= 0 && W.update_centers) { prev[2 * q] = kp2[m].x; prev[2 * q + 1] = kp2[m].y; }
    }
    if (removed) atomicSub(&s_nmatches, removed);
    __syncthreads();
    if (tid == 0) A.nmatches[pair] = s_nmatches;
}

size_t search_init_smem_bytes(int cap, int sort_n)
{
    return (size_t)sort_n * 4 + (size_t)((kGridCols * (kGridRows + 1) + 2) & ~1) * 2 + (size_t)cap * 16 + 16;
}

int launch_search_init(const SearchInitArgs &a, cudaStream_t s)
{
    const size_t smem = search_init_smem_bytes(a.cap, a.sort_n);
    if (smem > 227 * 1024) return -2;                     // per-pair tables do not fit one SM's shared memory (capacity beyond ~11 600)
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_search_init, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    k_search_init<<<a.npairs, kSiThreads, smem, s>>>(a);
    return 0;
}

} // namespace orbx

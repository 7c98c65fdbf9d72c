// octree.cu -- DistributeOctTree (ORBextractor.cpp:489-718) as a deterministic block-parallel
// algorithm: one thread block per (level, frame) problem.
//
// Parallel restatement.  Keys (packed candidates) live in a ping-pong array
// where every node owns a contiguous range whose internal order is the candidate order, so a
// DivideNode is a *stable 4-way partition inside the parent's range*.  The std::list is a node
// table in list order.  One pass =
//   A  exclusive scan over all keys of the one-hot quadrant vector (stable ranks): every warp streams a
//      contiguous segment, ranks come from ballots + running warp-uniform counters (two block barriers per
//      sweep); per key only the rank in its own quadrant (u32) and the quadrant (u8) are stored, the full
//      counter vectors only at the first and last key of every node
//   B  per node: child sizes = counters behind the last key - counters in front of the first key
//   C  choose the nodes to split and their processing rank
//        phase 1 (:558-617): every >1-key node, in list order
//        phase 2 (:629-691): >1-key nodes sorted by (size desc, later-created first = smaller
//                            list index first); cut at the first prefix with size >= N (:684)
//   D  scans over the processing order -> creation index of every child; push_front means the
//      new list is [children, latest created first] ++ [untouched nodes in old order]
//   E  write the new node table;  F  move the keys.
// The result (set AND order) equals the sequential reference under the defined tie-break
// "equal sizes: later-created node first" (= the reference under a monotonic allocator).
#include "orbx_internal.cuh"

#include <cstdlib>

namespace orbx {

constexpr int kOctMaxThreads = 1024;          // block size is chosen per geometry: 256 (VGA-sized problems) or 1024 (4K)
constexpr int kOctMaxWarps = kOctMaxThreads / 32;
constexpr int kOctKeyBytes = 4 + 4 + 4 + 2 + 2 + 1;   // per key in shared memory: two key arrays, own-quadrant rank, two node arrays, quadrant
int octree_smem_keys(const Geo &g);

struct __align__(16) Node {
    short x0, y0, x1, y1;
    int begin, count;
};

struct OctShared {
    int warp_i[kOctMaxWarps + 1];
    uint4 warp_v[kOctMaxWarps + 1];
    int n, nsplit, C, U, nToExpand, J, pending;
    int psize[8], pmulti[8];                              // closed-form phase 1: nodes / nodes with more than one key per depth
};

__device__ __forceinline__ int warp_incl_scan(int v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    return v;
}

// exclusive scan of one int per thread over the block; returns exclusive prefix, total via ref
template <int NW>
__device__ __forceinline__ int block_excl_scan(int v, OctShared &S, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int inc = warp_incl_scan(v);
    if (lane == 31) S.warp_i[warp] = inc;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) { const int t = S.warp_i[w]; if (w < warp) off += t; tot += t; }
    __syncthreads();
    total = tot;
    return off + inc - v;
}

__device__ __forceinline__ uint4 add4(uint4 a, uint4 b) { return make_uint4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

__device__ __forceinline__ int quadrant(uint32_t key, const Node &nd)
{
    // DivideNode :433-434, :462-476
    const int sx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1), sy = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
    const bool left = cand_x(key) < sx, up = cand_y(key) < sy;
    return left ? (up ? 0 : 2) : (up ? 1 : 3);
}

// A key as a 64-bit value whose maximum over a node is the key the reference keeps (:697-715): the highest score, and among equal
// scores the FIRST in candidate order -- candidates come cell by cell (row-major cells, ORBextractor.cpp:745-786) and row-major
// inside a cell, so the order is (cell row, cell column, y, x); (v - 3) / cell size is a multiply by a host-made reciprocal.
__device__ __forceinline__ unsigned long long best_key(uint32_t key, const LevelGeom &L)
{
    const int x = cand_x(key), y = cand_y(key);
    const unsigned long long i = (unsigned)(((y - 3) * L.inv_hCell) >> 18), j = (unsigned)(((x - 3) * L.inv_wCell) >> 18);
    const unsigned long long ord = (i << 32) | (j << 24) | ((unsigned long long)y << 12) | (unsigned long long)x;
    return ((unsigned long long)cand_score(key) << 40) | (0xffffffffffull - ord);
}
__device__ __forceinline__ uint32_t best_key_decode(unsigned long long v)
{
    const unsigned long long ord = 0xffffffffffull - (v & 0xffffffffffull);
    return pack_cand((int)(ord & 0xfff), (int)((ord >> 12) & 0xfff), (int)(v >> 40));
}

__device__ __forceinline__ uint32_t comp(const uint4 &v, int q) { return q == 0 ? v.x : q == 1 ? v.y : q == 2 ? v.z : v.w; }

// kU: key groups a warp loads before it processes them.  Problems whose keys live in global memory (4K-sized levels: tens of
// thousands of candidates, one block per SM because of the node tables) are bound by the latency of the dependent L2 round
// trips of every sweep (ncu, round 2: st_long 9.4 per issue at 39 % issue-active); kU = 4 puts four groups' loads in flight.
// kGlobalTables: the node tables live in a global scratch slice instead of shared memory (level quotas beyond one SM); a
// template parameter, not a runtime pointer choice, so that the normal instantiations keep shared-memory addressing (as a
// runtime select every node-table access became a generic load: VGA octree 0.140 -> 0.161 ms).
template <int kOctThreads, int kU, bool kGlobalTables = false>
__global__ void __launch_bounds__(kOctThreads, kOctThreads == 256 ? 5 : ((kOctThreads == 512 && kU == 1) ? 2 : 1))   // latency-bound: favour resident blocks over registers
k_octree(const __grid_constant__ Geo g, const int *__restrict__ cell_count, const uint32_t *__restrict__ cell_slots,
         uint32_t *keysA_all, uint32_t *keysB_all, uint16_t *nodeA_all, uint16_t *nodeB_all, uint4 *scanE_all,
         int *__restrict__ ncand_out, uint32_t *__restrict__ kept_out, int *__restrict__ nkept_out, const int smem_keys, const int level_lo,
         unsigned char *node_scratch, const unsigned long long node_scratch_stride)
{
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    // node tables: shared memory, or -- for a level quota whose tables do not fit one SM (more than ~2 400 features on one
    // level) -- this problem's slice of a global scratch buffer (slow, but the configuration is served instead of refused)
    unsigned char *smem_raw = kGlobalTables ? node_scratch + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * node_scratch_stride : smem_dyn;
    __shared__ OctShared S;

    const int level = level_lo + blockIdx.x, f = blockIdx.y + g.frame0, tid = threadIdx.x;
    constexpr int kOctWarps = kOctThreads / 32;
    const LevelGeom &L = g.lv[level];
    const int NC = L.node_cap;
    // dynamic shared memory carve-up (sized by the largest level's node_cap)
    Node *nodes = reinterpret_cast<Node *>(smem_raw);
    Node *nodesN = nodes + NC;
    int *childCnt = reinterpret_cast<int *>(nodesN + NC);   // [4*NC]
    int *newIdx = childCnt + 4 * NC;                        // [4*NC]
    uint4 *nodeFirst = reinterpret_cast<uint4 *>(newIdx + 4 * NC);   // [NC] quadrant counters in front of a node's first key (16-B aligned)
    int *rank = reinterpret_cast<int *>(nodeFirst + NC);   // [NC]
    int *arr = rank + NC;                                   // [NC] processing-order scratch
    int *ubase = arr + NC;                                  // [NC]
    unsigned char *nonEmpty = reinterpret_cast<unsigned char *>(ubase + NC); // [NC]
    unsigned char *split = nonEmpty + NC;                   // [NC]
    // key arrays of problems with at most smem_keys candidates live in shared memory (ncu, round 2: every sweep of a pass is a
    // dependent round trip to L2 on 1 000 keys -- long-scoreboard stalls behind 4.75 barrier stalls per issue); larger
    // problems keep them in global memory.  The carve-up is sized with the LAUNCH's largest node_cap (see octree_smem_bytes).
    unsigned char *key_smem = smem_raw + (size_t)g.oct_node_cap_max * (2 * sizeof(Node) + 4 * 4 * 2 + 16 + 3 * 4 + 2);
    key_smem += (16 - ((size_t)key_smem & 15)) & 15;

    uint32_t *kA = keysA_all + (size_t)f * g.keys_per_frame + L.key_base;
    uint32_t *kB = keysB_all + (size_t)f * g.keys_per_frame + L.key_base;
    uint16_t *nA = nodeA_all + (size_t)f * g.keys_per_frame + L.key_base;
    uint16_t *nB = nodeB_all + (size_t)f * g.keys_per_frame + L.key_base;
    // per-key scratch carved from the 16 B/key scan region: own-quadrant rank (u32) and quadrant (u8)
    uint32_t *E32 = reinterpret_cast<uint32_t *>(scanE_all + (size_t)f * (g.keys_per_frame + g.nlevels) + L.key_base + level);
    unsigned char *qbuf = reinterpret_cast<unsigned char *>(E32 + L.max_cand + 1);
    int *celloff = reinterpret_cast<int *>(E32);            // reused before the first scan
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;

    const int N = L.N;
    const int nCells = L.nCols * L.nRows;
    const int *ccount = cell_count + (size_t)f * g.total_cells + L.cell_base;
    const uint32_t *cslots = cell_slots + (size_t)f * g.slots_per_frame + L.slot_base;

    // ---- gather candidates in reference order (cell row-major, in-cell row-major) into kB ----
    {
        int carry = 0;
        for (int base = 0; base < nCells; base += kOctThreads) {
            const int c = base + tid;
            const int v = c < nCells ? ccount[c] : 0;
            int tot;
            const int ex = block_excl_scan<kOctWarps>(v, S, tot);
            if (c < nCells) celloff[c] = carry + ex;
            carry += tot;
        }
        if (tid == 0) { S.n = carry; ncand_out[f * g.nlevels + level] = carry; }
        __syncthreads();
        if (S.n <= smem_keys) {
            kA = reinterpret_cast<uint32_t *>(key_smem); kB = kA + smem_keys; E32 = kB + smem_keys;
            nA = reinterpret_cast<uint16_t *>(E32 + smem_keys); nB = nA + smem_keys;
            qbuf = reinterpret_cast<unsigned char *>(nB + smem_keys);
        }
        // eight lanes per cell (a cell holds ~4 candidates): four cells per warp pass instead of one, so a warp walks
        // a quarter as many dependent count -> slot load chains
        for (int c = tid >> 3; c < nCells; c += kOctThreads >> 3) {
            const int cnt = ccount[c], off = celloff[c];
            for (int i = tid & 7; i < cnt; i += 8) kB[off + i] = cslots[(size_t)c * L.cell_cap + i];
        }
        __syncthreads();
    }
    const int n = S.n;

    // ---- roots (:493-535): stable partition of kB by root index into kA ----
    const int nIni = L.nIni;
    if (nIni <= 0 || n == 0) { if (tid == 0) nkept_out[f * g.nlevels + level] = 0; return; }
    // every warp streams one contiguous segment of the key array; ranks come from ballots + running warp-uniform
    // counters, so a sweep over all keys costs two block barriers instead of two per 256 keys
    const int seg = (((n + kOctWarps - 1) / kOctWarps) + 31) & ~31;
    const int s0 = min(n, warp * seg), s1 = min(n, s0 + seg);
    int cur_size = 0;                                      // nodes in the list (lNodes.size()), kept in a register by every thread
    // ---- phase 1 (:545-625) in closed form.  A DivideNode split depends on the node's bounds only, never on its keys, so the path
    //      of every key (root, then one quadrant digit per depth) is known up front, a node of depth k is the set of keys sharing a
    //      k-digit prefix, and lNodes.size() after the pass that creates depth d is the number of distinct d-prefixes.  Counting the
    //      keys per depth-B prefix (bins) gives the sizes of every depth <= B, hence the depth e at which phase 1 ends (:621-626).  The
    //      list at that point is [depth-e nodes] ++ [single-key nodes that stopped at depth e-1] ++ ... ++ [single-key roots], and
    //      because children are pushed to the FRONT in quadrant order while the parents are walked front to back, the order inside
    //      the depth-k group is lexicographic in (root, q1 .. qk) with directions that alternate from the last digit backwards
    //      (descending, ascending, ...; the root follows q1): an XOR mask on the prefix.  So one flag scan over the concatenated
    //      prefix spaces yields every node's list index, and one counting scatter by depth-e prefix replaces the root partition and
    //      e partition passes.  The restatement is checked on the CPU against the sequential form (tools/proto/octree_closed_form.py).  When phase 1 does not
    //      end within depth B (few, clustered candidates) the sequential form below runs instead.
    int closed = 0;                                        // 0: sequential phase 1; 1: closed form, the run is finished; 2: closed form, phase 2 follows
    // The first pass of phase 2 runs from the bins as well when the prefixes one depth below e were counted (e + 1 <= B): the child
    // sizes come from the counts instead of two sweeps over the keys, and ONE scatter by depth-(e+1) prefix places every key in its
    // final node.  The tables that pass needs (child counts = fill cursors, start slots, list indices) are copied into the per-key
    // rank array of the sequential passes, which is dead until one of them runs.
    bool bins_pass = false;
    uint32_t *cfill = nullptr; uint16_t *cstart = nullptr, *clist = nullptr;
    int bins_e = 0;
    // when the scatter of the closed form is the last step of the run, it also takes each node's best key (atomicMax of best_key):
    // the final walk over every node's keys -- one thread per node, a dependent load per key -- is skipped
    unsigned long long *best64 = nullptr;
    if (!kGlobalTables && L.oct_B > 0) {
        const int B = L.oct_B;
        const float hX = L.hX;
        const int MB = nIni * (((1 << (2 * B + 2)) - 1) / 3);         // prefixes of all depths 0..B; depth k starts at nIni * (4^k - 1) / 3
        uint32_t *cnt = reinterpret_cast<uint32_t *>(childCnt);       // [MB] keys per prefix (scratch: the pass tables are not live yet)
        uint32_t *start = cnt + MB;                                   // [nIni * 4^B] first key slot of a depth-e prefix
        uint16_t *listpos = reinterpret_cast<uint16_t *>(start + (nIni << (2 * B)));   // [MB] list index of the node of a prefix
        auto lvoff = [&](int k) { return nIni * (((1 << (2 * k)) - 1) / 3); };
        for (int i = tid; i < MB; i += kOctThreads) cnt[i] = 0u;
        if (tid < 8) { S.psize[tid] = 0; S.pmulti[tid] = 0; }
        __syncthreads();
        {
            uint32_t *cB = cnt + lvoff(B);
            for (int p = tid; p < n; p += kOctThreads) {
                const uint32_t key = kB[p];
                const int x = cand_x(key), y = cand_y(key);
                const int r = min((int)((float)x / hX), nIni - 1);                                       // :519
                int x0 = (int)(hX * (float)r), x1 = (int)(hX * (float)(r + 1)), y0 = 0, y1 = L.regionH;   // :505-506
                int code = r;
                for (int d = 0; d < B; ++d) {                                                            // DivideNode :433-434, :462-476
                    const int sx = x0 + ((x1 - x0 + 1) >> 1), sy = y0 + ((y1 - y0 + 1) >> 1);
                    const bool left = x < sx, up = y < sy;
                    code = (code << 2) | (left ? (up ? 0 : 2) : (up ? 1 : 3));
                    x0 = left ? x0 : sx; x1 = left ? sx : x1; y0 = up ? y0 : sy; y1 = up ? sy : y1;
                }
                nB[p] = (uint16_t)code;
                atomicAdd(cB + code, 1u);
            }
        }
        __syncthreads();
        // keys per prefix of every depth, and per depth the number of prefixes with keys / with more than one key
        for (int k = B; k >= 0; --k) {
            const int G = nIni << (2 * k);
            uint32_t *ck = cnt + lvoff(k);
            for (int j0 = warp * 32; j0 < G; j0 += kOctThreads) {
                const int j = j0 + lane;
                uint32_t c = 0;
                if (j < G) {
                    if (k == B) c = ck[j];
                    else { const uint32_t *cc = cnt + lvoff(k + 1) + 4 * j; c = cc[0] + cc[1] + cc[2] + cc[3]; ck[j] = c; }
                }
                const uint32_t bz = __ballot_sync(0xffffffffu, c > 0), bm = __ballot_sync(0xffffffffu, c > 1);
                if (lane == 0) { if (bz) atomicAdd(&S.psize[k], __popc(bz)); if (bm) atomicAdd(&S.pmulti[k], __popc(bm)); }
            }
            __syncthreads();
        }
        // the pass loop of phase 1 on the sizes alone (:545-626); every thread runs it on the same numbers
        int e = 0;
        for (int d = 0; d < B; ++d) {
            const int sz = S.psize[d + 1];
            if (sz >= N || sz == S.psize[d]) { e = d + 1; closed = 1; break; }                            // :621
            if (sz + 3 * S.pmulti[d + 1] > N) { e = d + 1; closed = 2; break; }                           // :626
        }
        if (closed == 2 && e + 1 <= B) {
            const int G1 = nIni << (2 * e + 2), Me = nIni * (((1 << (2 * e + 2)) - 1) / 3);
            const long cap_bytes = 4L * (n <= smem_keys ? smem_keys : L.max_cand);
            bins_pass = n <= 65535 && 6L * G1 + 2L * Me + 8 <= cap_bytes;      // 16-bit start slots
        }
        if (bins_pass) {
            const int e1 = e + 1, G1 = nIni << (2 * e1), Me = nIni * (((1 << (2 * e + 2)) - 1) / 3);
            bins_e = e;
            cfill = E32; cstart = reinterpret_cast<uint16_t *>(cfill + G1); clist = cstart + G1;
            const uint32_t *c1 = cnt + lvoff(e1);
            {
                int carry = 0;
                for (int base = 0; base < G1; base += kOctThreads) {
                    const int j = base + tid;
                    const int v = j < G1 ? (int)c1[j] : 0;
                    int tot;
                    const int ex = block_excl_scan<kOctWarps>(v, S, tot);
                    if (j < G1) { cfill[j] = (uint32_t)v; cstart[j] = (uint16_t)(carry + ex); }
                    carry += tot;
                }
            }
            __syncthreads();
            {
                int carry = 0;
                for (int base = 0; base < Me; base += kOctThreads) {
                    const int t = base + tid;
                    int k = e, j = 0, flag = 0;
                    uint32_t c = 0;
                    if (t < Me) {
                        int rem = t;
                        while (rem >= (nIni << (2 * k))) { rem -= nIni << (2 * k); --k; }
                        int r = rem >> (2 * k);
                        const int path = (rem & ((1 << (2 * k)) - 1)) ^ (0x33333333 & ((1 << (2 * k)) - 1));
                        if (k & 1) r = nIni - 1 - r;
                        j = (r << (2 * k)) | path;
                        c = cnt[lvoff(k) + j];
                        const bool parent_multi = k == 0 || cnt[lvoff(k - 1) + (j >> 2)] > 1u;
                        flag = parent_multi && (k == e ? c > 0u : c == 1u);
                    }
                    int tot;
                    const int pos = carry + block_excl_scan<kOctWarps>(flag, S, tot);
                    carry += tot;
                    if (t < Me) clist[lvoff(k) + j] = flag ? (uint16_t)pos : (uint16_t)0xffffu;
                    if (flag) {
                        const int r = j >> (2 * k);
                        int x0 = (int)(hX * (float)r), x1 = (int)(hX * (float)(r + 1)), y0 = 0, y1 = L.regionH;
                        for (int d = 1; d <= k; ++d) {
                            const int q = (j >> (2 * (k - d))) & 3;
                            const int sx = x0 + ((x1 - x0 + 1) >> 1), sy = y0 + ((y1 - y0 + 1) >> 1);
                            if (q & 1) x0 = sx; else x1 = sx;
                            if (q & 2) y0 = sy; else y1 = sy;
                        }
                        const int pfx = j << (2 * (e1 - k));                 // first depth-(e+1) prefix below this node
                        Node nd;
                        nd.x0 = (short)x0; nd.x1 = (short)x1; nd.y0 = (short)y0; nd.y1 = (short)y1;
                        nd.begin = (int)cstart[pfx]; nd.count = (int)c;
                        nodes[pos] = nd;
                        nodesN[pos].begin = pfx;                             // parked until the child sizes are read (next step)
                    }
                }
            }
            cur_size = S.psize[e];
            if (tid == 0) { S.J = 0x7fffffff; S.pending = 0; S.nToExpand = 0; }
            __syncthreads();
            // child sizes of the nodes (step B of a pass) from the counts; the scratch of the analysis is dead from here on
            for (int gi = tid; gi < cur_size; gi += kOctThreads) {
                int ne = 0;
                if (nodes[gi].count > 1) {
                    const int pfx = nodesN[gi].begin;
                    const int c0 = (int)cfill[pfx], c1v = (int)cfill[pfx + 1], c2 = (int)cfill[pfx + 2], c3 = (int)cfill[pfx + 3];
                    childCnt[4 * gi] = c0; childCnt[4 * gi + 1] = c1v; childCnt[4 * gi + 2] = c2; childCnt[4 * gi + 3] = c3;
                    ne = (c0 > 0) + (c1v > 0) + (c2 > 0) + (c3 > 0);
                }
                nonEmpty[gi] = (unsigned char)ne;
            }
            __syncthreads();
        } else if (closed) {
            // first key slot of every depth-e prefix (exclusive scan in natural order)
            const int Ge = nIni << (2 * e);
            uint32_t *ce = cnt + lvoff(e);
            {
                int carry = 0;
                for (int base = 0; base < Ge; base += kOctThreads) {
                    const int j = base + tid;
                    const int v = j < Ge ? (int)ce[j] : 0;
                    int tot;
                    const int ex = block_excl_scan<kOctWarps>(v, S, tot);
                    if (j < Ge) start[j] = (uint32_t)(carry + ex);
                    carry += tot;
                }
            }
            __syncthreads();
            // the list: depth e first, then the single-key nodes of depth e-1, ..., 0; inside a depth the order is the XOR-masked prefix
            {
                const int Me = nIni * (((1 << (2 * e + 2)) - 1) / 3);
                int carry = 0;
                for (int base = 0; base < Me; base += kOctThreads) {
                    const int t = base + tid;
                    int k = e, j = 0, flag = 0;
                    uint32_t c = 0;
                    if (t < Me) {
                        int rem = t;
                        while (rem >= (nIni << (2 * k))) { rem -= nIni << (2 * k); --k; }            // segment k holds nIni * 4^k prefixes
                        int r = rem >> (2 * k);
                        const int path = (rem & ((1 << (2 * k)) - 1)) ^ (0x33333333 & ((1 << (2 * k)) - 1));
                        if (k & 1) r = nIni - 1 - r;
                        j = (r << (2 * k)) | path;
                        c = cnt[lvoff(k) + j];
                        const bool parent_multi = k == 0 || cnt[lvoff(k - 1) + (j >> 2)] > 1u;
                        flag = parent_multi && (k == e ? c > 0u : c == 1u);
                    }
                    int tot;
                    const int pos = carry + block_excl_scan<kOctWarps>(flag, S, tot);
                    carry += tot;
                    if (t < Me) listpos[lvoff(k) + j] = flag ? (uint16_t)pos : (uint16_t)0xffffu;
                    if (flag) {
                        if (closed == 1) reinterpret_cast<unsigned long long *>(E32)[pos] = 0ull;
                        const int r = j >> (2 * k);
                        int x0 = (int)(hX * (float)r), x1 = (int)(hX * (float)(r + 1)), y0 = 0, y1 = L.regionH;
                        for (int d = 1; d <= k; ++d) {
                            const int q = (j >> (2 * (k - d))) & 3;
                            const int sx = x0 + ((x1 - x0 + 1) >> 1), sy = y0 + ((y1 - y0 + 1) >> 1);
                            if (q & 1) x0 = sx; else x1 = sx;
                            if (q & 2) y0 = sy; else y1 = sy;
                        }
                        Node nd;
                        nd.x0 = (short)x0; nd.x1 = (short)x1; nd.y0 = (short)y0; nd.y1 = (short)y1;
                        nd.begin = (int)start[j << (2 * (e - k))]; nd.count = (int)c;
                        nodes[pos] = nd;
                    }
                }
            }
            __syncthreads();
            // keys to their nodes' ranges (any order inside a range: the final choice breaks score ties by candidate order explicitly)
            for (int p = tid; p < n; p += kOctThreads) {
                const uint32_t key = kB[p];
                const int code = nB[p];
                const int ge = code >> (2 * (B - e));
                // its depth-e node, or -- when an ancestor prefix already held this key alone -- that single-key node (exactly one
                // prefix of the key's path is a node)
                int node = listpos[lvoff(e) + ge];
                for (int k = e - 1; node == 0xffff && k >= 0; --k) node = listpos[lvoff(k) + (code >> (2 * (B - k)))];
                const int dst = (int)start[ge] + (int)atomicSub(ce + ge, 1u) - 1;
                kA[dst] = key; nA[dst] = (uint16_t)node;
                if (closed == 1) atomicMax(reinterpret_cast<unsigned long long *>(E32) + node, best_key(key, L));
            }
            if (closed == 1) best64 = reinterpret_cast<unsigned long long *>(E32);
            cur_size = S.psize[e];
            __syncthreads();
        }
    }
    if (!closed) {
        const float hX = L.hX;
        int placed = 0, nroots = 0;
        for (int r0 = 0; r0 < nIni; r0 += 4) {                            // four roots per sweep
            uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            for (int p0 = s0; p0 < s1; p0 += 32) {
                const int p = p0 + lane;
                int q = 4;
                if (p < s1) { const int ri = min((int)((float)cand_x(kB[p]) / hX), nIni - 1) - r0; if (ri >= 0 && ri < 4) q = ri; }   // :519
                c0 += __popc(__ballot_sync(0xffffffffu, q == 0)); c1 += __popc(__ballot_sync(0xffffffffu, q == 1));
                c2 += __popc(__ballot_sync(0xffffffffu, q == 2)); c3 += __popc(__ballot_sync(0xffffffffu, q == 3));
            }
            if (lane == 0) S.warp_v[warp] = make_uint4(c0, c1, c2, c3);
            __syncthreads();
            uint4 run = make_uint4(0, 0, 0, 0), tot = run;
            for (int w = 0; w < kOctWarps; ++w) { const uint4 t = S.warp_v[w]; if (w < warp) run = add4(run, t); tot = add4(tot, t); }
            const int off0 = placed, off1 = off0 + tot.x, off2 = off1 + tot.y, off3 = off2 + tot.z;
            const int id0 = nroots, id1 = id0 + (tot.x > 0), id2 = id1 + (tot.y > 0), id3 = id2 + (tot.z > 0);
            for (int p0 = s0; p0 < s1; p0 += 32) {
                const int p = p0 + lane;
                int q = 4; uint32_t key = 0;
                if (p < s1) { key = kB[p]; const int ri = min((int)((float)cand_x(key) / hX), nIni - 1) - r0; if (ri >= 0 && ri < 4) q = ri; }
                const uint32_t b0 = __ballot_sync(0xffffffffu, q == 0), b1 = __ballot_sync(0xffffffffu, q == 1);
                const uint32_t b2 = __ballot_sync(0xffffffffu, q == 2), b3 = __ballot_sync(0xffffffffu, q == 3);
                if (q < 4) {
                    const int dst = q == 0 ? off0 + run.x + __popc(b0 & lt) : q == 1 ? off1 + run.y + __popc(b1 & lt)
                                  : q == 2 ? off2 + run.z + __popc(b2 & lt) : off3 + run.w + __popc(b3 & lt);
                    kA[dst] = key; nA[dst] = (uint16_t)(q == 0 ? id0 : q == 1 ? id1 : q == 2 ? id2 : id3);
                }
                run.x += __popc(b0); run.y += __popc(b1); run.z += __popc(b2); run.w += __popc(b3);
            }
            if (tid == 0) {
                const int cnts[4] = { (int)tot.x, (int)tot.y, (int)tot.z, (int)tot.w }, offs[4] = { off0, off1, off2, off3 };
                int id = nroots;
                for (int k = 0; k < 4 && r0 + k < nIni; ++k)
                    if (cnts[k] > 0) {
                        Node nd;
                        nd.x0 = (short)(int)(hX * (float)(r0 + k)); nd.x1 = (short)(int)(hX * (float)(r0 + k + 1));   // :505-506
                        nd.y0 = 0; nd.y1 = (short)L.regionH;
                        nd.begin = offs[k]; nd.count = cnts[k];
                        nodes[id++] = nd;
                    }
            }
            nroots = id3 + (tot.w > 0);
            placed = off3 + tot.w;
            __syncthreads();
        }
        cur_size = nroots;                                 // block-uniform: every thread derived it from the same totals
    }

    bool phase2 = closed == 2;
#pragma unroll 1
    for (; closed != 1;) {
        const int size = cur_size;
        const int prevSize = size;

        // ---- A: stable rank of every key inside its future child (warp-streaming scan, see above) ----
        if (!bins_pass) {
            uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
            for (int p0 = s0; p0 < s1; p0 += 32 * kU) {
                int gis[kU]; uint32_t keys[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int p = p0 + 32 * u + lane;
                    gis[u] = 0; keys[u] = 0;
                    if (p < s1) { gis[u] = nA[p]; keys[u] = kA[p]; }
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int p = p0 + 32 * u + lane;
                    int q = 4;
                    if (p < s1) {
                        const Node nd = nodes[gis[u]];
                        if (nd.count > 1) q = quadrant(keys[u], nd);
                        qbuf[p] = (unsigned char)q;
                    }
                    c0 += __popc(__ballot_sync(0xffffffffu, q == 0)); c1 += __popc(__ballot_sync(0xffffffffu, q == 1));
                    c2 += __popc(__ballot_sync(0xffffffffu, q == 2)); c3 += __popc(__ballot_sync(0xffffffffu, q == 3));
                }
            }
            if (lane == 0) S.warp_v[warp] = make_uint4(c0, c1, c2, c3);
            if (tid == 0) { S.J = 0x7fffffff; S.pending = 0; }
            __syncthreads();
            if (tid == 0) S.nToExpand = 0;                 // behind a barrier: the previous pass's tail may still be reading it
            uint4 run = make_uint4(0, 0, 0, 0);
            for (int w = 0; w < warp; ++w) run = add4(run, S.warp_v[w]);
            for (int p0 = s0; p0 < s1; p0 += 32 * kU) {
                int qs[kU], gis[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int p = p0 + 32 * u + lane;
                    qs[u] = 4; gis[u] = 0;
                    if (p < s1) { qs[u] = (int)qbuf[p]; gis[u] = nA[p]; }
                }
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    const int p = p0 + 32 * u + lane;
                    const int q = qs[u];
                    const uint32_t b0 = __ballot_sync(0xffffffffu, q == 0), b1 = __ballot_sync(0xffffffffu, q == 1);
                    const uint32_t b2 = __ballot_sync(0xffffffffu, q == 2), b3 = __ballot_sync(0xffffffffu, q == 3);
                    if (q < 4) {
                        const uint4 ex = make_uint4(run.x + __popc(b0 & lt), run.y + __popc(b1 & lt), run.z + __popc(b2 & lt), run.w + __popc(b3 & lt));
                        E32[p] = comp(ex, q);
                        const int gi = gis[u];
                        const Node nd = nodes[gi];
                        if (p == nd.begin) nodeFirst[gi] = ex;
                        if (p == nd.begin + nd.count - 1)             // counters just behind the node's last key
                            *reinterpret_cast<uint4 *>(childCnt + 4 * gi) = make_uint4(ex.x + (q == 0), ex.y + (q == 1), ex.z + (q == 2), ex.w + (q == 3));
                    }
                    run.x += __popc(b0); run.y += __popc(b1); run.z += __popc(b2); run.w += __popc(b3);
                }
            }
            __syncthreads();
        }
        // ---- B: child sizes ----
        if (!bins_pass)
        for (int gi = tid; gi < size; gi += kOctThreads) {
            const Node nd = nodes[gi];
            int ne = 0;
            if (nd.count > 1) {
                const uint4 e0 = nodeFirst[gi], e1 = *reinterpret_cast<const uint4 *>(childCnt + 4 * gi);
                const int c0 = e1.x - e0.x, c1 = e1.y - e0.y, c2 = e1.z - e0.z, c3 = e1.w - e0.w;
                childCnt[4 * gi] = c0; childCnt[4 * gi + 1] = c1; childCnt[4 * gi + 2] = c2; childCnt[4 * gi + 3] = c3;
                ne = (c0 > 0) + (c1 > 0) + (c2 > 0) + (c3 > 0);
            }
            nonEmpty[gi] = (unsigned char)ne;
        }
        __syncthreads();
        // ---- C: which nodes split, and in which processing order ----
        if (!phase2) {
            int carry = 0;
            for (int base = 0; base < size; base += kOctThreads) {
                const int gi = base + tid;
                const int fl = gi < size ? (nodes[gi].count > 1) : 0;
                int tot;
                const int ex = block_excl_scan<kOctWarps>(fl, S, tot);
                if (gi < size) { split[gi] = (unsigned char)fl; rank[gi] = carry + ex; }
                carry += tot;
            }
            if (tid == 0) S.nsplit = carry;
            __syncthreads();
        } else {
            // sort key (:638 ascending, walked from the back): more keys first; ties -> later created
            // first.  Pending nodes were all created in the previous pass, where creation order is
            // the reverse of list order, so "later created" == smaller list index.
            for (int gi = tid; gi < size; gi += kOctThreads) {
                const int cg = nodes[gi].count;
                int r = -1;
                if (cg > 1) {
                    r = 0;
                    for (int j = 0; j < size; ++j) {
                        const int cj = nodes[j].count;
                        r += (cj > 1) && (cj > cg || (cj == cg && j < gi));
                    }
                    atomicAdd(&S.pending, 1);
                }
                rank[gi] = r;
                if (r >= 0) arr[r] = (int)nonEmpty[gi] - 1;   // gain of splitting this node
            }
            __syncthreads();
            const int P = S.pending;
            int carry = 0;
            for (int base = 0; base < P; base += kOctThreads) {
                const int r = base + tid;
                const int v = r < P ? arr[r] : 0;
                int tot;
                const int ex = block_excl_scan<kOctWarps>(v, S, tot);
                if (r < P && size + carry + ex + v >= N) atomicMin(&S.J, r);   // :684 break
                carry += tot;
            }
            __syncthreads();
            const int J = min(S.J, P - 1);
            for (int gi = tid; gi < size; gi += kOctThreads) split[gi] = (unsigned char)(rank[gi] >= 0 && rank[gi] <= J);
            if (tid == 0) S.nsplit = J + 1;
            __syncthreads();
        }
        const int nsplit = S.nsplit;
        // ---- D: creation index of every child, list slot of every untouched node ----
        for (int gi = tid; gi < size; gi += kOctThreads) if (split[gi]) arr[rank[gi]] = nonEmpty[gi];
        __syncthreads();
        {
            // one packed scan gives both: low half = children created before processing rank i, high half = untouched
            // nodes in front of list position i (sums stay below 2^16: node_cap < 65536, checked at geometry build)
            int carry = 0;
            const int span = nsplit > size ? nsplit : size;
            for (int base = 0; base < span; base += kOctThreads) {
                const int i = base + tid;
                const int v = (i < nsplit ? arr[i] : 0) | ((i < size ? !split[i] : 0) << 16);
                int tot;
                const int ex = carry + block_excl_scan<kOctWarps>(v, S, tot);
                if (i < nsplit) arr[i] = ex & 0xffff;
                if (i < size) ubase[i] = ex >> 16;
                carry += tot;
            }
            if (tid == 0) { S.C = carry & 0xffff; S.U = carry >> 16; }
            __syncthreads();
        }
        const int C = S.C, U = S.U;
        // ---- E: new node table ----
        {
            int expand = 0;
            for (int gi = tid; gi < size; gi += kOctThreads) {
                const Node nd = nodes[gi];
                if (split[gi]) {
                    const int sx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1), sy = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
                    int ci = arr[rank[gi]], off = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cnt = childCnt[4 * gi + q];
                        int pos = -1;
                        if (cnt > 0) {
                            pos = C - 1 - ci; ++ci;                 // push_front: latest created first
                            Node ch;
                            ch.x0 = (q & 1) ? (short)sx : nd.x0; ch.x1 = (q & 1) ? nd.x1 : (short)sx;
                            ch.y0 = (q & 2) ? (short)sy : nd.y0; ch.y1 = (q & 2) ? nd.y1 : (short)sy;
                            ch.begin = nd.begin + off; ch.count = cnt;
                            nodesN[pos] = ch;
                            if (bins_pass) reinterpret_cast<unsigned long long *>(nodeFirst)[pos] = 0ull;
                            off += cnt;
                            expand += cnt > 1;
                        }
                        newIdx[4 * gi + q] = pos;
                    }
                } else {
                    const int pos = C + ubase[gi];
                    nodesN[pos] = nd;
                    newIdx[4 * gi] = pos;
                    if (bins_pass) reinterpret_cast<unsigned long long *>(nodeFirst)[pos] = 0ull;
                }
            }
            if (expand) atomicAdd(&S.nToExpand, expand);
            __syncthreads();
        }
        // ---- F: move keys ----
        if (bins_pass) {
            // from the gathered keys (kB, prefix codes in nB) straight to the final ranges: a key's node is the single-key node of the
            // shallowest depth at which its prefix is alone, else its depth-e node; split nodes hand it on to the child of its next digit
            const int B = L.oct_B, e = bins_e;
            auto lvoff = [&](int k) { return nIni * (((1 << (2 * k)) - 1) / 3); };
            const bool last = C + U >= N || C + U == prevSize;      // the loop's exit test below: this scatter ends the run
            unsigned long long *b64 = reinterpret_cast<unsigned long long *>(nodeFirst);   // zeroed per new node in step E
            for (int p = tid; p < n; p += kOctThreads) {
                const uint32_t key = kB[p];
                const int code = nB[p];
                int gi = clist[lvoff(e) + (code >> (2 * (B - e)))];
                for (int k = e - 1; gi == 0xffff && k >= 0; --k) gi = clist[lvoff(k) + (code >> (2 * (B - k)))];
                const int pfx1 = code >> (2 * (B - e - 1));
                const int fin = split[gi] ? newIdx[4 * gi + (pfx1 & 3)] : newIdx[4 * gi];
                const int dst = (int)cstart[pfx1] + (int)atomicSub(cfill + pfx1, 1u) - 1;
                kA[dst] = key; nA[dst] = (uint16_t)fin;
                if (last) atomicMax(b64 + fin, best_key(key, L));
            }
            if (last) best64 = b64;
        } else
        for (int pb = tid; pb < n; pb += kOctThreads * kU) {
            int gis[kU], qs[kU]; uint32_t keys[kU], es[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int p = pb + u * kOctThreads;
                gis[u] = 0; qs[u] = 0; keys[u] = 0; es[u] = 0;
                if (p < n) {
                    gis[u] = nA[p]; keys[u] = kA[p];
                    if (kU > 1) { qs[u] = qbuf[p]; es[u] = E32[p]; }   // ahead of their use (stale for unsplit nodes and unused there); kU == 1 loads them on demand
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int p = pb + u * kOctThreads;
                if (p >= n) break;
                const int gi = gis[u];
                const uint32_t key = keys[u];
                if (split[gi]) {
                    const Node nd = nodes[gi];
                    const int q = kU > 1 ? qs[u] : (int)qbuf[p];
                    int off = 0;
#pragma unroll
                    for (int qq = 0; qq < 3; ++qq) if (qq < q) off += childCnt[4 * gi + qq];
                    const int np = nd.begin + off + (int)((kU > 1 ? es[u] : E32[p]) - comp(nodeFirst[gi], q));
                    kB[np] = key; nB[np] = (uint16_t)newIdx[4 * gi + q];
                } else {
                    kB[p] = key; nB[p] = (uint16_t)newIdx[4 * gi];
                }
            }
        }
        __syncthreads();
        if (bins_pass) { Node *v = nodes; nodes = nodesN; nodesN = v; bins_pass = false; }   // the keys went kB -> kA
        else { uint32_t *t = kA; kA = kB; kB = t; uint16_t *u = nA; nA = nB; nB = u; Node *v = nodes; nodes = nodesN; nodesN = v; }
        const int newSize = C + U;
        const int nToExpand = S.nToExpand;                 // reset only behind the next pass's first barrier
        cur_size = newSize;
        if (newSize >= N || newSize == prevSize) break;                 // :621-624 / :688
        if (!phase2 && newSize + nToExpand * 3 > N) phase2 = true;      // :626
    }

    // ---- :697-715 keep the best key of every node (first maximum wins), list order ----
    const int size = cur_size;
    uint32_t *kept = kept_out + (size_t)f * g.kept_total + L.kept_base;
    if (best64) {
        for (int gi = tid; gi < size && gi < L.kept_cap; gi += kOctThreads) kept[gi] = best_key_decode(best64[gi]);
    } else
    for (int gi = tid; gi < size && gi < L.kept_cap; gi += kOctThreads) {
        const Node nd = nodes[gi];
        // the first maximum in CANDIDATE order wins (:700-712).  The keys of a node are not kept in that order (closed-form phase 1), so
        // ties are broken explicitly: candidates come cell by cell (row-major cells, ORBextractor.cpp:745-786), row-major inside a cell
        uint32_t best = kA[nd.begin];
        for (int k = 1; k < nd.count; ++k) {
            const uint32_t key = kA[nd.begin + k];
            if (cand_score(key) > cand_score(best)) best = key;
            else if (cand_score(key) == cand_score(best)) {
                const int ya = cand_y(key), yb = cand_y(best), xa = cand_x(key), xb = cand_x(best);
                // (v - 3) / cell size by a host-made reciprocal (exact for v < 4096 and cells <= 64 px, see octree_closed_depth's neighbour)
                const int ia = ((ya - 3) * L.inv_hCell) >> 18, ib = ((yb - 3) * L.inv_hCell) >> 18;
                const int ja = ((xa - 3) * L.inv_wCell) >> 18, jb = ((xb - 3) * L.inv_wCell) >> 18;
                const bool before = ia != ib ? ia < ib : ja != jb ? ja < jb : ya != yb ? ya < yb : xa < xb;
                if (before) best = key;
            }
        }
        kept[gi] = best;
    }
    if (tid == 0) nkept_out[f * g.nlevels + level] = min(size, L.kept_cap);
}

// Depth B of the closed-form phase 1 for one level: the smallest depth whose prefix space holds twice the quota (a well
// spread level reaches its quota one or two depths earlier), shrunk until the scratch -- counts of all depths, start slots
// and list indices, laid over the pass tables behind the two node tables -- fits.  0 switches the closed form off.
int octree_closed_depth(const Geo &g, int level)
{
    static int enabled = -1;
    if (enabled < 0) { const char *e = std::getenv("ORBX_OCT_CLOSED"); enabled = e ? std::atoi(e) : 1; }
    const LevelGeom &L = g.lv[level];
    if (!enabled || L.nIni < 1 || L.nIni > 16 || L.N < 1) return 0;
    const long avail = (long)g.oct_node_cap_max * (2 * (long)sizeof(Node) + 4 * 4 * 2 + 16 + 3 * 4 + 2) - (long)L.node_cap * 2 * (long)sizeof(Node) - 64;
    int B = 1;
    while (B < 6 && ((long)L.nIni << (2 * B)) < 2L * L.N) ++B;
    for (; B >= 1; --B) {
        const long G = (long)L.nIni << (2 * B), M = (long)L.nIni * (((1L << (2 * B + 2)) - 1) / 3);
        if (G <= 65536 && 4 * M + 4 * G + 2 * M <= avail) break;
    }
    return B < 2 ? 0 : B;
}

// bytes of one problem's node tables (no key arrays): the slice size of the global-memory fallback
int octree_table_bytes(const Geo &g)
{
    int nc = 0;
    for (int l = 0; l < g.nlevels; ++l) nc = nc > g.lv[l].node_cap ? nc : g.lv[l].node_cap;
    return (nc * (2 * (int)sizeof(Node) + 4 * 4 * 2 + 3 * 4 + 16 + 2) + 64 + 16 + 255) & ~255;
}

int octree_smem_bytes(const Geo &g)
{
    int nc = 0;
    for (int l = 0; l < g.nlevels; ++l) nc = nc > g.lv[l].node_cap ? nc : g.lv[l].node_cap;
    // 2 node tables + childCnt + newIdx (4 ints each) + rank + arr + ubase + 2 byte flags, then the shared-memory key arrays
    return nc * (2 * (int)sizeof(Node) + 4 * 4 * 2 + 3 * 4 + 16 + 2) + 64 + 16 + octree_smem_keys(g) * kOctKeyBytes;
}

// Candidates per problem that get shared-memory key arrays: about six candidates per kept feature of the largest level
// (FAST delivers 4-6x the quota on textured frames), capped so that a few blocks still fit one SM.  Larger problems run
// from global memory, so this is a performance knob, never a capacity limit.
int octree_smem_keys(const Geo &g)
{
    int nmax = 0, big = 0;
    for (int l = 0; l < g.nlevels; ++l) {
        nmax = nmax > g.lv[l].N ? nmax : g.lv[l].N;
        big = big > g.lv[l].regionW * g.lv[l].regionH ? big : g.lv[l].regionW * g.lv[l].regionH;
    }
    if (big > 1500000) return 0;                            // 4K-sized levels: tens of thousands of candidates, 1024-thread blocks
    int k = (6 * nmax + 31) & ~31;
    if (const char *e = std::getenv("ORBX_OCT_SMEM_KEYS")) k = std::atoi(e) & ~31;
    k = k < 0 ? 0 : (k > 2048 ? 2048 : k);
    // node tables that already fill most of an SM (one huge level) leave no room: such problems run from global memory
    int nc = 0;
    for (int l = 0; l < g.nlevels; ++l) nc = nc > g.lv[l].node_cap ? nc : g.lv[l].node_cap;
    if ((size_t)nc * 94 + 80 + (size_t)k * kOctKeyBytes > 112 * 1024) k = 0;
    return k;
}

int octree_configure(int smem_bytes)
{
    cudaError_t e = cudaFuncSetAttribute(k_octree<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_octree<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_octree<512, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_octree<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_octree<1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_octree<1024, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    return e == cudaSuccess ? 0 : -1;
}

void launch_octree(const Geo &g, const DevBuffers &b, int nframes, int smem_bytes, cudaStream_t s, int level_lo, int level_hi)
{
    if (level_hi > g.nlevels) level_hi = g.nlevels;
    if (level_hi <= level_lo) return;
    dim3 grd(level_hi - level_lo, nframes);                 // levels [level_lo, level_hi); default: all
    if (b.oct_scratch) {
        // node tables in global memory (see the kernel): one slice per (frame, level) of the whole batch geometry
        unsigned char *base = b.oct_scratch + ((size_t)g.frame0 * g.nlevels + level_lo) * b.oct_scratch_stride;
        if (level_hi - level_lo == g.nlevels)
            k_octree<256, 1, true><<<grd, 256, 0, s>>>(g, b.cell_count, b.cell_slots, b.keysA, b.keysB, b.nodeA, b.nodeB, b.scanE, b.ncand, b.kept, b.nkept,
                                                       0, level_lo, base, b.oct_scratch_stride);
        else
            for (int l = level_lo; l < level_hi; ++l)     // a level range: slices of consecutive frames are nlevels apart, so one launch per level
                for (int f = 0; f < nframes; ++f) {
                    Geo g1 = g; g1.frame0 = g.frame0 + f;
                    k_octree<256, 1, true><<<dim3(1, 1), 256, 0, s>>>(g1, b.cell_count, b.cell_slots, b.keysA, b.keysB, b.nodeA, b.nodeB, b.scanE, b.ncand, b.kept,
                                                               b.nkept, 0, l, b.oct_scratch + ((size_t)(g.frame0 + f) * g.nlevels + l) * b.oct_scratch_stride, b.oct_scratch_stride);
                }
        return;
    }
    // candidate counts scale with the level area: large levels (4K) get 1024 threads per problem
    int big = 0;
    for (int l = 0; l < g.nlevels; ++l) big = big > g.lv[l].regionW * g.lv[l].regionH ? big : g.lv[l].regionW * g.lv[l].regionH;
    if (big > 1500000) {
        static int variant = -1;                             // ORBX_OCT_BIG: 0 = 1024 threads, 1 / 3 = 1024 threads with 2 / 4 key groups in flight (default 3: 0.864 -> 0.731 -> 0.638 -> 0.608 ms per 32 4K frames), 2 = 512 threads x 4 groups (0.873)
        if (variant < 0) { const char *e = std::getenv("ORBX_OCT_BIG"); variant = e ? std::atoi(e) : 3; }
#define OCT_ARGS g, b.cell_count, b.cell_slots, b.keysA, b.keysB, b.nodeA, b.nodeB, b.scanE, b.ncand, b.kept, b.nkept, 0, level_lo, nullptr, 0ull
        if (variant == 0) k_octree<1024, 1><<<grd, 1024, smem_bytes, s>>>(OCT_ARGS);
        else if (variant == 1) k_octree<1024, 2><<<grd, 1024, smem_bytes, s>>>(OCT_ARGS);
        else if (variant == 3) k_octree<1024, 4><<<grd, 1024, smem_bytes, s>>>(OCT_ARGS);
        else k_octree<512, 4><<<grd, 512, smem_bytes, s>>>(OCT_ARGS);
#undef OCT_ARGS
    }
    else if (nframes <= 4)
        // small batches (low-latency path): twice the threads per problem -- half the keys per warp in every sweep
        k_octree<512, 1><<<grd, 512, smem_bytes, s>>>(g, b.cell_count, b.cell_slots, b.keysA, b.keysB, b.nodeA, b.nodeB,
                                                   b.scanE, b.ncand, b.kept, b.nkept, octree_smem_keys(g), level_lo, nullptr, 0ull);
    else
        k_octree<256, 1><<<grd, 256, smem_bytes, s>>>(g, b.cell_count, b.cell_slots, b.keysA, b.keysB, b.nodeA, b.nodeB,
                                                   b.scanE, b.ncand, b.kept, b.nkept, octree_smem_keys(g), level_lo, nullptr, 0ull);
}

} // namespace orbx

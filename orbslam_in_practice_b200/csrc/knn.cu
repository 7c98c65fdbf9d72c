// knn.cu -- ORBmatcher::DescriptorDistance (ORBmatcher.cpp:128-144) and the brute-force best-2
// candidate scan of ORBmatcher.cpp:37-62, plus acceptance (:65-67) and the shard merge.
//
// k_knn2: each thread keeps QPT queries (8 x u32 each) in registers; the block streams a database
// segment through shared memory in 4 KB tiles (cp.async double buffer, 16-byte rows read back as
// broadcast LDS.128).  Per pair: 8 x (LOP3 xor + POPC) + adds, then a branch-free best-2 update on
// a packed (distance << 22 | row) key: min of keys == smallest distance, lowest index first, which
// is exactly the reference's strict '<' scan order.  The database is split into segments over
// blockIdx.y to fill the 148 SMs; k_knn2_merge folds the per-segment (k1, k2) pairs.
// Integer-pipe bound (POPC), not HBM bound: compulsory traffic is 32 B per row.
#include "orbx_internal.cuh"

#include <climits>
#include <cstdlib>

namespace orbx {

constexpr int kDefaultCsa = 15;               // carry-save form (see hamming_row / hamming_key): 13 / 14 = explicit 3- / 4-stage, 15 = alternating rows
constexpr int kKnnThreads = 256;
constexpr int kQPT = 4;                       // queries per thread
constexpr int kQPB = kKnnThreads * kQPT;      // queries per block
constexpr int kTileRows = 128;                // database rows per shared-memory tile (4 KB)
constexpr int kIdxBits = 22;                  // rows per segment < 2^22
constexpr uint32_t kNoKey = 0xffffffffu;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Hamming distance of two 256-bit rows.  CSA = number of carry-save adder stages folded in front of the POPCs
// (Harley-Seal): a CSA turns three words of equal weight into a sum word (a^b^c, one LOP3) and a carry word of
// twice the weight (maj(a,b,c), one LOP3), trading one POPC (4 lanes/clk/SMSP) for two LOP3 (16 lanes/clk/SMSP).
//   CSA 0: 8 POPC   CSA 2: 6 POPC + 4 LOP3   CSA 3: 5 POPC + 6 LOP3   CSA 4: 4 POPC + 8 LOP3
template <int CSA>
__device__ __forceinline__ int hamming_row(const uint32_t (&q)[8], const uint4 d0, const uint4 d1)
{
    const uint32_t x0 = q[0] ^ d0.x, x1 = q[1] ^ d0.y, x2 = q[2] ^ d0.z, x3 = q[3] ^ d0.w;
    const uint32_t x4 = q[4] ^ d1.x, x5 = q[5] ^ d1.y, x6 = q[6] ^ d1.z, x7 = q[7] ^ d1.w;
    if (CSA == 0)
        return __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3) + __popc(x4) + __popc(x5) + __popc(x6) + __popc(x7);
    const uint32_t s0 = x0 ^ x1 ^ x2, c0 = (x0 & x1) | (x2 & (x0 | x1));
    const uint32_t s1 = x3 ^ x4 ^ x5, c1 = (x3 & x4) | (x5 & (x3 | x4));
    if (CSA == 2)
        return __popc(s0) + __popc(s1) + __popc(x6) + __popc(x7) + 2 * (__popc(c0) + __popc(c1));
    const uint32_t s2 = s0 ^ s1 ^ x6, c2 = (s0 & s1) | (x6 & (s0 | s1));
    if (CSA == 3)
        return __popc(s2) + __popc(x7) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
    const uint32_t s3 = c0 ^ c1 ^ c2, c3 = (c0 & c1) | (c2 & (c0 | c1));
    return __popc(s2) + __popc(x7) + 2 * __popc(s3) + 4 * __popc(c3);
}

// Explicit-PTX form of the 3- and 4-stage variants (CSA 13 / 14).  Left to itself the compiler folds the eight XORs into
// the carry-save LOP3s and ends up with ~17 logic instructions per pair instead of 8 + 6, which made the 3-stage form
// ALU-pipe bound (SASS).  Here every XOR / sum / majority is one lop3 the compiler cannot re-associate, and the
// weighted sum of the POPCs is folded into the packed key with one mad each (FMA pipe): key = row + 2^22 * distance.
__device__ __forceinline__ uint32_t lop_xor(uint32_t a, uint32_t b) { uint32_t r; asm volatile("lop3.b32 %0, %1, %2, 0, 0x3c;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t lop_sum(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t lop_maj(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
template <int CSA>
__device__ __forceinline__ uint32_t hamming_key(const uint32_t (&q)[8], const uint4 d0, const uint4 d1, uint32_t row)
{
    const uint32_t x0 = lop_xor(q[0], d0.x), x1 = lop_xor(q[1], d0.y), x2 = lop_xor(q[2], d0.z), x3 = lop_xor(q[3], d0.w);
    const uint32_t x4 = lop_xor(q[4], d1.x), x5 = lop_xor(q[5], d1.y), x6 = lop_xor(q[6], d1.z), x7 = lop_xor(q[7], d1.w);
    const uint32_t s0 = lop_sum(x0, x1, x2), c0 = lop_maj(x0, x1, x2);
    const uint32_t s1 = lop_sum(x3, x4, x5), c1 = lop_maj(x3, x4, x5);
    const uint32_t s2 = lop_sum(s0, s1, x6), c2 = lop_maj(s0, s1, x6);
    constexpr uint32_t W1 = 1u << kIdxBits, W2 = 2u << kIdxBits, W4 = 4u << kIdxBits;
    uint32_t key = mad_u32(__popc(s2), W1, row);
    key = mad_u32(__popc(x7), W1, key);
    if (CSA == 13) {
        key = mad_u32(__popc(c0), W2, key);
        key = mad_u32(__popc(c1), W2, key);
        return mad_u32(__popc(c2), W2, key);
    }
    const uint32_t s3 = lop_sum(c0, c1, c2), c3 = lop_maj(c0, c1, c2);
    key = mad_u32(__popc(s3), W2, key);
    return mad_u32(__popc(c3), W4, key);
}

// one block's share of a scan: queries [q0, q0 + kQPB) against database segment `seg` (rows [seg * seg_rows, ...))
template <int CSA>
__device__ __forceinline__ void knn2_block(const uint4 *__restrict__ query, int nq, const uint4 *__restrict__ db, int ndb, int seg_rows,
                                           const int seg, const int q0, uint2 *__restrict__ partial)
{
    __shared__ __align__(16) uint4 tile[2][kTileRows * 2];

    const int tid = threadIdx.x;
    const int row0 = seg * seg_rows;
    const int rows = min(seg_rows, ndb - row0);

    uint32_t q[kQPT][8];
    uint32_t k1[kQPT], k2[kQPT];
#pragma unroll
    for (int i = 0; i < kQPT; ++i) {
        const int qi = min(q0 + i * kKnnThreads + tid, nq - 1);
        const uint4 a = __ldg(query + 2 * (size_t)qi), b = __ldg(query + 2 * (size_t)qi + 1);
        q[i][0] = a.x; q[i][1] = a.y; q[i][2] = a.z; q[i][3] = a.w;
        q[i][4] = b.x; q[i][5] = b.y; q[i][6] = b.z; q[i][7] = b.w;
        k1[i] = kNoKey; k2[i] = kNoKey;
    }

    const int ntiles = (rows + kTileRows - 1) / kTileRows;
    const uint4 *seg_db = db + 2 * (size_t)row0;
    // prologue: tile 0
    {
        const int r = tid;                        // one uint4 per thread: 256 x 16 B = 4 KB
        if (r < min(rows, kTileRows) * 2) cp_async16(&tile[0][r], seg_db + r);
        cp_async_commit();
    }
    for (int t = 0; t < ntiles; ++t) {
        const int cur = t & 1;
        if (t + 1 < ntiles) {
            const int nrow = min(rows - (t + 1) * kTileRows, kTileRows);
            if (tid < nrow * 2) cp_async16(&tile[cur ^ 1][tid], seg_db + 2 * (size_t)(t + 1) * kTileRows + tid);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const int trows = min(rows - t * kTileRows, kTileRows);
        const uint32_t rbase = (uint32_t)(t * kTileRows);
        if (CSA >= 10) {
            // two rows per step: five min/max fold both keys into the running best-2 (six when taken one at a time);
            // a padding row past the tile's end gets the no-key sentinel
            for (int r = 0; r < trows; r += 2) {
                uint32_t ka[kQPT];
                {
                    const uint4 d0 = tile[cur][2 * r], d1 = tile[cur][2 * r + 1];
#pragma unroll
                    for (int i = 0; i < kQPT; ++i) ka[i] = hamming_key<CSA == 15 ? 13 : CSA>(q[i], d0, d1, rbase + (uint32_t)r);
                }
                const uint4 d0 = tile[cur][2 * r + 2], d1 = tile[cur][2 * r + 3];
                const bool two = r + 1 < trows;
#pragma unroll
                for (int i = 0; i < kQPT; ++i) {
                    const uint32_t kb = two ? hamming_key<CSA == 15 ? 14 : CSA>(q[i], d0, d1, rbase + (uint32_t)r + 1u) : kNoKey;
                    const uint32_t lo = min(ka[i], kb), hi = max(ka[i], kb);
                    const uint32_t t1 = max(k1[i], lo);
                    k1[i] = min(k1[i], lo);
                    k2[i] = min(min(k2[i], t1), hi);
                }
            }
        } else {
#pragma unroll 2
            for (int r = 0; r < trows; ++r) {
                const uint4 d0 = tile[cur][2 * r], d1 = tile[cur][2 * r + 1];
#pragma unroll
                for (int i = 0; i < kQPT; ++i) {
                    const int dist = hamming_row<CSA>(q[i], d0, d1);
                    const uint32_t key = ((uint32_t)dist << kIdxBits) | (rbase + (uint32_t)r);
                    const uint32_t hi = max(k1[i], key);
                    k1[i] = min(k1[i], key);
                    k2[i] = min(k2[i], hi);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < kQPT; ++i) {
        const int qi = q0 + i * kKnnThreads + tid;
        if (qi < nq) partial[(size_t)seg * nq + qi] = make_uint2(k1[i], k2[i]);
    }
}

template <int CSA>
__global__ void __launch_bounds__(kKnnThreads, 4)
k_knn2(const uint4 *__restrict__ query, int nq, const uint4 *__restrict__ db, int ndb, int seg_rows,
       uint2 *__restrict__ partial)
{
    knn2_block<CSA>(query, nq, db, ndb, seg_rows, blockIdx.y, blockIdx.x * kQPB, partial);
}

// Many small, independent scans in ONE launch (blockIdx.z = pair): descriptor rows of frame pair_a[p] against those of frame
// pair_b[p], both inside the extractor's output layout desc[F][cap][32] with counts[F] on the device -- the left -> right
// matching of a batch of stereo pairs (BASELINE configs[2]).  A 2 000 x 2 000 scan alone is 32 blocks: launch-latency bound
// (64 us per pair through k_knn2 + merge + select); a batch of them fills the GPU.
template <int CSA>
__global__ void __launch_bounds__(kKnnThreads, 4)
k_knn2_pairs(const uint4 *__restrict__ desc, const int *__restrict__ counts, int cap, const int *__restrict__ pair_a, const int *__restrict__ pair_b,
             int seg_rows, int nseg_max, uint2 *__restrict__ partial)
{
    const int p = blockIdx.z;
    const int a = pair_a[p], b = pair_b[p];
    const int nq = min(max(counts[a], 0), cap), ndb = min(max(counts[b], 0), cap);
    const int q0 = blockIdx.x * kQPB, seg = blockIdx.y;
    if (q0 >= nq || seg * seg_rows >= ndb) return;                      // block-uniform
    knn2_block<CSA>(desc + 2 * (size_t)a * cap, nq, desc + 2 * (size_t)b * cap, ndb, seg_rows, seg, q0,
                    partial + (size_t)p * nseg_max * cap + (size_t)seg * (cap - nq));   // knn2_block indexes [seg * nq + qi]: land on [seg * cap + qi]
}

// fold a pair's segments (ascending row ranges) and apply TH_LOW / the float ratio test (ORBmatcher.cpp:65-67) in the same pass
__global__ void __launch_bounds__(256)
k_knn2_pairs_merge(const uint2 *__restrict__ partial, const int *__restrict__ counts, int cap, const int *__restrict__ pair_a,
                   const int *__restrict__ pair_b, int seg_rows, int nseg_max, int *__restrict__ d1, int *__restrict__ idx1, int *__restrict__ d2,
                   int th_low, float ratio, int *__restrict__ match)
{
    const int p = blockIdx.y, qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= cap) return;
    const int nq = min(max(counts[pair_a[p]], 0), cap), ndb = min(max(counts[pair_b[p]], 0), cap);
    const size_t o = (size_t)p * cap + qi;
    if (qi >= nq) { d1[o] = INT_MAX; idx1[o] = -1; d2[o] = INT_MAX; if (match) match[o] = -1; return; }   // rows past the frame's count: defined, empty
    const int nseg = (ndb + seg_rows - 1) / seg_rows;
    uint32_t b1 = kNoKey, b2 = kNoKey; int bseg = 0;
    const uint2 *ps = partial + (size_t)p * nseg_max * cap + qi;
    for (int s = 0; s < nseg; ++s) {
        const uint2 k = ps[(size_t)s * cap];
        // keys are (distance << 22 | row in segment): compare distances first, earlier segment wins ties (ascending rows)
        const uint32_t ks[2] = { k.x, k.y };
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (ks[j] == kNoKey) continue;
            const uint32_t dist = ks[j] >> kIdxBits;
            if (b1 == kNoKey || dist < (b1 >> kIdxBits)) { b2 = b1; b1 = ks[j]; bseg = s; }
            else if (b2 == kNoKey || dist < (b2 >> kIdxBits)) b2 = ks[j];
        }
    }
    const int best = b1 == kNoKey ? INT_MAX : (int)(b1 >> kIdxBits), best2 = b2 == kNoKey ? INT_MAX : (int)(b2 >> kIdxBits);
    const int bidx = b1 == kNoKey ? -1 : (int)(b1 & ((1u << kIdxBits) - 1u)) + bseg * seg_rows;
    d1[o] = best; idx1[o] = bidx; d2[o] = best2;
    if (match) match[o] = (bidx >= 0 && best <= th_low && (float)best < __fmul_rn((float)best2, ratio)) ? bidx : -1;
}

// fold the per-segment pairs: segments are ascending index ranges, so (distance, segment, row)
// lexicographic order == (distance, global index) order.
__global__ void __launch_bounds__(256)
k_knn2_merge(const uint2 *__restrict__ partial, int nq, int nseg, int seg_rows, int index_base,
             int *__restrict__ d1, int *__restrict__ idx1, int *__restrict__ d2)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    unsigned long long b1 = ~0ull, b2 = ~0ull;
    for (int s = 0; s < nseg; ++s) {
        const uint2 p = partial[(size_t)s * nq + qi];
        const uint32_t ks[2] = { p.x, p.y };
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (ks[j] == kNoKey) continue;
            const unsigned long long dist = ks[j] >> kIdxBits, row = (ks[j] & ((1u << kIdxBits) - 1u)) + (unsigned long long)s * seg_rows;
            const unsigned long long key = (dist << 40) | row;
            const unsigned long long hi = key > b1 ? key : b1;
            b1 = key < b1 ? key : b1;
            b2 = hi < b2 ? hi : b2;
        }
    }
    d1[qi] = b1 == ~0ull ? INT_MAX : (int)(b1 >> 40);
    idx1[qi] = b1 == ~0ull ? -1 : (int)(b1 & ((1ull << 40) - 1)) + index_base;
    d2[qi] = b2 == ~0ull ? INT_MAX : (int)(b2 >> 40);
}

__global__ void k_knn2_empty(int nq, int *d1, int *idx1, int *d2)
{
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi < nq) { d1[qi] = INT_MAX; idx1[qi] = -1; d2[qi] = INT_MAX; }
}

__global__ void k_hamming_pairs(const uint4 *__restrict__ a, const uint4 *__restrict__ b, int n, int *__restrict__ dist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 a0 = a[2 * (size_t)i], a1 = a[2 * (size_t)i + 1], b0 = b[2 * (size_t)i], b1 = b[2 * (size_t)i + 1];
    dist[i] = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
              __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void k_ratio_select(const int *__restrict__ d1, const int *__restrict__ idx1, const int *__restrict__ d2,
                               int nq, int th_low, float ratio, int *__restrict__ match)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int m = -1;
    // ORBmatcher.cpp:65-67: int <= int, then float compare of bestDist against (float)bestDist2 * ratio
    if (idx1[i] >= 0 && d1[i] <= th_low && (float)d1[i] < __fmul_rn((float)d2[i], ratio)) m = idx1[i];
    match[i] = m;
}

// shard merge: re-run the :52-61 update over (d1, d2) of each shard in ascending shard order
__global__ void k_merge_shards(const int *__restrict__ d1, const int *__restrict__ idx1, const int *__restrict__ d2,
                               int nshards, int nq, size_t stride, int *__restrict__ od1, int *__restrict__ oidx1, int *__restrict__ od2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    int best = INT_MAX, best2 = INT_MAX, bidx = -1;
    for (int s = 0; s < nshards; ++s) {
        const size_t k = (size_t)s * stride + i;
        const int id = idx1[k];
        if (id < 0) continue;
        const int a = d1[k], b = d2[k];
        if (a < best) { best2 = best; best = a; bidx = id; } else if (a < best2) best2 = a;
        if (b < best2) best2 = b;     // b >= a, so it can never become 'best'
    }
    od1[i] = best; oidx1[i] = bidx; od2[i] = best2;
}

// ---- peer-memory exchange for the database-sharded search ----
// Every rank scans its shard into its own exchange buffer, then stores the call's epoch into a flag word in
// EVERY peer's buffer (NVLink P2P store).  The fused kernel below waits until all shards of this epoch are
// published, then reads the world's (d1, idx1, d2) triples straight from peer memory over NVLink and folds
// them (same rule as k_merge_shards) together with the TH_LOW / ratio acceptance: gather + merge + select in
// one kernel, no NCCL call and no staging copy on the data path.
__global__ void k_exchange_signal(uint32_t *const *__restrict__ peer_flags, int rank, int world, uint32_t epoch)
{
    const int p = threadIdx.x;
    if (p < world) {
        __threadfence_system();                                   // the triples of this epoch are written (earlier kernels, same stream)
        *reinterpret_cast<volatile uint32_t *>(peer_flags[p] + rank) = epoch;
    }
}

__global__ void __launch_bounds__(256)
k_merge_peers(const int *const *__restrict__ peer_tri, const volatile uint32_t *my_flags, int world, int nq, size_t arr_stride,
              uint32_t epoch, int *__restrict__ od1, int *__restrict__ oidx1, int *__restrict__ od2,
              int th_low, float ratio, int *__restrict__ match, int *__restrict__ err)
{
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        int ok = 1;
        const long long t0 = clock64();
        for (int r = 0; r < world && ok; ++r)
            while ((int)(my_flags[r] - epoch) < 0) {              // peer r has not published this epoch yet
                if (clock64() - t0 > 4000000000ll) { ok = 0; break; }   // ~2 s: report instead of hanging the GPU
                __nanosleep(200);
            }
        __threadfence_system();
        s_ok = ok;
        if (!ok) *err = 1;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    if (!s_ok) {                                                  // a shard never arrived: no stale results of an earlier call survive
        od1[i] = INT_MAX; oidx1[i] = -1; od2[i] = INT_MAX;
        if (match) match[i] = -1;
        return;
    }
    int best = INT_MAX, best2 = INT_MAX, bidx = -1;
    for (int r = 0; r < world; ++r) {                             // ranks own ascending index ranges
        const int *t = peer_tri[r];
        const int id = __ldcv(t + arr_stride + i);                // peer memory: bypass L1
        if (id < 0) continue;
        const int a = __ldcv(t + i), b = __ldcv(t + 2 * arr_stride + i);
        if (a < best) { best2 = best; best = a; bidx = id; } else if (a < best2) best2 = a;
        if (b < best2) best2 = b;
    }
    od1[i] = best; oidx1[i] = bidx; od2[i] = best2;
    if (match) match[i] = (bidx >= 0 && best <= th_low && (float)best < __fmul_rn((float)best2, ratio)) ? bidx : -1;
}

void launch_knn2_empty(int nq, int *d1, int *idx1, int *d2, cudaStream_t s)
{
    if (nq > 0) k_knn2_empty<<<(nq + 255) / 256, 256, 0, s>>>(nq, d1, idx1, d2);
}
void launch_exchange_signal(uint32_t *const *peer_flags, int rank, int world, uint32_t epoch, cudaStream_t s)
{
    k_exchange_signal<<<1, 32, 0, s>>>(peer_flags, rank, world, epoch);
}
void launch_merge_peers(const int *const *peer_tri, const uint32_t *my_flags, int world, int nq, size_t arr_stride, uint32_t epoch,
                        int *od1, int *oidx1, int *od2, int th_low, float ratio, int *match, int *err, cudaStream_t s)
{
    if (nq > 0) k_merge_peers<<<(nq + 255) / 256, 256, 0, s>>>(peer_tri, my_flags, world, nq, arr_stride, epoch, od1, oidx1, od2, th_low, ratio, match, err);
}

// ---- integer-pipe microbenchmarks (roofline denominators for k_knn2) ----
template <int MODE>
__global__ void __launch_bounds__(256)
k_popc_bench(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u + blockIdx.x;
    uint32_t k = seed;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = __popc(a[i]);                             // POPC only (dependent per chain, 16 chains)
            if (MODE == 1) a[i] = __popc(a[i] ^ k);                         // LOP3 + POPC
            if (MODE == 2) a[i] = (a[i] ^ k) + (a[i] >> 1);                 // ALU only (LOP3/SHF/IADD)
        }
        k = k * 1664525u + 1013904223u;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- host launchers (called from abi.cu) ----
int knn_segments(int nq, int ndb, int sm_count, int *seg_rows_out)
{
    const int gx = (nq + kQPB - 1) / kQPB;
    int target = (sm_count * 4 * 4 + gx - 1) / gx;            // ~4 waves of 4 blocks/SM
    int max_by_rows = (ndb + kTileRows - 1) / kTileRows;      // at least one tile per segment
    int nseg = target < 1 ? 1 : target;
    if (nseg > max_by_rows) nseg = max_by_rows;
    const int min_seg = (int)(((long long)ndb + (1 << kIdxBits) - 2) / ((1 << kIdxBits) - 1));
    if (nseg < min_seg) nseg = min_seg;
    if (nseg < 1) nseg = 1;
    if (nseg > 65535) nseg = 65535;
    int seg_rows = (ndb + nseg - 1) / nseg;
    seg_rows = (seg_rows + kTileRows - 1) / kTileRows * kTileRows;
    nseg = (ndb + seg_rows - 1) / seg_rows;
    *seg_rows_out = seg_rows;
    return nseg;
}

// Worst case of nseg * nq over every query count a handle created for (max_q, max_db) accepts.  The segment count GROWS
// when the query count shrinks (fewer blocks in x -> more database segments to fill the SMs), so sizing for nq = max_q
// alone rejects valid smaller calls.  nseg is non-decreasing in ndb, so ndb = max_db is the worst database.
size_t knn_partial_elems(int max_q, int max_db, int sm_count)
{
    size_t worst = 0;
    const int gx_max = (max_q + kQPB - 1) / kQPB;
    for (int gx = 1; gx <= gx_max; ++gx) {
        const int nq = gx * kQPB < max_q ? gx * kQPB : max_q;      // the largest query count with this many blocks in x
        int seg_rows = 0;
        const size_t need = (size_t)knn_segments(nq, max_db > 0 ? max_db : 1, sm_count, &seg_rows) * (size_t)nq;
        worst = need > worst ? need : worst;
    }
    return worst;
}

void launch_knn2(const uint8_t *d_query, int nq, const uint8_t *d_db, int ndb, int index_base, int nseg, int seg_rows,
                 uint2 *partial, int *d1, int *idx1, int *d2, cudaStream_t s, cudaEvent_t *ev)
{
    if (nq <= 0) return;
    if (ndb <= 0) { k_knn2_empty<<<(nq + 255) / 256, 256, 0, s>>>(nq, d1, idx1, d2); return; }
    dim3 grd((nq + kQPB - 1) / kQPB, nseg);
    if (ev) cudaEventRecord(ev[0], s);
    static int csa = -1;
    if (csa < 0) { const char *e = getenv("ORBX_KNN_CSA"); csa = e ? atoi(e) : kDefaultCsa; }
    if (csa == 0) k_knn2<0><<<grd, kKnnThreads, 0, s>>>((const uint4 *)d_query, nq, (const uint4 *)d_db, ndb, seg_rows, partial);
    else if (csa == 2) k_knn2<2><<<grd, kKnnThreads, 0, s>>>((const uint4 *)d_query, nq, (const uint4 *)d_db, ndb, seg_rows, partial);
    else if (csa == 3) k_knn2<3><<<grd, kKnnThreads, 0, s>>>((const uint4 *)d_query, nq, (const uint4 *)d_db, ndb, seg_rows, partial);
    else if (csa == 13) k_knn2<13><<<grd, kKnnThreads, 0, s>>>((const uint4 *)d_query, nq, (const uint4 *)d_db, ndb, seg_rows, partial);
    else if (csa == 15) k_knn2<15><<<grd, kKnnThreads, 0, s>>>((const uint4 *)d_query, nq, (const uint4 *)d_db, ndb, seg_rows, partial);
    else if (csa == 14) k_knn2<14><<<grd, kKnnThreads, 0, s>>>((const uint4 *)d_query, nq, (const uint4 *)d_db, ndb, seg_rows, partial);
    else k_knn2<4><<<grd, kKnnThreads, 0, s>>>((const uint4 *)d_query, nq, (const uint4 *)d_db, ndb, seg_rows, partial);
    if (ev) cudaEventRecord(ev[1], s);
    k_knn2_merge<<<(nq + 255) / 256, 256, 0, s>>>(partial, nq, nseg, seg_rows, index_base, d1, idx1, d2);
    if (ev) cudaEventRecord(ev[2], s);
}

// seg_rows / nseg_max of the batched scan depend on the row capacity only (the counts live on the device)
void knn2_pairs_geometry(int cap, int *seg_rows, int *nseg_max)
{
    *seg_rows = kTileRows;                                   // one 4 KB tile per block: the most blocks a small scan can give
    *nseg_max = (cap + kTileRows - 1) / kTileRows;
}
size_t knn2_pairs_workspace_bytes(int cap, int npairs)
{
    int sr, ns;
    knn2_pairs_geometry(cap, &sr, &ns);
    return (size_t)npairs * ns * cap * sizeof(uint2);
}
void launch_knn2_pairs(const uint8_t *desc, const int *counts, int cap, const int *pair_a, const int *pair_b, int npairs, uint2 *partial,
                       int *d1, int *idx1, int *d2, int th_low, float ratio, int *match, cudaStream_t s)
{
    if (npairs <= 0 || cap <= 0) return;
    int seg_rows, nseg_max;
    knn2_pairs_geometry(cap, &seg_rows, &nseg_max);
    dim3 grd((cap + kQPB - 1) / kQPB, nseg_max, npairs);
    k_knn2_pairs<kDefaultCsa><<<grd, kKnnThreads, 0, s>>>((const uint4 *)desc, counts, cap, pair_a, pair_b, seg_rows, nseg_max, partial);
    k_knn2_pairs_merge<<<dim3((cap + 255) / 256, npairs), 256, 0, s>>>(partial, counts, cap, pair_a, pair_b, seg_rows, nseg_max, d1, idx1, d2, th_low, ratio, match);
}

void launch_hamming_pairs(const uint8_t *a, const uint8_t *b, int n, int *dist, cudaStream_t s)
{
    if (n > 0) k_hamming_pairs<<<(n + 255) / 256, 256, 0, s>>>((const uint4 *)a, (const uint4 *)b, n, dist);
}
void launch_ratio_select(const int *d1, const int *idx1, const int *d2, int nq, int th, float ratio, int *match, cudaStream_t s)
{
    if (nq > 0) k_ratio_select<<<(nq + 255) / 256, 256, 0, s>>>(d1, idx1, d2, nq, th, ratio, match);
}
void launch_merge_shards(const int *d1, const int *idx1, const int *d2, int nshards, int nq, size_t stride, int *od1, int *oidx1, int *od2, cudaStream_t s)
{
    if (nq > 0) k_merge_shards<<<(nq + 255) / 256, 256, 0, s>>>(d1, idx1, d2, nshards, nq, stride ? stride : (size_t)nq, od1, oidx1, od2);
}

int run_popc_bench(int mode, int sm_count, double *ops_per_second)
{
    const int blocks = sm_count * 8, iters = 4096;
    uint32_t *out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * 256 * 4) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k_popc_bench<0><<<blocks, 256>>>(out, iters, 12345u + rep);
        else if (mode == 1) k_popc_bench<1><<<blocks, 256>>>(out, iters, 12345u + rep);
        else k_popc_bench<2><<<blocks, 256>>>(out, iters, 12345u + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return -1;
    *ops_per_second = (double)blocks * 256.0 * iters * 16.0 / (best * 1e-3);
    return 0;
}

} // namespace orbx

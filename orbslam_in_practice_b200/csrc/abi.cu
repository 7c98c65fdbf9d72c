// abi.cu -- the C ABI of liborbx.so (include/orbx.h): handles, geometry, launch sequencing.
// Host-side table math follows the reference constructor (ORBextractor.cpp:360-420) and
// OpenCV's resize coefficient generation; everything per-pixel runs on the device.
#include "orbx_internal.cuh"

#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: named ranges for Nsight timelines, no-ops without a tool attached

namespace {
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};
} // namespace

namespace orbx {
// knn.cu
int knn_segments(int nq, int ndb, int sm_count, int *seg_rows_out);
int octree_closed_depth(const Geo &g, int level);
size_t knn_partial_elems(int max_q, int max_db, int sm_count);
void launch_knn2_empty(int nq, int *d1, int *idx1, int *d2, cudaStream_t s);
void launch_knn2(const uint8_t *d_query, int nq, const uint8_t *d_db, int ndb, int index_base, int nseg, int seg_rows,
                 uint2 *partial, int *d1, int *idx1, int *d2, cudaStream_t s, cudaEvent_t *ev);
void launch_hamming_pairs(const uint8_t *a, const uint8_t *b, int n, int *dist, cudaStream_t s);
size_t knn2_pairs_workspace_bytes(int cap, int npairs);
void launch_knn2_pairs(const uint8_t *desc, const int *counts, int cap, const int *pair_a, const int *pair_b, int npairs, uint2 *partial,
                       int *d1, int *idx1, int *d2, int th_low, float ratio, int *match, cudaStream_t s);
void launch_ratio_select(const int *d1, const int *idx1, const int *d2, int nq, int th, float ratio, int *match, cudaStream_t s);
void launch_merge_shards(const int *d1, const int *idx1, const int *d2, int nshards, int nq, size_t stride, int *od1, int *oidx1, int *od2, cudaStream_t s);
int run_popc_bench(int mode, int sm_count, double *ops_per_second);
void launch_exchange_signal(uint32_t *const *peer_flags, int rank, int world, uint32_t epoch, cudaStream_t s);
void launch_merge_peers(const int *const *peer_tri, const uint32_t *my_flags, int world, int nq, size_t arr_stride, uint32_t epoch,
                        int *od1, int *oidx1, int *od2, int th_low, float ratio, int *match, int *err, cudaStream_t s);
} // namespace orbx

using namespace orbx;

static thread_local char g_cuda_err[256] = "";

static int cuda_fail(cudaError_t e, const char *what)
{
    std::snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
    return ORBX_E_CUDA;
}
#define CK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(e__, #call); } while (0)

struct orbx_extractor {
    orbx_params params;
    int device, max_w, max_h, max_batch;
    int cur_w, cur_h;            // geometry currently built
    int last_frames;
    Geo geo;                     // geometry of the current frame size
    Geo full;                    // geometry of (max_w, max_h): sized every buffer
    DevBuffers buf;
    std::vector<void *> allocs;
    std::vector<size_t> alloc_bytes;             // payload sizes, recorded only with guard zones
    int guard;                                   // ORBX_GUARD=1: canary zones around every device buffer (orbx_debug_guard_check)
    cudaStream_t stream, stream2, s_h2d, s_d2h;   // compute (two, alternating chunks) / upload / download
    static const int kMaxChunks = 8;
    cudaEvent_t ev_h2d[kMaxChunks], ev_done[kMaxChunks];
    cudaStream_t s_aux[2];                        // blur runs here, concurrent with FAST + octree
    cudaEvent_t ev_fork[kMaxChunks], ev_join[kMaxChunks];
    unsigned fork_slot;
    int oct_smem;
    unsigned char *oct_scratch_mem; size_t oct_scratch_bytes;   // global-memory node tables (only for quotas beyond one SM's shared memory)
    size_t tables_cap;                            // int2 entries allocated for buf.tables
    int device_split;                             // orbx_extract_device: independent sub-batches on two streams
    cudaEvent_t ev_split_fork, ev_split_join;
    long long launches;
    // reference tables
    float scale[ORBX_MAX_LEVELS], inv_scale[ORBX_MAX_LEVELS], sigma2[ORBX_MAX_LEVELS], inv_sigma2[ORBX_MAX_LEVELS];
    int nfeat[ORBX_MAX_LEVELS];
    int umax[16];
    int border_on;
    int in_channels, in_rgb;     // input pixel format (orbx_set_input_format)
    uint8_t *staging_color; size_t staging_color_bytes;   // host-path staging for colour input
    int profiling;
    static const int kProfCalls = 64;            // event sets kept (ring) while profiling
    cudaEvent_t ev[kProfCalls][ORBX_NUM_STAGES + 1];
    int prof_calls;                              // calls recorded since profiling was enabled
    // ---- low-latency path for small batches (the reference's call pattern is ONE frame per call, src/Frame.cpp:75-78) ----
    // The kernels of a call are replayed as a CUDA graph whose nodes follow the real dependencies: FAST + octree of level l
    // hang off level l of the resize chain (not off its end), the blur off the chain's end, describe joins everything.
    struct LatGraph {
        cudaGraphExec_t exec; int launches;
        const void *imgs, *kps, *desc, *counts; size_t pitch, fstride; int w, h, n, ch, rgb, border;
    };
    LatGraph lat[2];                             // [0] host entry points (handle-owned staging / outputs), [1] orbx_extract_device
    int low_latency, lat_warm;
    cudaStream_t s_cap;                          // capture origin (the graph is replayed on the call's stream)
    cudaStream_t s_branch[ORBX_MAX_LEVELS];
    cudaEvent_t ev_lvl[ORBX_MAX_LEVELS], ev_lvl_done[ORBX_MAX_LEVELS];
    uint8_t *pin_in, *pin_out; size_t pin_in_bytes, pin_out_bytes;   // pinned staging for pageable caller buffers / the single D2H
    struct { int active, n; orbx_keypoint *kps; uint8_t *desc; int32_t *counts; } lat_pending;   // copy-out owed by _end
};
static const int kLatMaxFrames = 8;

extern "C" const char *orbx_strerror(int code)
{
    switch (code) {
    case ORBX_OK: return "ok";
    case ORBX_E_INVALID: return "invalid argument";
    case ORBX_E_CUDA: return "CUDA runtime error";
    case ORBX_E_NOMEM: return "out of memory";
    case ORBX_E_CAPACITY: return "input exceeds the capacity the handle was created with";
    case ORBX_E_NODEVICE: return "no sm_100 CUDA device available (this library has no CPU fallback)";
    case ORBX_E_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown error";
    }
}
extern "C" const char *orbx_last_cuda_error(void) { return g_cuda_err; }
extern "C" int orbx_version(void) { return 100; }
#ifndef ORBX_BUILD_ID
#define ORBX_BUILD_ID "unknown"
#endif
extern "C" const char *orbx_build_id(void) { return ORBX_BUILD_ID; }

extern "C" int orbx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

static int cv_round(double v) { return (int)lrint(v); }
static int cv_floor(double v) { int i = (int)v; return i - (i > v); }
static int cv_ceil(double v) { int i = (int)v; return i + (i < v); }
static short sat_short(int v) { return (short)(v < -32768 ? -32768 : v > 32767 ? 32767 : v); }

// ORBextractor::ORBextractor, ORBextractor.cpp:360-420
static void build_reference_tables(orbx_extractor *ex)
{
    const int nlevels = ex->params.nlevels, nfeatures = ex->params.nfeatures;
    const double scaleFactor = ex->params.scale_factor;      // ORBextractor.h:84: double member
    ex->scale[0] = 1.0f; ex->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; ++i) {
        ex->scale[i] = (float)(ex->scale[i - 1] * scaleFactor);
        ex->sigma2[i] = ex->scale[i] * ex->scale[i];
    }
    for (int i = 0; i < nlevels; ++i) { ex->inv_scale[i] = 1.0f / ex->scale[i]; ex->inv_sigma2[i] = 1.0f / ex->sigma2[i]; }
    const float factor = (float)(1.0f / scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) { ex->nfeat[l] = cv_round(nDesired); sum += ex->nfeat[l]; nDesired *= factor; }
    ex->nfeat[nlevels - 1] = std::max(nfeatures - sum, 0);
    int v, v0;
    const int vmax = cv_floor(kHalfPatch * std::sqrt(2.f) / 2 + 1), vmin = cv_ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (v = 0; v <= vmax; ++v) ex->umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (ex->umax[v0] == ex->umax[v0 + 1]) ++v0;
        ex->umax[v] = v0;
        ++v0;
    }
}

// OpenCV resize INTER_LINEAR coefficient tables
static void linear_table(int ssize, int dsize, bool clamp, std::vector<int2> &tab)
{
    const double scale = 1.0 / ((double)dsize / ssize);
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor(f);
        f -= s;
        if (clamp && s < 0) { s = 0; f = 0.f; }
        if (clamp && s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        const short c0 = sat_short((int)lrintf((1.f - f) * 2048.f)), c1 = sat_short((int)lrintf(f * 2048.f));
        tab.push_back(make_int2(s, (int)((unsigned)(unsigned short)c0 | ((unsigned)(unsigned short)c1 << 16))));
    }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int host_reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// Padded tables and per-block source ranges of the staged resize kernel (pyramid.cu).  tabx / taby of level L must
// already be in `tab`.  Everything a block needs to know about its source tile is precomputed here, for both border
// widths (variant 0: kMinBlurBorder, variant 1: kBorder).
static void staged_resize_tables(const LevelGeom &S, LevelGeom &L, std::vector<int2> &tab, int max_batch)
{
    while (tab.size() & 3) tab.push_back(make_int2(0, 0));
    L.tabxp = (int)tab.size();
    for (int i = 0; i < L.pitch; ++i) {
        const int X = std::min(std::max(i - kPadX, -kBorder), L.w + kBorder - 1);
        const int2 t = tab[(size_t)L.tabx + (size_t)host_reflect101(X, L.w)];
        tab.push_back(t);
    }
    while (tab.size() & 3) tab.push_back(make_int2(0, 0));
    L.tabyp = (int)tab.size();
    for (int i = 0; i < L.h + 2 * kBorder; ++i) {
        const int2 t = tab[(size_t)L.taby + (size_t)host_reflect101(i - kBorder, L.h)];
        const int sy0 = std::min(std::max(t.x, 0), S.h - 1), sy1 = std::min(std::max(t.x + 1, 0), S.h - 1);
        tab.push_back(make_int2(sy0 | (sy1 << 16), t.y));
    }
    const int chunks = L.pitch / 4;
    L.rs_bw = chunks > 64 ? 128 : 64;                 // small levels: narrower blocks so that few lanes idle past the end of a row
    L.rs_nbx = (chunks + L.rs_bw - 1) / L.rs_bw;
    L.rs_staged = S.h < 32768 ? 1 : 0;
    // Destination rows per block.  16 amortises the staged source rows best and wins for every level of a large batch (measured on
    // 256 VGA frames: 0.174 ms with 16 rows everywhere, 0.180 with 32, 0.183 with 8 rows on levels 5-7, 0.190 / 0.235 with 8 / 4 everywhere).  The
    // handles of the low-latency path (batches of <= 8 frames, a handful of blocks per launch) are bound by the serial depth of a
    // thread instead: 4 rows per block cut the resize chain of one VGA frame from 46 to 32 us.
    {
        static int forced = -1;
        if (forced < 0) { const char *e = std::getenv("ORBX_RS_ROWS"); forced = e ? std::atoi(e) : 0; }
        L.rs_rows = max_batch <= 8 ? 4 : 16;
        if (forced == 4 || forced == 8 || forced == 16) L.rs_rows = forced;
    }
    const int kRows = L.rs_rows;
    const int Bv[2] = { kMinBlurBorder, kBorder };
    int tile_w = 16;
    L.rs_xr = (int)tab.size();
    for (int v = 0; v < 2; ++v)
        for (int bx = 0; bx < L.rs_nbx; ++bx) {
            int lo = INT_MAX, hi = -1;
            for (int c = bx * L.rs_bw; c < std::min((bx + 1) * L.rs_bw, chunks); ++c) {
                const int X0 = c * 4 - kPadX;
                if (X0 + 3 < -Bv[v] || X0 >= L.w + Bv[v]) continue;
                for (int k = 0; k < 4; ++k) {
                    const int o = tab[(size_t)L.tabxp + (size_t)(c * 4 + k)].x;
                    lo = std::min(lo, o); hi = std::max(hi, o);
                }
                // every pair of neighbouring pixels must fit an 8-byte window: source offsets at most 3 apart
                for (int k = 0; k < 4; k += 2)
                    if (std::abs(tab[(size_t)L.tabxp + (size_t)(c * 4 + k)].x - tab[(size_t)L.tabxp + (size_t)(c * 4 + k + 1)].x) > 3) L.rs_staged = 0;
            }
            if (hi >= 0) tile_w = std::max(tile_w, hi + 1 - (lo & ~15) + 8);
            tab.push_back(make_int2(hi >= 0 ? lo : 0, hi));           // hi = -1: no pixel of this block is inside the bordered level
        }
    L.rs_tile_w = (tile_w + 15) / 16 * 16;
    int tile_h = 1;
    L.rs_yr = (int)tab.size();
    for (int v = 0; v < 2; ++v) {
        const int B = Bv[v], nby = (L.h + 2 * B + kRows - 1) / kRows;
        if (v == 0) L.rs_nby0 = nby;
        for (int by = 0; by < nby; ++by) {
            int lo = INT_MAX, hi = -1;
            for (int Y = by * kRows - B; Y < std::min((by + 1) * kRows - B, L.h + B); ++Y) {
                const int e = tab[(size_t)L.tabyp + (size_t)(Y + kBorder)].x;
                lo = std::min(lo, e & 0xffff); hi = std::max(hi, e >> 16);
            }
            tile_h = std::max(tile_h, hi - lo + 1);
            tab.push_back(make_int2(lo, hi));
        }
    }
    L.rs_tile_h = tile_h;
    if ((size_t)L.rs_tile_w * (size_t)L.rs_tile_h + 1024 > 46 * 1024) L.rs_staged = 0;
}

// FAST cell descriptors of a frame geometry (ORBextractor.cpp:745-762: the cell rectangle and its skip rules), one int4
// per cell: the kernel's per-warp prologue becomes one 16-byte load instead of a level search, a division and the clamps.
static void build_cell_table(const Geo &g, std::vector<int4> &tab)
{
    tab.assign((size_t)g.total_cells, make_int4(0, 0, 0, 0));
    for (int l = 0; l < g.nlevels; ++l) {
        const LevelGeom &L = g.lv[l];
        for (int ci = 0; ci < L.nRows; ++ci)
            for (int cj = 0; cj < L.nCols; ++cj) {
                const int c = ci * L.nCols + cj;
                const int iniY = kMinBorder + ci * L.hCell, iniX = kMinBorder + cj * L.wCell;
                int cw = 0, ch = 0;
                if (!(iniY >= L.maxBorderY - 3 || iniX >= L.maxBorderX - 6)) {
                    cw = std::min(iniX + L.wCell + 6, L.maxBorderX) - iniX;
                    ch = std::min(iniY + L.hCell + 6, L.maxBorderY) - iniY;
                    if (cw - 6 <= 0 || ch - 6 <= 0) cw = ch = 0;
                }
                tab[(size_t)L.cell_base + c] = make_int4(l, iniX | (iniY << 16), cw | (ch << 16), (int)(L.slot_base + (unsigned long long)c * L.cell_cap));
            }
    }
}

static int build_geometry(const orbx_extractor *ex, int w, int h, Geo &g, std::vector<int2> *tables)
{
    std::memset(&g, 0, sizeof(g));
    g.nlevels = ex->params.nlevels; g.ini_th = ex->params.ini_th_fast; g.min_th = ex->params.min_th_fast;
    g.border_on = ex->border_on;
    for (int i = 0; i < 16; ++i) g.umax[i] = ex->umax[i];
    size_t pyr_off = 0, blur_off = 0, slot_off = 0, key_off = 0;
    int cell_off = 0, kept_off = 0;
    const size_t F = (size_t)ex->max_batch;
    for (int l = 0; l < g.nlevels; ++l) {
        LevelGeom &L = g.lv[l];
        const float s = ex->inv_scale[l];
        L.w = cv_round((float)w * s); L.h = cv_round((float)h * s);           // :1075-1076
        if (L.w < 1 || L.h < 1) return ORBX_E_UNSUPPORTED;
        L.pitch = (int)align_up((size_t)kPadX + L.w + kBorder, 64);
        L.blur_pitch = (int)align_up((size_t)L.w, 64);
        L.base = pyr_off; L.frame_stride = (size_t)L.pitch * (L.h + 2 * kPadY);
        pyr_off += L.frame_stride * F;
        L.blur_base = blur_off; L.blur_frame_stride = (size_t)L.blur_pitch * L.h;
        blur_off += L.blur_frame_stride * F;
        // cell grid :729-743
        L.maxBorderX = L.w - kBorder + 3; L.maxBorderY = L.h - kBorder + 3;
        const float width = (float)(L.maxBorderX - kMinBorder), height = (float)(L.maxBorderY - kMinBorder);
        const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
        if (width <= 0 || height <= 0 || nCols < 1 || nRows < 1) {
            L.nCols = L.nRows = 0; L.wCell = L.hCell = 1; L.cell_cap = 1;      // level too small: no cells
        } else {
            L.nCols = nCols; L.nRows = nRows;
            L.wCell = (int)std::ceil(width / nCols); L.hCell = (int)std::ceil(height / nRows);
            if (L.wCell > 64 || L.hCell > 64) return ORBX_E_UNSUPPORTED;
            L.cell_cap = ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);
        }
        L.cell_base = cell_off; cell_off += L.nCols * L.nRows;
        L.slot_base = slot_off; L.max_cand = L.nCols * L.nRows * L.cell_cap;
        slot_off += (size_t)L.max_cand;
        L.key_base = key_off; key_off += (size_t)L.max_cand;
        // octree :493-495
        L.N = ex->nfeat[l];
        L.regionW = L.maxBorderX - kMinBorder; L.regionH = L.maxBorderY - kMinBorder;
        L.nIni = (L.regionW > 0 && L.regionH > 0) ? (int)std::round((float)L.regionW / (float)L.regionH) : 0;
        L.hX = L.nIni > 0 ? (float)L.regionW / L.nIni : 1.f;
        if (L.regionW > 4095 || L.regionH > 4095) return ORBX_E_UNSUPPORTED;   // 12-bit packed coordinates
        L.kept_cap = std::max(L.N + 3, 4 * L.nIni) + 1;
        L.node_cap = L.kept_cap + 1;
        if (L.node_cap > 65535) return ORBX_E_UNSUPPORTED;
        L.kept_base = kept_off; kept_off += L.kept_cap;
        L.scale = ex->scale[l];
        L.patch_size = (int)(31 * ex->scale[l]);                                // :794
        if (l > 0 && tables) {
            while (tables->size() & 3) tables->push_back(make_int2(0, 0));   // 32-byte aligned runs for int4 loads
            L.tabx = (int)tables->size(); linear_table(g.lv[l - 1].w, L.w, true, *tables);
            L.taby = (int)tables->size(); linear_table(g.lv[l - 1].h, L.h, false, *tables);
            staged_resize_tables(g.lv[l - 1], L, *tables, ex->max_batch);
        }
    }
    g.total_cells = cell_off;
    g.oct_node_cap_max = 0;
    for (int l = 0; l < g.nlevels; ++l) g.oct_node_cap_max = std::max(g.oct_node_cap_max, g.lv[l].node_cap);
    for (int l = 0; l < g.nlevels; ++l) {
        LevelGeom &L = g.lv[l];
        L.oct_B = octree_closed_depth(g, l);
        L.inv_wCell = ((1 << 18) + L.wCell - 1) / L.wCell; L.inv_hCell = ((1 << 18) + L.hCell - 1) / L.hCell;
    }
    g.capacity = kept_off; g.kept_total = kept_off;
    g.slots_per_frame = slot_off; g.keys_per_frame = key_off;
    g.pyr_frame_total = pyr_off; g.blur_frame_total = blur_off;
    return ORBX_OK;
}

// Debug aid (ORBX_GUARD=1 at orbx_create; compute-sanitizer is not available on every pool): every device buffer of the handle
// gets a 4 KB canary zone in front of and behind it, and orbx_debug_guard_check() verifies that no kernel wrote into one.
static const size_t kGuardBytes = 4096;
static const int kGuardByte = 0xA5;
template <typename T> static int dev_alloc(orbx_extractor *ex, T **p, size_t count)
{
    void *q = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T), gz = ex->guard ? kGuardBytes : 0;
    cudaError_t e = cudaMalloc(&q, bytes + 2 * gz);
    if (e != cudaSuccess) { cuda_fail(e, "cudaMalloc"); return ORBX_E_NOMEM; }
    ex->allocs.push_back(q);
    if (gz) {
        if (cudaMemset(q, kGuardByte, gz) != cudaSuccess || cudaMemset((char *)q + gz + bytes, kGuardByte, gz) != cudaSuccess) return cuda_fail(cudaGetLastError(), "cudaMemset");
        ex->alloc_bytes.push_back(bytes);
    }
    *p = (T *)((char *)q + gz);
    return ORBX_OK;
}

extern "C" int orbx_debug_guard_check(orbx_extractor *ex, int *bad_buffer)
{
    if (!ex) return ORBX_E_INVALID;
    if (bad_buffer) *bad_buffer = -1;
    if (!ex->guard) return ORBX_E_UNSUPPORTED;          // the handle was created without ORBX_GUARD=1
    CK(cudaSetDevice(ex->device));
    CK(cudaDeviceSynchronize());
    std::vector<unsigned char> h(2 * kGuardBytes);
    for (size_t i = 0; i < ex->allocs.size() && i < ex->alloc_bytes.size(); ++i) {
        const char *q = (const char *)ex->allocs[i];
        CK(cudaMemcpy(h.data(), q, kGuardBytes, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h.data() + kGuardBytes, q + kGuardBytes + ex->alloc_bytes[i], kGuardBytes, cudaMemcpyDeviceToHost));
        for (unsigned char b : h)
            if (b != (unsigned char)kGuardByte) {
                if (bad_buffer) *bad_buffer = (int)i;
                std::snprintf(g_cuda_err, sizeof(g_cuda_err), "guard zone of device buffer %d (allocation order in orbx_create) was overwritten", (int)i);
                return ORBX_E_CUDA;
            }
    }
    return ORBX_OK;
}

static int upload_geometry(orbx_extractor *ex, int w, int h)
{
    std::vector<int2> tables;
    Geo ng;
    int rc = build_geometry(ex, w, h, ng, &tables);
    if (rc) return rc;
    // every buffer was sized by the creation geometry; a smaller frame with a wider aspect ratio can
    // still need more octree roots / keypoint slots than that
    if (ng.capacity > ex->full.capacity || ng.slots_per_frame > ex->full.slots_per_frame || ng.total_cells > ex->full.total_cells ||
        ng.keys_per_frame > ex->full.keys_per_frame || ng.pyr_frame_total > ex->full.pyr_frame_total ||
        ng.blur_frame_total > ex->full.blur_frame_total)
        return ORBX_E_CAPACITY;
    ex->geo = ng;
    ex->geo.capacity = ex->full.capacity;   // output row stride stays the creation capacity
    if (tables.size() > ex->tables_cap) return ORBX_E_CAPACITY;
    {
        std::vector<int4> cells;
        build_cell_table(ng, cells);
        if (!cells.empty()) CK(cudaMemcpyAsync(ex->buf.cell_tab, cells.data(), cells.size() * sizeof(int4), cudaMemcpyHostToDevice, ex->stream));
        CK(cudaStreamSynchronize(ex->stream));      // pageable host memory about to go out of scope
    }
    if (!tables.empty()) CK(cudaMemcpyAsync(ex->buf.tables, tables.data(), tables.size() * sizeof(int2), cudaMemcpyHostToDevice, ex->stream));
    CK(cudaStreamSynchronize(ex->stream));   // `tables` is pageable host memory about to go out of scope
    ex->cur_w = w; ex->cur_h = h;
    ex->oct_smem = octree_smem_bytes(ex->geo);
    // the octree keeps its node tables in shared memory; a level quota beyond ~2 400 features does not fit one SM and runs
    // with the tables in a global scratch buffer instead (one slice per frame and level; slow, but served)
    ex->buf.oct_scratch = nullptr; ex->buf.oct_scratch_stride = 0;
    if (ex->oct_smem > 227 * 1024) {
        const size_t stride = (size_t)octree_table_bytes(ex->geo), need = stride * (size_t)ex->max_batch * (size_t)ex->geo.nlevels;
        if (ex->oct_scratch_bytes < need) {
            if (ex->oct_scratch_mem) CK(cudaFree(ex->oct_scratch_mem));
            ex->oct_scratch_mem = nullptr; ex->oct_scratch_bytes = 0;
            if (cudaMalloc(&ex->oct_scratch_mem, need) != cudaSuccess) { cuda_fail(cudaGetLastError(), "cudaMalloc (octree node tables)"); return ORBX_E_NOMEM; }
            ex->oct_scratch_bytes = need;
        }
        ex->buf.oct_scratch = ex->oct_scratch_mem; ex->buf.oct_scratch_stride = stride;
        ex->oct_smem = 0;
    }
    if (octree_configure(ex->oct_smem)) return cuda_fail(cudaGetLastError(), "octree smem");
    return ORBX_OK;
}

extern "C" int orbx_create(const orbx_params *p, int max_width, int max_height, int max_batch, int device, orbx_extractor **out)
{
    if (!p || !out || max_width < 1 || max_height < 1 || max_batch < 1) return ORBX_E_INVALID;
    if (p->nlevels < 1 || p->nlevels > ORBX_MAX_LEVELS || p->nfeatures < 0 || !(p->scale_factor > 1.0f)) return ORBX_E_INVALID;
    if (p->ini_th_fast < 1 || p->ini_th_fast > 255 || p->min_th_fast < 1 || p->min_th_fast > 255) return ORBX_E_INVALID;
    if (max_batch > 65535) return ORBX_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return ORBX_E_NODEVICE; }
    int major = 0;
    CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return ORBX_E_NODEVICE;
    CK(cudaSetDevice(device));
    orbx_extractor *ex = new (std::nothrow) orbx_extractor();
    if (!ex) return ORBX_E_NOMEM;
    ex->params = *p; ex->device = device; ex->max_w = max_width; ex->max_h = max_height; ex->max_batch = max_batch;
    ex->guard = 0;
    if (const char *e = std::getenv("ORBX_GUARD")) ex->guard = std::atoi(e) != 0;
    ex->launches = 0; ex->last_frames = 0; ex->border_on = 0; ex->in_channels = 1; ex->in_rgb = 0;
    ex->staging_color = nullptr; ex->staging_color_bytes = 0; ex->profiling = 0; ex->prof_calls = 0;
    for (auto &set : ex->ev) for (auto &e : set) e = nullptr;
    std::memset(ex->lat, 0, sizeof(ex->lat)); ex->low_latency = 1; ex->lat_warm = 0; ex->s_cap = nullptr;
    if (const char *e = std::getenv("ORBX_LOW_LATENCY")) ex->low_latency = std::atoi(e) != 0;
    for (int i = 0; i < ORBX_MAX_LEVELS; ++i) { ex->s_branch[i] = nullptr; ex->ev_lvl[i] = nullptr; ex->ev_lvl_done[i] = nullptr; }
    ex->pin_in = ex->pin_out = nullptr; ex->pin_in_bytes = ex->pin_out_bytes = 0;
    std::memset(&ex->lat_pending, 0, sizeof(ex->lat_pending));
    std::memset(&ex->buf, 0, sizeof(ex->buf));
    ex->oct_scratch_mem = nullptr; ex->oct_scratch_bytes = 0;
    ex->stream = ex->stream2 = ex->s_h2d = ex->s_d2h = nullptr; ex->s_aux[0] = ex->s_aux[1] = nullptr; ex->fork_slot = 0;
    for (int i = 0; i < orbx_extractor::kMaxChunks; ++i) { ex->ev_h2d[i] = nullptr; ex->ev_done[i] = nullptr; ex->ev_fork[i] = nullptr; ex->ev_join[i] = nullptr; }
    build_reference_tables(ex);
    int rc = build_geometry(ex, max_width, max_height, ex->full, nullptr);
    if (rc) { delete ex; return rc; }
    ex->geo = ex->full;
    if (cudaStreamCreateWithFlags(&ex->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ex->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ex->s_aux[0], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ex->s_aux[1], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ex->s_h2d, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ex->s_d2h, cudaStreamNonBlocking) != cudaSuccess) { delete ex; return cuda_fail(cudaGetLastError(), "stream"); }
    cudaEventCreateWithFlags(&ex->ev_split_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ex->ev_split_join, cudaEventDisableTiming);
    ex->device_split = 2;
    if (const char *e = std::getenv("ORBX_DEVICE_SPLIT")) { const int v = std::atoi(e); if (v >= 1 && v <= orbx_extractor::kMaxChunks) ex->device_split = v; }
    for (int i = 0; i < orbx_extractor::kMaxChunks; ++i) {
        cudaEventCreateWithFlags(&ex->ev_h2d[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ex->ev_done[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ex->ev_fork[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ex->ev_join[i], cudaEventDisableTiming);
    }
    const Geo &g = ex->full;
    const size_t F = (size_t)max_batch;
    DevBuffers &b = ex->buf;
    size_t ntab = 0;
    for (int l = 1; l < g.nlevels; ++l) ntab += (size_t)g.lv[l].w + 2 * (size_t)g.lv[l].h + (size_t)g.lv[l].pitch + 2 * kBorder + 16 +
                                               2 * ((size_t)g.lv[l].pitch / 256 + 2) + 2 * ((size_t)(g.lv[l].h + 2 * kBorder) / 4 + 2);   // plain + padded tables + block ranges
#define TRY(x) do { rc = (x); if (rc) { orbx_destroy(ex); return rc; } } while (0)
    TRY(dev_alloc(ex, &b.pyr, g.pyr_frame_total));
    TRY(dev_alloc(ex, &b.blur, g.blur_frame_total + 256));   // the descriptor stage reads 64-byte row segments: up to 26 bytes past a row's end
    ex->tables_cap = ntab + 8 * ORBX_MAX_LEVELS;
    TRY(dev_alloc(ex, &b.tables, ex->tables_cap));
    TRY(dev_alloc(ex, &b.cell_tab, (size_t)g.total_cells));
    TRY(dev_alloc(ex, &b.cell_count, F * g.total_cells));
    TRY(dev_alloc(ex, &b.cell_slots, F * g.slots_per_frame));
    TRY(dev_alloc(ex, &b.keysA, F * g.keys_per_frame));
    TRY(dev_alloc(ex, &b.keysB, F * g.keys_per_frame));
    TRY(dev_alloc(ex, &b.nodeA, F * g.keys_per_frame));
    TRY(dev_alloc(ex, &b.nodeB, F * g.keys_per_frame));
    TRY(dev_alloc(ex, &b.scanE, F * (g.keys_per_frame + g.nlevels)));
    TRY(dev_alloc(ex, &b.ncand, F * g.nlevels));
    TRY(dev_alloc(ex, &b.kept, F * g.kept_total));
    TRY(dev_alloc(ex, &b.nkept, F * g.nlevels));
    TRY(dev_alloc(ex, &b.staging, F * (size_t)max_width * max_height));
    {   // outputs of the host path as ONE block [kps | desc | counts]: a full batch goes back in a single D2H copy
        uint8_t *blk = nullptr;
        TRY(dev_alloc(ex, &blk, F * g.capacity * (sizeof(orbx_keypoint) + 32) + F * sizeof(int)));
        b.out_kps = reinterpret_cast<orbx_keypoint *>(blk);
        b.out_desc = blk + F * g.capacity * sizeof(orbx_keypoint);
        b.out_counts = reinterpret_cast<int *>(b.out_desc + F * g.capacity * 32);
    }
#undef TRY
    rc = upload_geometry(ex, max_width, max_height);
    if (rc) { orbx_destroy(ex); return rc; }
    *out = ex;
    return ORBX_OK;
}

extern "C" int orbx_destroy(orbx_extractor *ex)
{
    if (!ex) return ORBX_OK;
    cudaSetDevice(ex->device);
    if (ex->stream) { cudaStreamSynchronize(ex->stream); cudaStreamDestroy(ex->stream); }
    if (ex->stream2) cudaStreamDestroy(ex->stream2);
    for (auto &a : ex->s_aux) if (a) { cudaStreamSynchronize(a); cudaStreamDestroy(a); }
    if (ex->s_h2d) cudaStreamDestroy(ex->s_h2d);
    if (ex->s_d2h) cudaStreamDestroy(ex->s_d2h);
    for (int i = 0; i < orbx_extractor::kMaxChunks; ++i) { if (ex->ev_h2d[i]) cudaEventDestroy(ex->ev_h2d[i]); if (ex->ev_done[i]) cudaEventDestroy(ex->ev_done[i]);
        if (ex->ev_fork[i]) cudaEventDestroy(ex->ev_fork[i]); if (ex->ev_join[i]) cudaEventDestroy(ex->ev_join[i]); }
    for (auto &lg : ex->lat) if (lg.exec) cudaGraphExecDestroy(lg.exec);
    for (int i = 0; i < ORBX_MAX_LEVELS; ++i) {
        if (ex->s_branch[i]) cudaStreamDestroy(ex->s_branch[i]);
        if (ex->ev_lvl[i]) cudaEventDestroy(ex->ev_lvl[i]);
        if (ex->ev_lvl_done[i]) cudaEventDestroy(ex->ev_lvl_done[i]);
    }
    if (ex->s_cap) cudaStreamDestroy(ex->s_cap);
    if (ex->pin_in) cudaFreeHost(ex->pin_in);
    if (ex->pin_out) cudaFreeHost(ex->pin_out);
    for (void *p : ex->allocs) cudaFree(p);
    if (ex->staging_color) cudaFree(ex->staging_color);
    if (ex->oct_scratch_mem) cudaFree(ex->oct_scratch_mem);
    if (ex->ev_split_fork) cudaEventDestroy(ex->ev_split_fork);
    if (ex->ev_split_join) cudaEventDestroy(ex->ev_split_join);
    for (auto &set : ex->ev) for (auto &e : set) if (e) cudaEventDestroy(e);
    delete ex;
    return ORBX_OK;
}

extern "C" int orbx_nlevels(const orbx_extractor *ex) { return ex ? ex->params.nlevels : ORBX_E_INVALID; }

extern "C" int orbx_capacity(const orbx_extractor *ex)
{
    if (!ex) return ORBX_E_INVALID;
    // the capacity depends on nIni (aspect ratio); report the one of the creation geometry, which
    // also sized the output buffers.  Geometry rebuilds for smaller frames never exceed it unless
    // the aspect ratio grows, which orbx_extract_* rejects with ORBX_E_CAPACITY.
    return ex->full.capacity;
}

extern "C" int orbx_tables(const orbx_extractor *ex, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2,
                           int32_t *features_per_level, int32_t *umax)
{
    if (!ex) return ORBX_E_INVALID;
    for (int i = 0; i < ex->params.nlevels; ++i) {
        if (scale) scale[i] = ex->scale[i];
        if (inv_scale) inv_scale[i] = ex->inv_scale[i];
        if (sigma2) sigma2[i] = ex->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = ex->inv_sigma2[i];
        if (features_per_level) features_per_level[i] = ex->nfeat[i];
    }
    if (umax) for (int i = 0; i < 16; ++i) umax[i] = ex->umax[i];
    return ORBX_OK;
}

extern "C" int orbx_set_input_format(orbx_extractor *ex, int channels, int rgb_order)
{
    if (!ex || (channels != 1 && channels != 3 && channels != 4)) return ORBX_E_INVALID;
    ex->in_channels = channels; ex->in_rgb = rgb_order ? 1 : 0;
    return ORBX_OK;
}

extern "C" int orbx_set_pyramid_border(orbx_extractor *ex, int enabled)
{
    if (!ex) return ORBX_E_INVALID;
    ex->border_on = enabled ? 1 : 0;
    ex->geo.border_on = ex->border_on; ex->full.border_on = ex->border_on;
    return ORBX_OK;
}

extern "C" int orbx_set_device_split(orbx_extractor *ex, int nsplit)
{
    if (!ex || nsplit < 1 || nsplit > orbx_extractor::kMaxChunks) return ORBX_E_INVALID;
    ex->device_split = nsplit;
    return ORBX_OK;
}

static int ensure_geometry(orbx_extractor *ex, int w, int h)
{
    if (w == ex->cur_w && h == ex->cur_h) return ORBX_OK;
    CK(cudaStreamSynchronize(ex->stream));
    return upload_geometry(ex, w, h);
}

// frames [frame0, frame0 + nframes) of the batch; all buffers (inputs, scratch, outputs) are indexed
// by the absolute frame number, so chunks of one batch can be in flight on different streams
static int run_pipeline(orbx_extractor *ex, const uint8_t *d_imgs, size_t pitch, size_t fstride, int w, int h, int frame0, int nframes,
                        orbx_keypoint *d_kps, uint8_t *d_desc, int *d_counts, cudaStream_t s)
{
    NvtxRange range("orbx::extract (level0, resize x7, FAST, octree, blur, describe)");
    int rc = ensure_geometry(ex, w, h);
    if (rc) return rc;
    Geo g = ex->geo;
    g.frame0 = frame0; g.in_channels = ex->in_channels; g.in_rgb = ex->in_rgb;
    const bool prof = ex->profiling != 0;
    cudaEvent_t *evs = ex->ev[ex->prof_calls % orbx_extractor::kProfCalls];
#define STAGE_EVENT(i) do { if (prof) cudaEventRecord(evs[i], s); } while (0)
    STAGE_EVENT(0);
    launch_level0(g, ex->buf, d_imgs, pitch, fstride, nframes, s);
    STAGE_EVENT(1);
    for (int l = 1; l < g.nlevels; ++l) launch_resize(g, ex->buf, l, nframes, s);
    STAGE_EVENT(2);
    if (prof) {
        // profiling: strictly serial so that every stage's event interval is that stage alone
        if (g.total_cells > 0) launch_fast(g, ex->buf, nframes, s);
        STAGE_EVENT(3);
        launch_octree(g, ex->buf, nframes, ex->oct_smem, s);
        STAGE_EVENT(4);
        launch_blur(g, ex->buf, nframes, s);
        STAGE_EVENT(5);
    } else {
        // the blur needs only the pyramid: fork it to an auxiliary stream so it fills the SMs the
        // latency-bound octree blocks and the kernel tails leave idle; join before describe
        const int slot = ex->fork_slot++ % orbx_extractor::kMaxChunks;
        cudaStream_t aux = ex->s_aux[slot & 1];
        CK(cudaEventRecord(ex->ev_fork[slot], s));
        CK(cudaStreamWaitEvent(aux, ex->ev_fork[slot], 0));
        launch_blur(g, ex->buf, nframes, aux);
        CK(cudaEventRecord(ex->ev_join[slot], aux));
        if (g.total_cells > 0) launch_fast(g, ex->buf, nframes, s);
        launch_octree(g, ex->buf, nframes, ex->oct_smem, s);
        CK(cudaStreamWaitEvent(s, ex->ev_join[slot], 0));
    }
    launch_describe(g, ex->buf, nframes, d_kps, d_desc, d_counts, s);
    STAGE_EVENT(6);
#undef STAGE_EVENT
    if (prof) ex->prof_calls++;
    ex->launches += 1 + (g.nlevels - 1) + (g.total_cells > 0) + 1 + 1 + 1;   // level0, resize chain, FAST, octree, blur, describe
    ex->last_frames = frame0 + nframes;
    CK(cudaGetLastError());
    return ORBX_OK;
}

// The same kernels as run_pipeline with the dependencies a single small batch really has: level l's FAST + octree start as
// soon as level l exists.  Issued on `s` and the handle's branch streams; meant to be stream-captured into a graph (the
// event record / wait pairs become graph edges).  Critical path of one VGA frame: level0 -> resize chain -> FAST_7 ->
// octree_7 -> describe, with the large levels' octrees (the longest kernels) running beside the chain.
static int run_pipeline_dag(orbx_extractor *ex, const uint8_t *d_imgs, size_t pitch, size_t fstride, int nframes,
                            orbx_keypoint *d_kps, uint8_t *d_desc, int *d_counts, cudaStream_t s, int *nlaunch)
{
    Geo g = ex->geo;
    g.frame0 = 0; g.in_channels = ex->in_channels; g.in_rgb = ex->in_rgb;
    int launches = 0;
    launch_level0(g, ex->buf, d_imgs, pitch, fstride, nframes, s); ++launches;
    for (int l = 0; l < g.nlevels; ++l) {
        if (l) { launch_resize(g, ex->buf, l, nframes, s); ++launches; }
        cudaStream_t br = ex->s_branch[l];
        CK(cudaEventRecord(ex->ev_lvl[l], s));
        CK(cudaStreamWaitEvent(br, ex->ev_lvl[l], 0));
        if (g.lv[l].nCols * g.lv[l].nRows > 0) { launch_fast(g, ex->buf, nframes, br, l, l + 1); ++launches; }
        launch_octree(g, ex->buf, nframes, ex->oct_smem, br, l, l + 1); ++launches;
        CK(cudaEventRecord(ex->ev_lvl_done[l], br));
    }
    launch_blur(g, ex->buf, nframes, s); ++launches;
    for (int l = 0; l < g.nlevels; ++l) CK(cudaStreamWaitEvent(s, ex->ev_lvl_done[l], 0));
    launch_describe(g, ex->buf, nframes, d_kps, d_desc, d_counts, s); ++launches;
    CK(cudaGetLastError());
    *nlaunch = launches;
    return ORBX_OK;
}

static int lat_resources(orbx_extractor *ex)
{
    for (int i = 0; i < ex->params.nlevels; ++i) {
        if (!ex->s_branch[i]) CK(cudaStreamCreateWithFlags(&ex->s_branch[i], cudaStreamNonBlocking));
        if (!ex->ev_lvl[i]) CK(cudaEventCreateWithFlags(&ex->ev_lvl[i], cudaEventDisableTiming));
        if (!ex->ev_lvl_done[i]) CK(cudaEventCreateWithFlags(&ex->ev_lvl_done[i], cudaEventDisableTiming));
    }
    if (!ex->s_cap) CK(cudaStreamCreateWithFlags(&ex->s_cap, cudaStreamNonBlocking));
    return ORBX_OK;
}

// Graph of one call's kernels for exactly these buffers and this geometry; rebuilt when anything in the key changes.
// Returns ORBX_OK with slot.exec == nullptr if graphs are unusable here (the caller then takes the stream path).
static int lat_graph_ensure(orbx_extractor *ex, int which, const uint8_t *d_imgs, size_t pitch, size_t fstride, int w, int h, int n,
                            orbx_keypoint *d_kps, uint8_t *d_desc, int *d_counts, cudaStream_t user_stream)
{
    orbx_extractor::LatGraph &lg = ex->lat[which];
    if (lg.exec && lg.imgs == d_imgs && lg.kps == d_kps && lg.desc == d_desc && lg.counts == d_counts && lg.pitch == pitch &&
        lg.fstride == fstride && lg.w == w && lg.h == h && lg.n == n && lg.ch == ex->in_channels && lg.rgb == ex->in_rgb && lg.border == ex->border_on)
        return ORBX_OK;
    if (lg.exec) { cudaGraphExecDestroy(lg.exec); lg.exec = nullptr; }
    int rc = lat_resources(ex);
    if (rc) return rc;
    if (!ex->lat_warm) {
        // one plain pass first: the launchers' one-time initialisation (table uploads, function attributes) must not
        // happen inside a capture.  It runs on the caller's stream (ordered behind whatever produces the inputs or still
        // uses the handle) and computes the same outputs the graph will overwrite.
        rc = run_pipeline(ex, d_imgs, pitch, fstride, w, h, 0, n, d_kps, d_desc, d_counts, user_stream);
        if (rc) return rc;
        CK(cudaStreamSynchronize(user_stream));
        CK(cudaStreamSynchronize(ex->s_aux[0])); CK(cudaStreamSynchronize(ex->s_aux[1]));
        ex->lat_warm = 1;
    }
    cudaGraph_t graph = nullptr;
    int nl = 0;
    if (cudaStreamBeginCapture(ex->s_cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); ex->low_latency = 0; return ORBX_OK; }
    rc = run_pipeline_dag(ex, d_imgs, pitch, fstride, n, d_kps, d_desc, d_counts, ex->s_cap, &nl);
    const cudaError_t ce = cudaStreamEndCapture(ex->s_cap, &graph);
    if (rc || ce != cudaSuccess || !graph) { cudaGetLastError(); if (graph) cudaGraphDestroy(graph); ex->low_latency = 0; return ORBX_OK; }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { cudaGetLastError(); ex->low_latency = 0; return ORBX_OK; }
    lg.exec = exec; lg.launches = nl;
    lg.imgs = d_imgs; lg.kps = d_kps; lg.desc = d_desc; lg.counts = d_counts; lg.pitch = pitch; lg.fstride = fstride;
    lg.w = w; lg.h = h; lg.n = n; lg.ch = ex->in_channels; lg.rgb = ex->in_rgb; lg.border = ex->border_on;
    return ORBX_OK;
}

extern "C" int orbx_set_low_latency(orbx_extractor *ex, int enabled)
{
    if (!ex) return ORBX_E_INVALID;
    ex->low_latency = enabled ? 1 : 0;
    return ORBX_OK;
}

extern "C" int orbx_extract_device(orbx_extractor *ex, const uint8_t *d_imgs, size_t row_pitch, size_t frame_stride,
                                   int width, int height, int nframes,
                                   orbx_keypoint *d_kps, uint8_t *d_desc, int32_t *d_counts, void *stream)
{
    if (!ex || nframes < 0 || !d_kps || !d_desc || !d_counts) return ORBX_E_INVALID;
    if (nframes == 0) return ORBX_OK;
    CK(cudaSetDevice(ex->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ex->stream;
    if (width <= 0 || height <= 0 || !d_imgs) {               // empty image: :1004-1005
        CK(cudaMemsetAsync(d_counts, 0, sizeof(int) * (size_t)nframes, s));
        return ORBX_OK;
    }
    if (width > ex->max_w || height > ex->max_h || nframes > ex->max_batch) return ORBX_E_CAPACITY;
    if (row_pitch < (size_t)width * ex->in_channels) return ORBX_E_INVALID;
    // Large batches run as two independent halves on two streams: frames are independent, every kernel here is
    // instruction-issue bound at 65-77 % issue utilisation, and the halves' different kernels (and their tail waves)
    // fill each other's idle issue slots.  Profiling keeps one serial pass so that stage times stay attributable.
    int nsplit = (!ex->profiling && nframes >= 32) ? ex->device_split : 1;
    if (ex->low_latency && !ex->profiling && nframes <= kLatMaxFrames) {
        int rc = ensure_geometry(ex, width, height);
        if (rc) return rc;
        rc = lat_graph_ensure(ex, 1, d_imgs, row_pitch, frame_stride, width, height, nframes, d_kps, d_desc, d_counts, s);
        if (rc) return rc;
        if (ex->lat[1].exec) {
            CK(cudaGraphLaunch(ex->lat[1].exec, s));
            ex->launches += ex->lat[1].launches; ex->last_frames = nframes;
            return ORBX_OK;
        }
    }
    if (nsplit <= 1) return run_pipeline(ex, d_imgs, row_pitch, frame_stride, width, height, 0, nframes, d_kps, d_desc, d_counts, s);
    CK(cudaEventRecord(ex->ev_split_fork, s));
    CK(cudaStreamWaitEvent(ex->stream2, ex->ev_split_fork, 0));
    for (int c = 0; c < nsplit; ++c) {
        const int f0 = (int)((long long)nframes * c / nsplit), f1 = (int)((long long)nframes * (c + 1) / nsplit);
        int rc = run_pipeline(ex, d_imgs, row_pitch, frame_stride, width, height, f0, f1 - f0, d_kps, d_desc, d_counts, (c & 1) ? ex->stream2 : s);
        if (rc) return rc;
    }
    CK(cudaEventRecord(ex->ev_split_join, ex->stream2));
    CK(cudaStreamWaitEvent(s, ex->ev_split_join, 0));
    return ORBX_OK;
}

extern "C" int orbx_extract_host_end(orbx_extractor *ex)
{
    if (!ex) return ORBX_E_INVALID;
    CK(cudaSetDevice(ex->device));
    if (ex->lat_pending.active) {
        // low-latency path: everything ran on ex->stream; hand the caller exactly counts[f] rows per frame
        ex->lat_pending.active = 0;
        CK(cudaStreamSynchronize(ex->stream));
        const size_t cap = (size_t)ex->full.capacity; const int n = ex->lat_pending.n;
        const orbx_keypoint *pk = reinterpret_cast<const orbx_keypoint *>(ex->pin_out);
        const uint8_t *pd = ex->pin_out + (size_t)n * cap * sizeof(orbx_keypoint);
        const int *pc = reinterpret_cast<const int *>(pd + (size_t)n * cap * 32);
        for (int f = 0; f < n; ++f) {
            int c = pc[f]; if (c < 0) c = 0; if ((size_t)c > cap) c = (int)cap;
            ex->lat_pending.counts[f] = c;
            std::memcpy(ex->lat_pending.kps + f * cap, pk + f * cap, (size_t)c * sizeof(orbx_keypoint));
            std::memcpy(ex->lat_pending.desc + f * cap * 32, pd + f * cap * 32, (size_t)c * 32);
        }
        return ORBX_OK;
    }
    CK(cudaStreamSynchronize(ex->s_d2h));
    CK(cudaStreamSynchronize(ex->stream));
    CK(cudaStreamSynchronize(ex->stream2));
    return ORBX_OK;
}

static bool host_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int pin_ensure(uint8_t **buf, size_t *have, size_t need)
{
    if (*have >= need) return ORBX_OK;
    if (*buf) { cudaFreeHost(*buf); *buf = nullptr; *have = 0; }
    if (cudaHostAlloc((void **)buf, need, cudaHostAllocDefault) != cudaSuccess) { cuda_fail(cudaGetLastError(), "cudaHostAlloc"); return ORBX_E_NOMEM; }
    *have = need;
    return ORBX_OK;
}

// Small batches (<= kLatMaxFrames frames; the reference's caller passes ONE): a single stream, the upload, the kernel graph
// and ONE download into handle-owned pinned memory.  Pageable caller images are first copied into pinned staging (a
// cudaMemcpyAsync from pageable memory stages through the driver anyway, and synchronously).  Measured and rejected:
// describe storing its rows straight into pinned host memory (no download node) -- 5 us slower for one frame, 40 % for eight.
static int extract_host_small(orbx_extractor *ex, const uint8_t *imgs, size_t row_pitch, size_t frame_stride, int width, int height, int n,
                              uint8_t *staging, orbx_keypoint *kps, uint8_t *desc, int32_t *counts, bool *taken)
{
    *taken = false;
    const size_t C = (size_t)ex->in_channels, rbytes = (size_t)width * C, fbytes = rbytes * height, cap = (size_t)ex->full.capacity;
    int rc = lat_graph_ensure(ex, 0, staging, rbytes, fbytes, width, height, n, ex->buf.out_kps, ex->buf.out_desc, ex->buf.out_counts, ex->stream);
    if (rc) return rc;
    if (!ex->lat[0].exec) return ORBX_OK;           // graphs unavailable: the caller falls through to the stream path
    cudaStream_t s = ex->stream;
    const uint8_t *src = imgs; size_t sp = row_pitch, sf = frame_stride;
    if (!host_pinned(imgs)) {
        rc = pin_ensure(&ex->pin_in, &ex->pin_in_bytes, (size_t)n * fbytes);
        if (rc) return rc;
        for (int f = 0; f < n; ++f) {
            const uint8_t *fs = imgs + (size_t)f * frame_stride; uint8_t *fd = ex->pin_in + (size_t)f * fbytes;
            if (row_pitch == rbytes) std::memcpy(fd, fs, fbytes);
            else for (int y = 0; y < height; ++y) std::memcpy(fd + (size_t)y * rbytes, fs + (size_t)y * row_pitch, rbytes);
        }
        src = ex->pin_in; sp = rbytes; sf = fbytes;
    }
    if (sp == rbytes && (sf == fbytes || n == 1)) CK(cudaMemcpyAsync(staging, src, (size_t)n * fbytes, cudaMemcpyHostToDevice, s));
    else for (int f = 0; f < n; ++f)
        CK(cudaMemcpy2DAsync(staging + (size_t)f * fbytes, rbytes, src + (size_t)f * sf, sp, rbytes, (size_t)height, cudaMemcpyHostToDevice, s));
    CK(cudaGraphLaunch(ex->lat[0].exec, s));
    ex->launches += ex->lat[0].launches; ex->last_frames = n;
    const size_t kb = (size_t)n * cap * sizeof(orbx_keypoint), db = (size_t)n * cap * 32, cb = (size_t)n * sizeof(int);
    rc = pin_ensure(&ex->pin_out, &ex->pin_out_bytes, kb + db + cb);
    if (rc) return rc;
    if (n == ex->max_batch) {
        CK(cudaMemcpyAsync(ex->pin_out, ex->buf.out_kps, kb + db + cb, cudaMemcpyDeviceToHost, s));    // the block is contiguous for a full batch
    } else {
        CK(cudaMemcpyAsync(ex->pin_out, ex->buf.out_kps, kb, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ex->pin_out + kb, ex->buf.out_desc, db, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(ex->pin_out + kb + db, ex->buf.out_counts, cb, cudaMemcpyDeviceToHost, s));
    }
    ex->lat_pending.active = 1; ex->lat_pending.n = n; ex->lat_pending.kps = kps; ex->lat_pending.desc = desc; ex->lat_pending.counts = counts;
    *taken = true;
    return ORBX_OK;
}

extern "C" int orbx_extract_host(orbx_extractor *ex, const uint8_t *imgs, size_t row_pitch, size_t frame_stride,
                                 int width, int height, int nframes,
                                 orbx_keypoint *kps, uint8_t *desc, int32_t *counts)
{
    int rc = orbx_extract_host_begin(ex, imgs, row_pitch, frame_stride, width, height, nframes, kps, desc, counts);
    if (rc) return rc;
    return orbx_extract_host_end(ex);
}

extern "C" int orbx_extract_host_begin(orbx_extractor *ex, const uint8_t *imgs, size_t row_pitch, size_t frame_stride,
                                       int width, int height, int nframes,
                                       orbx_keypoint *kps, uint8_t *desc, int32_t *counts)
{
    if (!ex || nframes < 0 || !counts) return ORBX_E_INVALID;
    if (nframes == 0) return ORBX_OK;
    if (width <= 0 || height <= 0 || !imgs) { std::memset(counts, 0, sizeof(int) * (size_t)nframes); return ORBX_OK; }
    if (!kps || !desc) return ORBX_E_INVALID;
    if (width > ex->max_w || height > ex->max_h || nframes > ex->max_batch) return ORBX_E_CAPACITY;
    const size_t C = (size_t)ex->in_channels;
    if (row_pitch < (size_t)width * C) return ORBX_E_INVALID;
    CK(cudaSetDevice(ex->device));
    const size_t cap = (size_t)ex->full.capacity;
    uint8_t *staging = ex->buf.staging;
    if (C != 1) {                                   // colour input needs C x the gray staging: allocated on first use
        const size_t need = (size_t)ex->max_batch * ex->max_w * ex->max_h * C;
        if (ex->staging_color_bytes < need) {
            if (ex->staging_color) CK(cudaFree(ex->staging_color));
            ex->staging_color = nullptr; ex->staging_color_bytes = 0;
            if (cudaMalloc(&ex->staging_color, need) != cudaSuccess) { cuda_fail(cudaGetLastError(), "cudaMalloc"); return ORBX_E_NOMEM; }
            ex->staging_color_bytes = need;
        }
        staging = ex->staging_color;
    }
    int rc = ensure_geometry(ex, width, height);
    if (rc) return rc;
    if (ex->low_latency && !ex->profiling && nframes <= kLatMaxFrames) {
        bool taken = false;
        rc = extract_host_small(ex, imgs, row_pitch, frame_stride, width, height, nframes, staging, kps, desc, counts, &taken);
        if (rc || taken) return rc;
    }
    // Chunked three-stream pipeline: the H2D copy of chunk c+1 and the D2H copy of chunk c-1 overlap
    // the kernels of chunk c (PCIe is full duplex; frames are independent).
    // (256 VGA frames, blocking call, final kernels of round 2: 3 / 4 / 5 / 6 / 8 chunks = 127 / 134 / 138 / 140 / 143 k frames/s: the tail behind the last
    // upload is one chunk's kernels and download)
    int nchunks = nframes >= 256 ? 8 : nframes >= 64 ? 4 : (nframes >= 8 ? 2 : 1);
    if (const char *e = std::getenv("ORBX_HOST_CHUNKS")) { const int v = std::atoi(e); if (v >= 1 && v <= orbx_extractor::kMaxChunks && v <= nframes) nchunks = v; }
    const size_t rbytes = (size_t)width * C;           // bytes per tight row
    const size_t fbytes = rbytes * height;
    for (int c = 0; c < nchunks; ++c) {
        const int f0 = (int)((long long)nframes * c / nchunks), f1 = (int)((long long)nframes * (c + 1) / nchunks);
        const int n = f1 - f0;
        if (n <= 0) continue;
        if (row_pitch == rbytes && (frame_stride == fbytes || n == 1)) {
            // fully contiguous frames: one linear copy (a 2-D copy of 640-byte rows is slower on the DMA engine)
            CK(cudaMemcpyAsync(staging + f0 * fbytes, imgs + (size_t)f0 * frame_stride, (size_t)n * fbytes,
                               cudaMemcpyHostToDevice, ex->s_h2d));
        } else if (frame_stride == row_pitch * (size_t)height || n == 1) {
            CK(cudaMemcpy2DAsync(staging + f0 * fbytes, rbytes, imgs + (size_t)f0 * frame_stride, row_pitch,
                                 rbytes, (size_t)height * n, cudaMemcpyHostToDevice, ex->s_h2d));
        } else {
            for (int f = f0; f < f1; ++f)
                CK(cudaMemcpy2DAsync(staging + f * fbytes, rbytes, imgs + (size_t)f * frame_stride, row_pitch,
                                     rbytes, (size_t)height, cudaMemcpyHostToDevice, ex->s_h2d));
        }
        CK(cudaEventRecord(ex->ev_h2d[c], ex->s_h2d));
        // chunks alternate between two compute streams so one chunk's kernels fill the launch gaps
        // and tail waves of the other (per-frame buffers are disjoint)
        cudaStream_t cs = (c & 1) ? ex->stream2 : ex->stream;
        CK(cudaStreamWaitEvent(cs, ex->ev_h2d[c], 0));
        rc = run_pipeline(ex, staging, rbytes, fbytes, width, height, f0, n,
                          ex->buf.out_kps, ex->buf.out_desc, ex->buf.out_counts, cs);
        if (rc) return rc;
        CK(cudaEventRecord(ex->ev_done[c], cs));
        CK(cudaStreamWaitEvent(ex->s_d2h, ex->ev_done[c], 0));
        CK(cudaMemcpyAsync(kps + f0 * cap, ex->buf.out_kps + f0 * cap, (size_t)n * cap * sizeof(orbx_keypoint), cudaMemcpyDeviceToHost, ex->s_d2h));
        CK(cudaMemcpyAsync(desc + f0 * cap * 32, ex->buf.out_desc + f0 * cap * 32, (size_t)n * cap * 32, cudaMemcpyDeviceToHost, ex->s_d2h));
        CK(cudaMemcpyAsync(counts + f0, ex->buf.out_counts + f0, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ex->s_d2h));
    }
    return ORBX_OK;                                  // copies and kernels are in flight: orbx_extract_host_end waits
}

extern "C" int orbx_level_dims(const orbx_extractor *ex, int level, int *width, int *height)
{
    if (!ex || level < 0 || level >= ex->params.nlevels || !width || !height) return ORBX_E_INVALID;
    *width = ex->geo.lv[level].w; *height = ex->geo.lv[level].h;
    return ORBX_OK;
}

extern "C" int orbx_download_level(orbx_extractor *ex, int frame, int level, int blurred, int border, uint8_t *dst, size_t dst_pitch)
{
    if (!ex || !dst || level < 0 || level >= ex->params.nlevels || frame < 0 || frame >= ex->last_frames) return ORBX_E_INVALID;
    if (border != 0 && (border != kBorder || blurred)) return ORBX_E_INVALID;
    CK(cudaSetDevice(ex->device));
    const LevelGeom &L = ex->geo.lv[level];
    const size_t w = (size_t)L.w + 2 * border, h = (size_t)L.h + 2 * border;
    if (dst_pitch < w) return ORBX_E_INVALID;
    CK(cudaStreamSynchronize(ex->stream));
    if (blurred) {
        CK(cudaMemcpy2D(dst, dst_pitch, ex->buf.blur + L.blur_base + (size_t)frame * L.blur_frame_stride, L.blur_pitch, w, h, cudaMemcpyDeviceToHost));
    } else {
        // lazy mode (default): the hot path writes only the 4-px border its kernels read; the full 19-px
        // reflect-101 border of mvImagePyramid is materialised here, when somebody actually asks for it
        if (border && !ex->border_on) {
            launch_fill_border(ex->geo, ex->buf, level, frame, ex->stream);
            ex->launches += 1;
            CK(cudaStreamSynchronize(ex->stream));
        }
        const uint8_t *src = ex->buf.pyr + L.base + (size_t)frame * L.frame_stride + (size_t)(kPadY - border) * L.pitch + kPadX - border;
        CK(cudaMemcpy2D(dst, dst_pitch, src, L.pitch, w, h, cudaMemcpyDeviceToHost));
    }
    return ORBX_OK;
}

extern "C" int orbx_level_device_ptr(orbx_extractor *ex, int frame, int level, const uint8_t **ptr, size_t *pitch)
{
    if (!ex || !ptr || !pitch || level < 0 || level >= ex->params.nlevels || frame < 0 || frame >= ex->max_batch) return ORBX_E_INVALID;
    const LevelGeom &L = ex->geo.lv[level];
    *ptr = ex->buf.pyr + L.base + (size_t)frame * L.frame_stride + (size_t)kPadY * L.pitch + kPadX;
    *pitch = (size_t)L.pitch;
    return ORBX_OK;
}

extern "C" int orbx_max_candidates(const orbx_extractor *ex, int level)
{
    if (!ex || level < 0 || level >= ex->params.nlevels) return ORBX_E_INVALID;
    return ex->geo.lv[level].max_cand;
}

static int download_keys(orbx_extractor *ex, const uint32_t *d_src, const int *d_count, int count_cap, orbx_cand *out, int cap, int *n)
{
    int cnt = 0;
    CK(cudaStreamSynchronize(ex->stream));
    CK(cudaMemcpy(&cnt, d_count, sizeof(int), cudaMemcpyDeviceToHost));
    if (cnt > count_cap) cnt = count_cap;
    *n = cnt;
    if (cnt > cap) return ORBX_E_CAPACITY;
    std::vector<uint32_t> tmp((size_t)cnt);
    if (cnt) CK(cudaMemcpy(tmp.data(), d_src, sizeof(uint32_t) * (size_t)cnt, cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt; ++i) { out[i].x = (int16_t)cand_x(tmp[i]); out[i].y = (int16_t)cand_y(tmp[i]); out[i].score = cand_score(tmp[i]); }
    return ORBX_OK;
}

extern "C" int orbx_download_candidates(orbx_extractor *ex, int frame, int level, orbx_cand *out, int cap, int *n)
{
    if (!ex || !out || !n || level < 0 || level >= ex->params.nlevels || frame < 0 || frame >= ex->last_frames) return ORBX_E_INVALID;
    CK(cudaSetDevice(ex->device));
    const Geo &g = ex->geo; const LevelGeom &L = g.lv[level];
    // after the octree kernel the ordered candidate list is no longer contiguous (keys are permuted
    // inside node ranges); rebuild it from the per-cell slots, which are left untouched
    CK(cudaStreamSynchronize(ex->stream));
    const int nCells = L.nCols * L.nRows;
    std::vector<int> counts((size_t)nCells);
    if (nCells) CK(cudaMemcpy(counts.data(), ex->buf.cell_count + (size_t)frame * g.total_cells + L.cell_base, sizeof(int) * (size_t)nCells, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> slots((size_t)L.max_cand);
    if (L.max_cand) CK(cudaMemcpy(slots.data(), ex->buf.cell_slots + (size_t)frame * g.slots_per_frame + L.slot_base, sizeof(uint32_t) * (size_t)L.max_cand, cudaMemcpyDeviceToHost));
    int total = 0;
    for (int c = 0; c < nCells; ++c) total += counts[c];
    *n = total;
    if (total > cap) return ORBX_E_CAPACITY;
    int k = 0;
    for (int c = 0; c < nCells; ++c)
        for (int i = 0; i < counts[c]; ++i) {
            const uint32_t key = slots[(size_t)c * L.cell_cap + i];
            out[k].x = (int16_t)cand_x(key); out[k].y = (int16_t)cand_y(key); out[k].score = cand_score(key); ++k;
        }
    return ORBX_OK;
}

extern "C" int orbx_download_kept(orbx_extractor *ex, int frame, int level, orbx_cand *out, int cap, int *n)
{
    if (!ex || !out || !n || level < 0 || level >= ex->params.nlevels || frame < 0 || frame >= ex->last_frames) return ORBX_E_INVALID;
    CK(cudaSetDevice(ex->device));
    const Geo &g = ex->geo; const LevelGeom &L = g.lv[level];
    return download_keys(ex, ex->buf.kept + (size_t)frame * g.kept_total + L.kept_base, ex->buf.nkept + (size_t)frame * g.nlevels + level, L.kept_cap, out, cap, n);
}

extern "C" int orbx_undistort_keypoints_device(orbx_extractor *ex, const orbx_keypoint *d_in, orbx_keypoint *d_out, int n,
                                               const float *cam, const float *dist, int literal_bug, void *stream)
{
    if (!ex || n < 0 || !cam || !dist || (n && (!d_in || !d_out))) return ORBX_E_INVALID;
    CK(cudaSetDevice(ex->device));
    launch_undistort(d_in, d_out, n, cam, dist, literal_bug, stream ? (cudaStream_t)stream : ex->stream);
    ex->launches += n > 0;
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" long long orbx_launch_count(const orbx_extractor *ex) { return ex ? ex->launches : 0; }

extern "C" int orbx_set_profiling(orbx_extractor *ex, int enabled)
{
    if (!ex) return ORBX_E_INVALID;
    CK(cudaSetDevice(ex->device));
    if (enabled) {
        for (auto &set : ex->ev) for (auto &e : set) if (!e) CK(cudaEventCreate(&e));
        ex->prof_calls = 0;
    }
    ex->profiling = enabled ? 1 : 0;
    return ORBX_OK;
}

extern "C" int orbx_stage_times(orbx_extractor *ex, float *ms)
{
    if (!ex || !ms || ex->prof_calls < 1) return ORBX_E_INVALID;
    CK(cudaSetDevice(ex->device));
    // average over the calls recorded since profiling was enabled (at most the last kProfCalls)
    const int n = ex->prof_calls < orbx_extractor::kProfCalls ? ex->prof_calls : orbx_extractor::kProfCalls;
    for (int i = 0; i < ORBX_NUM_STAGES; ++i) ms[i] = 0.f;
    for (int c = 0; c < n; ++c) {
        CK(cudaEventSynchronize(ex->ev[c][ORBX_NUM_STAGES]));
        for (int i = 0; i < ORBX_NUM_STAGES; ++i) {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, ex->ev[c][i], ex->ev[c][i + 1]));
            ms[i] += t / n;
        }
    }
    return ORBX_OK;
}

// ------------------------------------------------------------------------------------------------
// matcher
// ------------------------------------------------------------------------------------------------
struct orbm_matcher {
    int device, max_q, max_db, sm_count;
    cudaStream_t stream;
    uint2 *partial; size_t partial_elems;
    uint8_t *d_q, *d_db; int *d_out;   // device staging for the host-pointer entry points
    void *si_buf; size_t si_bytes;     // scratch of orbm_search_init_host
    // peer-memory exchange (sharded search): own buffer = [2 parities][3][xq_max] int32 + flags[64] + err
    unsigned char *x_buf; int x_qmax, x_rank, x_world; uint32_t x_epoch;
    void *x_peer[64];                  // opened peer mappings (own entry = x_buf)
    int **x_tri_tab; uint32_t **x_flag_tab;   // device tables: [2][world] triple pointers, [world] flag pointers
    long long launches;
    int profiling; bool ev_valid;
    cudaEvent_t ev[3];
};

extern "C" int orbm_create(int max_queries, int max_db, int device, orbm_matcher **out)
{
    if (!out || max_queries < 1 || max_db < 0) return ORBX_E_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return ORBX_E_NODEVICE; }
    int major = 0;
    CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) return ORBX_E_NODEVICE;
    CK(cudaSetDevice(device));
    orbm_matcher *m = new (std::nothrow) orbm_matcher();
    if (!m) return ORBX_E_NOMEM;
    std::memset(m, 0, sizeof(*m));
    m->device = device; m->max_q = max_queries; m->max_db = max_db;
    CK(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device));
    CK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    m->partial_elems = knn_partial_elems(max_queries, max_db, m->sm_count);   // worst case over every nq <= max_queries
    if (cudaMalloc(&m->partial, m->partial_elems * sizeof(uint2)) != cudaSuccess) { cuda_fail(cudaGetLastError(), "cudaMalloc"); orbm_destroy(m); return ORBX_E_NOMEM; }
    *out = m;
    return ORBX_OK;
}

extern "C" int orbm_destroy(orbm_matcher *m)
{
    if (!m) return ORBX_OK;
    cudaSetDevice(m->device);
    if (m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
    cudaFree(m->partial); cudaFree(m->d_q); cudaFree(m->d_db); cudaFree(m->d_out); cudaFree(m->si_buf);
    for (int r = 0; r < m->x_world; ++r) if (m->x_peer[r] && r != m->x_rank) cudaIpcCloseMemHandle(m->x_peer[r]);
    cudaFree(m->x_buf); cudaFree(m->x_tri_tab); cudaFree(m->x_flag_tab);
    for (auto &e : m->ev) if (e) cudaEventDestroy(e);
    delete m;
    return ORBX_OK;
}

extern "C" long long orbm_launch_count(const orbm_matcher *m) { return m ? m->launches : 0; }

extern "C" int orbm_set_profiling(orbm_matcher *m, int enabled)
{
    if (!m) return ORBX_E_INVALID;
    CK(cudaSetDevice(m->device));
    if (enabled) for (auto &e : m->ev) if (!e) CK(cudaEventCreate(&e));
    m->profiling = enabled ? 1 : 0; m->ev_valid = false;
    return ORBX_OK;
}

extern "C" int orbm_knn2_times(orbm_matcher *m, float *scan_ms, float *merge_ms)
{
    if (!m || !m->ev_valid) return ORBX_E_INVALID;
    CK(cudaSetDevice(m->device));
    CK(cudaEventSynchronize(m->ev[2]));
    if (scan_ms) CK(cudaEventElapsedTime(scan_ms, m->ev[0], m->ev[1]));
    if (merge_ms) CK(cudaEventElapsedTime(merge_ms, m->ev[1], m->ev[2]));
    return ORBX_OK;
}

extern "C" int orbm_knn2_device(orbm_matcher *m, const uint8_t *d_query, int nq, const uint8_t *d_db, int ndb,
                                int index_base, int32_t *d_d1, int32_t *d_idx1, int32_t *d_d2, void *stream)
{
    if (!m || nq < 0 || ndb < 0 || (nq && (!d_query || !d_d1 || !d_idx1 || !d_d2)) || (ndb && !d_db)) return ORBX_E_INVALID;
    if (nq > m->max_q || ndb > m->max_db) return ORBX_E_CAPACITY;
    if (((uintptr_t)d_query | (uintptr_t)d_db) & 15) return ORBX_E_INVALID;   // 16-byte vector loads
    if (nq == 0) return ORBX_OK;
    NvtxRange range("orbm::knn2 (scan + segment merge)");
    CK(cudaSetDevice(m->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : m->stream;
    int seg_rows = 0;
    const int nseg = ndb > 0 ? knn_segments(nq, ndb, m->sm_count, &seg_rows) : 0;
    if ((size_t)nseg * nq > m->partial_elems) return ORBX_E_CAPACITY;
    launch_knn2(d_query, nq, d_db, ndb, index_base, nseg, seg_rows, m->partial, d_d1, d_idx1, d_d2, s, m->profiling ? m->ev : nullptr);
    m->ev_valid = m->profiling && ndb > 0;
    m->launches += ndb > 0 ? 2 : 1;
    CK(cudaGetLastError());
    return ORBX_OK;
}

static int ensure_staging(orbm_matcher *m)
{
    if (m->d_out) return ORBX_OK;
    CK(cudaMalloc(&m->d_q, (size_t)m->max_q * 32));
    CK(cudaMalloc(&m->d_db, (size_t)std::max(std::max(m->max_db, m->max_q), 1) * 32));
    CK(cudaMalloc(&m->d_out, (size_t)m->max_q * 3 * sizeof(int)));
    return ORBX_OK;
}

extern "C" int orbm_knn2_host(orbm_matcher *m, const uint8_t *query, int nq, const uint8_t *db, int ndb,
                              int index_base, int32_t *d1, int32_t *idx1, int32_t *d2)
{
    if (!m || nq < 0 || ndb < 0 || (nq && (!query || !d1 || !idx1 || !d2)) || (ndb && !db)) return ORBX_E_INVALID;
    if (nq > m->max_q || ndb > m->max_db) return ORBX_E_CAPACITY;
    if (nq == 0) return ORBX_OK;
    CK(cudaSetDevice(m->device));
    int rc = ensure_staging(m);
    if (rc) return rc;
    cudaStream_t s = m->stream;
    CK(cudaMemcpyAsync(m->d_q, query, (size_t)nq * 32, cudaMemcpyHostToDevice, s));
    if (ndb) CK(cudaMemcpyAsync(m->d_db, db, (size_t)ndb * 32, cudaMemcpyHostToDevice, s));
    int *o = m->d_out;
    rc = orbm_knn2_device(m, m->d_q, nq, m->d_db, ndb, index_base, o, o + m->max_q, o + 2 * (size_t)m->max_q, s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(d1, o, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(idx1, o + m->max_q, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(d2, o + 2 * (size_t)m->max_q, (size_t)nq * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" int orbm_hamming_pairs_host(orbm_matcher *m, const uint8_t *a, const uint8_t *b, int n, int32_t *dist)
{
    if (!m || n < 0 || (n && (!a || !b || !dist))) return ORBX_E_INVALID;
    if (n > m->max_q) return ORBX_E_CAPACITY;
    if (n == 0) return ORBX_OK;
    CK(cudaSetDevice(m->device));
    int rc = ensure_staging(m);
    if (rc) return rc;
    cudaStream_t s = m->stream;
    CK(cudaMemcpyAsync(m->d_q, a, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(m->d_db, b, (size_t)n * 32, cudaMemcpyHostToDevice, s));
    launch_hamming_pairs(m->d_q, m->d_db, n, m->d_out, s);
    m->launches += 1;
    CK(cudaMemcpyAsync(dist, m->d_out, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" int orbm_ratio_select_device(orbm_matcher *m, const int32_t *d_d1, const int32_t *d_idx1, const int32_t *d_d2,
                                        int nq, int th_low, float ratio, int32_t *d_match, void *stream)
{
    if (!m || nq < 0 || (nq && (!d_d1 || !d_idx1 || !d_d2 || !d_match))) return ORBX_E_INVALID;
    CK(cudaSetDevice(m->device));
    launch_ratio_select(d_d1, d_idx1, d_d2, nq, th_low, ratio, d_match, stream ? (cudaStream_t)stream : m->stream);
    m->launches += nq > 0;
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbm_merge_shards_device(orbm_matcher *m, const int32_t *d_d1, const int32_t *d_idx1, const int32_t *d_d2,
                                        int nshards, int nq, size_t shard_stride,
                                        int32_t *d_od1, int32_t *d_oidx1, int32_t *d_od2, void *stream)
{
    if (!m || nq < 0 || nshards < 1 || (nq && (!d_d1 || !d_idx1 || !d_d2 || !d_od1 || !d_oidx1 || !d_od2))) return ORBX_E_INVALID;
    CK(cudaSetDevice(m->device));
    launch_merge_shards(d_d1, d_idx1, d_d2, nshards, nq, shard_stride, d_od1, d_oidx1, d_od2, stream ? (cudaStream_t)stream : m->stream);
    m->launches += nq > 0;
    CK(cudaGetLastError());
    return ORBX_OK;
}

static size_t x_tri_bytes(const orbm_matcher *m) { return (size_t)2 * 3 * m->x_qmax * sizeof(int); }

extern "C" int orbm_exchange_create(orbm_matcher *m, int max_queries, int rank, int world, unsigned char *handle_out)
{
    if (!m || !handle_out || max_queries < 1 || world < 1 || world > 64 || rank < 0 || rank >= world || m->x_buf) return ORBX_E_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == ORBM_IPC_HANDLE_BYTES, "IPC handle size");
    CK(cudaSetDevice(m->device));
    m->x_qmax = max_queries; m->x_rank = rank; m->x_world = world; m->x_epoch = 0;
    const size_t bytes = x_tri_bytes(m) + 64 * sizeof(uint32_t) + 64;
    if (cudaMalloc(&m->x_buf, bytes) != cudaSuccess) { cuda_fail(cudaGetLastError(), "cudaMalloc"); return ORBX_E_NOMEM; }
    CK(cudaMemset(m->x_buf, 0, bytes));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, m->x_buf));
    std::memcpy(handle_out, &h, sizeof(h));
    return ORBX_OK;
}

extern "C" int orbm_exchange_open(orbm_matcher *m, const unsigned char *handles)
{
    if (!m || !handles || !m->x_buf || m->x_tri_tab) return ORBX_E_INVALID;
    CK(cudaSetDevice(m->device));
    const int W = m->x_world;
    for (int r = 0; r < W; ++r) {
        if (r == m->x_rank) { m->x_peer[r] = m->x_buf; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)r * ORBM_IPC_HANDLE_BYTES, sizeof(h));
        CK(cudaIpcOpenMemHandle(&m->x_peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    std::vector<int *> tri((size_t)2 * W);
    std::vector<uint32_t *> flg((size_t)W);
    for (int r = 0; r < W; ++r) {
        unsigned char *b = (unsigned char *)m->x_peer[r];
        tri[r] = (int *)b; tri[W + r] = (int *)b + (size_t)3 * m->x_qmax;
        flg[r] = (uint32_t *)(b + x_tri_bytes(m));
    }
    CK(cudaMalloc(&m->x_tri_tab, tri.size() * sizeof(int *)));
    CK(cudaMalloc(&m->x_flag_tab, flg.size() * sizeof(uint32_t *)));
    CK(cudaMemcpy(m->x_tri_tab, tri.data(), tri.size() * sizeof(int *), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(m->x_flag_tab, flg.data(), flg.size() * sizeof(uint32_t *), cudaMemcpyHostToDevice));
    return ORBX_OK;
}

extern "C" int orbm_knn2_sharded_device(orbm_matcher *m, const uint8_t *d_query, int nq, const uint8_t *d_db_shard, int ndb_shard,
                                        int index_base, int32_t *d_d1, int32_t *d_idx1, int32_t *d_d2,
                                        int th_low, float ratio, int32_t *d_match, void *stream)
{
    if (!m || !m->x_tri_tab || nq < 1 || nq > m->x_qmax || !d_d1 || !d_idx1 || !d_d2) return ORBX_E_INVALID;
    CK(cudaSetDevice(m->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : m->stream;
    const uint32_t epoch = ++m->x_epoch;
    const int par = (int)(epoch & 1u), W = m->x_world;
    int *mine = (int *)m->x_buf + (size_t)par * 3 * m->x_qmax;       // this epoch's triple, arrays strided by x_qmax
    int rc = orbm_knn2_device(m, d_query, nq, d_db_shard, ndb_shard, index_base, mine, mine + m->x_qmax, mine + 2 * (size_t)m->x_qmax, s);
    // A local failure must not leave the peers spinning until their timeout: publish an EMPTY shard for this epoch
    // (idx = -1 rows are skipped by the merge) and the flag, run the merge like everybody else, then report the error.
    const int scan_rc = rc;
    if (rc) launch_knn2_empty(nq, mine, mine + m->x_qmax, mine + 2 * (size_t)m->x_qmax, s);
    launch_exchange_signal(m->x_flag_tab, m->x_rank, W, epoch, s);
    unsigned char *own = m->x_buf;
    launch_merge_peers((const int *const *)(m->x_tri_tab + (size_t)par * W), (const uint32_t *)(own + x_tri_bytes(m)), W, nq, (size_t)m->x_qmax,
                       epoch, d_d1, d_idx1, d_d2, th_low, ratio, d_match, (int *)(own + x_tri_bytes(m) + 64 * sizeof(uint32_t)), s);
    m->launches += 2;
    CK(cudaGetLastError());
    return scan_rc;
}

extern "C" int orbm_exchange_status(orbm_matcher *m)
{
    if (!m || !m->x_buf) return ORBX_E_INVALID;
    CK(cudaSetDevice(m->device));
    int err = 0;
    CK(cudaMemcpy(&err, m->x_buf + x_tri_bytes(m) + 64 * sizeof(uint32_t), sizeof(int), cudaMemcpyDeviceToHost));
    if (err) { std::snprintf(g_cuda_err, sizeof(g_cuda_err), "peer exchange timed out waiting for a shard"); return ORBX_E_CUDA; }
    return ORBX_OK;
}

extern "C" size_t orbm_search_init_workspace_bytes(int capacity, int npairs)
{
    if (capacity < 1 || npairs < 1) return 0;
    return (size_t)capacity * ((size_t)capacity + 1) * sizeof(uint32_t) * (size_t)npairs;
}

static int window_params_ok(const orbm_window_params *p)
{
    const bool bounds_ok = p && (p->use_bounds ? (p->max_x > p->min_x && p->max_y > p->min_y) : (p->width >= 1 && p->height >= 1));
    return p && bounds_ok && p->radius >= 0.f && (p->gate == 0 || p->gate == 1) &&
           p->query_level_min >= 0 && p->query_level_max < 16 && p->query_level_min <= p->query_level_max;
}

extern "C" int orbm_search_window_device(orbm_matcher *m, const orbx_keypoint *d_kps, const uint8_t *d_desc, const int32_t *d_counts,
                                         int capacity, const int32_t *d_pair_a, const int32_t *d_pair_b, int npairs,
                                         float *d_centers, int32_t *d_matches12, int32_t *d_nmatches,
                                         const orbm_window_params *params, void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (!m || npairs < 0 || capacity < 1 || capacity >= 65536 || !window_params_ok(params)) return ORBX_E_INVALID;
    if (npairs == 0) return ORBX_OK;
    if (!d_kps || !d_desc || !d_counts || !d_pair_a || !d_pair_b || !d_centers || !d_matches12 || !d_nmatches || !d_workspace)
        return ORBX_E_INVALID;
    if (((uintptr_t)d_desc & 15) || ((uintptr_t)d_workspace & 3)) return ORBX_E_INVALID;
    CK(cudaSetDevice(m->device));
    SearchInitArgs a;
    a.kps = d_kps; a.desc = d_desc; a.counts = d_counts; a.cap = capacity;
    a.pair_a = d_pair_a; a.pair_b = d_pair_b; a.npairs = npairs;
    a.prev_matched = d_centers; a.matches12 = d_matches12; a.nmatches = d_nmatches;
    a.w = *params;
    // octaves of F2 any query can reach; the grid (and the sort) only holds those
    a.grid_level_min = params->level_below < 0 ? 0 : params->query_level_min - params->level_below;
    a.grid_level_max = params->level_above < 0 ? INT_MAX : params->query_level_max + params->level_above;
    a.groups = nullptr; a.second_init = INT_MAX;
    a.workspace = (uint32_t *)d_workspace; a.ws_words_per_pair = workspace_bytes / sizeof(uint32_t) / (size_t)npairs;
    int sn = 32; while (sn < capacity) sn <<= 1;
    a.sort_n = sn;
    if (const int lrc = launch_search_init(a, stream ? (cudaStream_t)stream : m->stream)) {
        if (lrc == -2) {
            std::snprintf(g_cuda_err, sizeof(g_cuda_err), "keypoint capacity %d too large for the search kernel's shared-memory tables (limit about 11 600)", capacity);
            return ORBX_E_UNSUPPORTED;
        }
        return cuda_fail(cudaGetLastError(), "search_init smem");
    }
    m->launches += 1;
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbm_search_groups_device(orbm_matcher *m, const orbx_keypoint *d_kps, const uint8_t *d_desc, const int32_t *d_counts,
                                         const uint16_t *d_groups, int capacity, const int32_t *d_pair_a, const int32_t *d_pair_b, int npairs,
                                         int32_t *d_matches12, int32_t *d_nmatches, int th_dist, float nnratio, int check_orientation,
                                         void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (!m || npairs < 0 || capacity < 1 || capacity >= 65536 || th_dist < 0) return ORBX_E_INVALID;
    if (npairs == 0) return ORBX_OK;
    if (!d_kps || !d_desc || !d_counts || !d_groups || !d_pair_a || !d_pair_b || !d_matches12 || !d_nmatches || !d_workspace) return ORBX_E_INVALID;
    if (((uintptr_t)d_desc & 15) || ((uintptr_t)d_workspace & 3) || ((uintptr_t)d_groups & 1)) return ORBX_E_INVALID;
    CK(cudaSetDevice(m->device));
    SearchInitArgs a;
    std::memset(&a, 0, sizeof(a));
    a.kps = d_kps; a.desc = d_desc; a.counts = d_counts; a.cap = capacity;
    a.pair_a = d_pair_a; a.pair_b = d_pair_b; a.npairs = npairs;
    a.prev_matched = nullptr; a.matches12 = d_matches12; a.nmatches = d_nmatches;
    a.w.gate = 1; a.w.th_dist = th_dist; a.w.nnratio = nnratio; a.w.check_orientation = check_orientation ? 1 : 0;
    a.w.update_centers = 0; a.w.width = 1; a.w.height = 1;
    a.groups = d_groups; a.second_init = 256;                 // upstream starts both best distances at 256
    a.workspace = (uint32_t *)d_workspace; a.ws_words_per_pair = workspace_bytes / sizeof(uint32_t) / (size_t)npairs;
    int sn = 32; while (sn < capacity) sn <<= 1;
    a.sort_n = sn;
    if (const int lrc = launch_search_init(a, stream ? (cudaStream_t)stream : m->stream)) {
        if (lrc == -2) {
            std::snprintf(g_cuda_err, sizeof(g_cuda_err), "keypoint capacity %d too large for the search kernel's shared-memory tables (limit about 11 600)", capacity);
            return ORBX_E_UNSUPPORTED;
        }
        return cuda_fail(cudaGetLastError(), "search_init smem");
    }
    m->launches += 1;
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbm_search_groups_host(orbm_matcher *m, const orbx_keypoint *kp1, const uint8_t *desc1, const uint16_t *group1, int n1,
                                       const orbx_keypoint *kp2, const uint8_t *desc2, const uint16_t *group2, int n2,
                                       int32_t *matches12, int32_t *nmatches, int th_dist, float nnratio, int check_orientation)
{
    if (!m || n1 < 0 || n2 < 0 || n1 >= 65536 || n2 >= 65536 || !nmatches) return ORBX_E_INVALID;
    if (n1 == 0) { *nmatches = 0; return ORBX_OK; }
    if (!kp1 || !desc1 || !group1 || !matches12 || (n2 && (!kp2 || !desc2 || !group2))) return ORBX_E_INVALID;
    CK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    const int cap = (n1 > n2 ? n1 : n2) < 1 ? 1 : (n1 > n2 ? n1 : n2);
    const size_t wsb = orbm_search_init_workspace_bytes(cap, 1);
    // one scratch allocation laid out as [kps 2][desc 2][groups 2][counts 2 + pairs 2][m12][nm][workspace]
    const size_t o_kps = 0, o_desc = (o_kps + 2 * (size_t)cap * sizeof(orbx_keypoint) + 15) & ~(size_t)15,
                 o_grp = (o_desc + 2 * (size_t)cap * 32 + 15) & ~(size_t)15, o_cnt = (o_grp + 2 * (size_t)cap * 2 + 15) & ~(size_t)15,
                 o_m12 = o_cnt + 64, o_nm = o_m12 + (size_t)cap * 4, o_ws = (o_nm + 16 + 15) & ~(size_t)15;
    const size_t need = o_ws + wsb;
    if (need > m->si_bytes) {
        if (m->si_buf) CK(cudaFree(m->si_buf));
        m->si_buf = nullptr; m->si_bytes = 0;
        if (cudaMalloc(&m->si_buf, need) != cudaSuccess) { cuda_fail(cudaGetLastError(), "cudaMalloc"); return ORBX_E_NOMEM; }
        m->si_bytes = need;
    }
    unsigned char *b = (unsigned char *)m->si_buf;
    const int meta[4] = { n1, n2, 0, 1 };
    CK(cudaMemcpyAsync(b + o_kps, kp1, (size_t)n1 * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b + o_desc, desc1, (size_t)n1 * 32, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b + o_grp, group1, (size_t)n1 * 2, cudaMemcpyHostToDevice, s));
    if (n2) {
        CK(cudaMemcpyAsync(b + o_kps + (size_t)cap * sizeof(orbx_keypoint), kp2, (size_t)n2 * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(b + o_desc + (size_t)cap * 32, desc2, (size_t)n2 * 32, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(b + o_grp + (size_t)cap * 2, group2, (size_t)n2 * 2, cudaMemcpyHostToDevice, s));
    }
    CK(cudaMemcpyAsync(b + o_cnt, meta, sizeof(meta), cudaMemcpyHostToDevice, s));
    int rc = orbm_search_groups_device(m, (const orbx_keypoint *)(b + o_kps), b + o_desc, (const int *)(b + o_cnt), (const uint16_t *)(b + o_grp), cap,
                                       (const int *)(b + o_cnt) + 2, (const int *)(b + o_cnt) + 3, 1, (int *)(b + o_m12), (int *)(b + o_nm),
                                       th_dist, nnratio, check_orientation, b + o_ws, wsb, s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(matches12, b + o_m12, (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(nmatches, b + o_nm, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ORBX_OK;
}

// SearchForInitialization as an instance of the windowed search: octave-0 queries, octave-0 candidates, fixed window,
// TH_LOW = 50 (src/ORBmatcher.cpp:7), displacement gate, centres updated
static orbm_window_params search_init_params(int window, float nnratio, int check_orientation, int width, int height, int literal_gridid_bug)
{
    orbm_window_params p;
    std::memset(&p, 0, sizeof(p));
    p.radius = (float)window;
    for (int i = 0; i < 16; ++i) p.level_scale[i] = 1.0f;
    p.query_level_min = 0; p.query_level_max = 0; p.level_below = 0; p.level_above = 0;
    p.gate = 0; p.th_dist = 50; p.nnratio = nnratio; p.check_orientation = check_orientation ? 1 : 0; p.update_centers = 1;
    p.width = width; p.height = height; p.literal_gridid_bug = literal_gridid_bug ? 1 : 0;
    return p;
}

extern "C" size_t orbm_knn2_pairs_workspace_bytes(int capacity, int npairs)
{
    if (capacity < 1 || npairs < 1) return 0;
    return knn2_pairs_workspace_bytes(capacity, npairs);
}

extern "C" int orbm_knn2_pairs_device(orbm_matcher *m, const uint8_t *d_desc, const int32_t *d_counts, int capacity,
                                      const int32_t *d_pair_a, const int32_t *d_pair_b, int npairs,
                                      int32_t *d_d1, int32_t *d_idx1, int32_t *d_d2, int th_low, float ratio, int32_t *d_match,
                                      void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (!m || npairs < 0 || capacity < 1 || capacity >= (1 << 22)) return ORBX_E_INVALID;
    if (npairs == 0) return ORBX_OK;
    if (!d_desc || !d_counts || !d_pair_a || !d_pair_b || !d_d1 || !d_idx1 || !d_d2 || !d_workspace) return ORBX_E_INVALID;
    if (npairs > 65535) return ORBX_E_CAPACITY;                        // grid.z / grid.y
    if (workspace_bytes < knn2_pairs_workspace_bytes(capacity, npairs)) return ORBX_E_CAPACITY;
    CK(cudaSetDevice(m->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : m->stream;
    launch_knn2_pairs(d_desc, d_counts, capacity, d_pair_a, d_pair_b, npairs, (uint2 *)d_workspace, d_d1, d_idx1, d_d2, th_low, ratio, d_match, s);
    m->launches += 2;
    CK(cudaGetLastError());
    return ORBX_OK;
}

extern "C" int orbm_search_init_device(orbm_matcher *m, const orbx_keypoint *d_kps, const uint8_t *d_desc, const int32_t *d_counts,
                                       int capacity, const int32_t *d_pair_a, const int32_t *d_pair_b, int npairs,
                                       float *d_prev_matched, int32_t *d_matches12, int32_t *d_nmatches,
                                       int window, float nnratio, int check_orientation, int width, int height,
                                       int literal_gridid_bug, void *d_workspace, size_t workspace_bytes, void *stream)
{
    if (width < 1 || height < 1 || window < 0) return ORBX_E_INVALID;
    const orbm_window_params p = search_init_params(window, nnratio, check_orientation, width, height, literal_gridid_bug);
    return orbm_search_window_device(m, d_kps, d_desc, d_counts, capacity, d_pair_a, d_pair_b, npairs, d_prev_matched, d_matches12,
                                     d_nmatches, &p, d_workspace, workspace_bytes, stream);
}

extern "C" int orbm_search_window_host(orbm_matcher *m, const orbx_keypoint *kp1, const uint8_t *desc1, int n1,
                                       const orbx_keypoint *kp2, const uint8_t *desc2, int n2,
                                       float *centers, int32_t *matches12, int32_t *nmatches, const orbm_window_params *params)
{
    if (!m || n1 < 0 || n2 < 0 || n1 >= 65536 || n2 >= 65536 || !nmatches || !window_params_ok(params)) return ORBX_E_INVALID;
    if (n1 == 0) { *nmatches = 0; return ORBX_OK; }
    if (!kp1 || !desc1 || !centers || !matches12 || (n2 && (!kp2 || !desc2))) return ORBX_E_INVALID;
    CK(cudaSetDevice(m->device));
    cudaStream_t s = m->stream;
    const int cap = (n1 > n2 ? n1 : n2) < 1 ? 1 : (n1 > n2 ? n1 : n2);
    const size_t wsb = orbm_search_init_workspace_bytes(cap, 1);
    // one scratch allocation laid out as [kps 2][desc 2][counts 2][pairs 2][centres][m12][nm][workspace]
    const size_t o_kps = 0, o_desc = (o_kps + 2 * (size_t)cap * sizeof(orbx_keypoint) + 15) & ~(size_t)15, o_cnt = (o_desc + 2 * (size_t)cap * 32 + 15) & ~(size_t)15,
                 o_prev = o_cnt + 64, o_m12 = o_prev + (size_t)cap * 8, o_nm = o_m12 + (size_t)cap * 4, o_ws = (o_nm + 16 + 15) & ~(size_t)15;
    const size_t need = o_ws + wsb;
    if (need > m->si_bytes) {
        if (m->si_buf) CK(cudaFree(m->si_buf));
        m->si_buf = nullptr; m->si_bytes = 0;
        if (cudaMalloc(&m->si_buf, need) != cudaSuccess) { cuda_fail(cudaGetLastError(), "cudaMalloc"); return ORBX_E_NOMEM; }
        m->si_bytes = need;
    }
    unsigned char *b = (unsigned char *)m->si_buf;
    const int meta[4] = { n1, n2, 0, 1 };                                  // counts[2], pair_a = 0, pair_b = 1
    CK(cudaMemcpyAsync(b + o_kps, kp1, (size_t)n1 * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b + o_desc, desc1, (size_t)n1 * 32, cudaMemcpyHostToDevice, s));
    if (n2) {
        CK(cudaMemcpyAsync(b + o_kps + (size_t)cap * sizeof(orbx_keypoint), kp2, (size_t)n2 * sizeof(orbx_keypoint), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(b + o_desc + (size_t)cap * 32, desc2, (size_t)n2 * 32, cudaMemcpyHostToDevice, s));
    }
    CK(cudaMemcpyAsync(b + o_cnt, meta, sizeof(meta), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(b + o_prev, centers, (size_t)n1 * 8, cudaMemcpyHostToDevice, s));
    int rc = orbm_search_window_device(m, (const orbx_keypoint *)(b + o_kps), b + o_desc, (const int *)(b + o_cnt), cap,
                                       (const int *)(b + o_cnt) + 2, (const int *)(b + o_cnt) + 3, 1,
                                       (float *)(b + o_prev), (int *)(b + o_m12), (int *)(b + o_nm), params, b + o_ws, wsb, s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(matches12, b + o_m12, (size_t)n1 * 4, cudaMemcpyDeviceToHost, s));
    if (params->update_centers) CK(cudaMemcpyAsync(centers, b + o_prev, (size_t)n1 * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(nmatches, b + o_nm, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return ORBX_OK;
}

extern "C" int orbm_search_init_host(orbm_matcher *m, const orbx_keypoint *kp1, const uint8_t *desc1, int n1,
                                     const orbx_keypoint *kp2, const uint8_t *desc2, int n2,
                                     float *prev_matched, int32_t *matches12, int32_t *nmatches,
                                     int window, float nnratio, int check_orientation, int width, int height, int literal_gridid_bug)
{
    if (width < 1 || height < 1 || window < 0) return ORBX_E_INVALID;
    const orbm_window_params p = search_init_params(window, nnratio, check_orientation, width, height, literal_gridid_bug);
    return orbm_search_window_host(m, kp1, desc1, n1, kp2, desc2, n2, prev_matched, matches12, nmatches, &p);
}

extern "C" int orbm_popc_peak(int device, double *popc_per_second, double *lop3_per_second)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { cudaGetLastError(); return ORBX_E_NODEVICE; }
    CK(cudaSetDevice(device));
    int sm = 0;
    CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device));
    double a = 0, b = 0;
    if (run_popc_bench(0, sm, &a)) return cuda_fail(cudaGetLastError(), "popc bench");
    if (run_popc_bench(1, sm, &b)) return cuda_fail(cudaGetLastError(), "popc+lop3 bench");
    if (popc_per_second) *popc_per_second = a;
    if (lop3_per_second) *lop3_per_second = b;
    return ORBX_OK;
}

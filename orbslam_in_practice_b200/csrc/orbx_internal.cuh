// orbx_internal.cuh -- shared definitions of the sm_100a ORB front end (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/orbx.h"

namespace orbx {

constexpr int kPadX = 32;            // left pad of every pyramid row (>= 19, keeps the interior 32 B aligned)
constexpr int kPadY = ORBX_EDGE_THRESHOLD;
constexpr int kBorder = ORBX_EDGE_THRESHOLD;
constexpr int kMinBlurBorder = 4;     // border always written: the blur kernel reads 3 px (+1 for word alignment) outside the level
constexpr int kHalfPatch = 15;       // HALF_PATCH_SIZE, ORBextractor.cpp:23
constexpr int kMinBorder = 16;       // EDGE_THRESHOLD - 3, ORBextractor.cpp:729
constexpr int kResizeRows = 16;      // most destination rows per block of the staged resize kernel (LevelGeom::rs_rows is the level's choice)

// packed FAST candidate: x (12 bit) | y (12 bit) << 12 | score (8 bit) << 24, coordinates
// relative to (16,16) of the level (the values of vToDistributeKeys).
__host__ __device__ inline uint32_t pack_cand(int x, int y, int score) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)score << 24); }
__host__ __device__ inline int cand_x(uint32_t k) { return (int)(k & 0xfffu); }
__host__ __device__ inline int cand_y(uint32_t k) { return (int)((k >> 12) & 0xfffu); }
__host__ __device__ inline int cand_score(uint32_t k) { return (int)(k >> 24); }

struct LevelGeom {
    int w, h;                 // level size (cvRound((float)dim * invScale), ORBextractor.cpp:1075-1076)
    int pitch;                // bytes per row of the bordered level buffer
    int blur_pitch;
    unsigned long long base;          // byte offset of frame 0's buffer for this level inside pyr
    unsigned long long frame_stride;  // bytes between consecutive frames of this level
    unsigned long long blur_base, blur_frame_stride;
    // per-cell FAST grid, ORBextractor.cpp:729-743
    int nCols, nRows, wCell, hCell, maxBorderX, maxBorderY;
    int cell_base;            // first cell index of this level inside a frame's cell arrays
    int cell_cap;             // candidate slots per cell: ceil(wCell/2)*ceil(hCell/2) (NMS bound)
    unsigned long long slot_base;     // u32 offset of this level's slots inside a frame's slot region
    // octree, ORBextractor.cpp:489-513
    int N, nIni, regionW, regionH;
    float hX;
    int max_cand;             // nCells * cell_cap
    unsigned long long key_base;      // u32 offset of this level's key scratch inside a frame's region
    int kept_cap, kept_base;  // survivors bound and offset inside a frame's kept array
    int node_cap;             // octree node table capacity
    int inv_wCell, inv_hCell; // ceil(2^18 / cell size): (v * inv) >> 18 == v / size for 0 <= v < 4096 and sizes <= 64
    int oct_B;                // depth of the closed-form phase 1 of the octree kernel (bins = nIni * 4^B), 0 = off
    float scale;              // mvScaleFactor[level]
    int patch_size;           // (int)(31 * scale), ORBextractor.cpp:794
    // resize tables (levels >= 1): entries {src offset, c0 | c1 << 16}
    int tabx, taby;           // offsets into the int2 table array
    // staged resize (pyramid.cu k_resize): padded tables indexed by padded coordinates, tile bounds from the host
    int tabxp;                // [pitch] entry(i) = tabx[reflect101(clamp(i - kPadX, -19, w + 18))]
    int tabyp;                // [h + 38] entry(Y + 19) = {sy0 | sy1 << 16 (clamped source rows), cy0 | cy1 << 16}
    int rs_bw;                // threads per block (4 pixels each)
    int rs_rows;              // destination rows per block (4, 8 or 16 <= kResizeRows), chosen per level by the host
    int rs_tile_w, rs_tile_h; // shared-memory source tile: bytes per row (multiple of 16) and rows, maxima over all blocks
    int rs_xr, rs_yr;         // per-block source ranges {lo, hi}: [2 border variants][blocks in x], [variant 0 blocks in y][variant 1 ...]
    int rs_nbx, rs_nby0;      // blocks in x; blocks in y of variant 0 (border 4)
    int rs_staged;            // 0: the level needs the gather kernel (adjacent source offsets further than 3 apart: scale > 3)
};

struct Geo {
    int nlevels, ini_th, min_th, border_on;
    int frame0;               // first frame handled by this launch (chunked host pipeline)
    int in_channels, in_rgb;  // input pixel format: 1 / 3 / 4 interleaved channels, RGB(A) or BGR(A) order
    int total_cells;          // cells per frame, all levels
    int capacity;             // output keypoint slots per frame
    int kept_total;           // kept slots per frame (= capacity)
    int oct_node_cap_max;     // largest node_cap over the levels (sizes the octree's shared-memory carve-up)
    int fast_cell_lo, fast_cell_hi;   // cell range of one k_fast_cells launch (set by launch_fast)
    unsigned long long pyr_frame_total, blur_frame_total; // allocation sizes (bytes, all frames) of the pyramid / blur buffers
    unsigned long long slots_per_frame, keys_per_frame;
    LevelGeom lv[ORBX_MAX_LEVELS];
    int umax[16];
};

struct DevBuffers {
    uint8_t *pyr;        // bordered pyramid levels, level-major then frame
    uint8_t *blur;       // blurred levels
    int2 *tables;        // resize tables
    int4 *cell_tab;      // [total_cells] FAST cell descriptors {level, iniX | iniY << 16, cw | ch << 16 (0: skipped cell), slot offset}
    int *cell_count;     // [F][total_cells]
    uint32_t *cell_slots;// [F][slots_per_frame]
    uint32_t *keysA, *keysB;   // [F][keys_per_frame]
    uint16_t *nodeA, *nodeB;   // [F][keys_per_frame]
    uint4 *scanE;        // [F][keys_per_frame + nlevels] octree per-key scratch (16 B/key: u32 rank + u8 quadrant are carved from it)
    int *ncand;          // [F][nlevels]
    uint32_t *kept;      // [F][kept_total]
    int *nkept;          // [F][nlevels]
    uint8_t *staging;    // device staging for host inputs [max_batch][max_h][max_w]
    orbx_keypoint *out_kps; uint8_t *out_desc; int *out_counts; // device outputs for the host path
    unsigned char *oct_scratch; unsigned long long oct_scratch_stride;   // octree node tables in global memory when they do not fit one SM (else NULL)
};

// kernels (defined in the .cu files)
void launch_level0(const Geo &g, const DevBuffers &b, const uint8_t *d_imgs, size_t pitch, size_t fstride, int nframes, cudaStream_t s);
void launch_resize(const Geo &g, const DevBuffers &b, int level, int nframes, cudaStream_t s);
void launch_fill_border(const Geo &g, const DevBuffers &b, int level, int frame, cudaStream_t s);
void launch_fast(const Geo &g, const DevBuffers &b, int nframes, cudaStream_t s, int level_lo = 0, int level_hi = ORBX_MAX_LEVELS);
void launch_octree(const Geo &g, const DevBuffers &b, int nframes, int smem_bytes, cudaStream_t s, int level_lo = 0, int level_hi = ORBX_MAX_LEVELS);
void launch_blur(const Geo &g, const DevBuffers &b, int nframes, cudaStream_t s);
void launch_describe(const Geo &g, const DevBuffers &b, int nframes, orbx_keypoint *d_kps, uint8_t *d_desc, int *d_counts, cudaStream_t s);
void launch_undistort(const orbx_keypoint *in, orbx_keypoint *out, int n, const float *cam, const float *dist, int literal_bug, cudaStream_t s);
int octree_smem_bytes(const Geo &g);
int octree_table_bytes(const Geo &g);
int octree_configure(int smem_bytes);

// search_init.cu: one windowed-search launch (SearchForInitialization and the projection-style search share it)
struct SearchInitArgs {
    const orbx_keypoint *kps; const uint8_t *desc; const int *counts; int cap;      // extractor outputs [F][cap]
    const int *pair_a, *pair_b; int npairs;
    float *prev_matched; int *matches12; int *nmatches;                              // [npairs][cap][2], [npairs][cap], [npairs]
    orbm_window_params w;                                                            // see include/orbx.h
    int grid_level_min, grid_level_max;                                              // octaves of F2 worth putting in the grid
    const unsigned short *groups;                                                    // [F][cap] group (vocabulary node) per keypoint, 0xffff = none; NULL = windowed search
    int second_init;                                                                 // value of the second-best distance when there is none (INT_MAX / 256)
    uint32_t *workspace; unsigned long long ws_words_per_pair;
    int sort_n;                                                                      // power of two >= cap
};
int launch_search_init(const SearchInitArgs &a, cudaStream_t s);

} // namespace orbx

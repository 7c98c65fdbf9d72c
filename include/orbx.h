/*
 * orbx.h -- C ABI of the B200-native ORB front end (liborbx.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, int status codes, no C++/torch types.
 * The C++ classes ORBSlam::ORBextractor / ORBSlam::ORBmatcher in orbslam_in_practice_b200/cpp/
 * keep the reference's signatures and forward to these entry points.  Reference interfaces
 * replaced (paths relative to the reference tree):
 *
 *   orbx_create / orbx_tables        ORBextractor::ORBextractor        include/ORBextractor.h:35-36, src/ORBextractor.cpp:360-420
 *                                    GetLevels/GetScaleFactor(s)/...   include/ORBextractor.h:47-69
 *   orbx_extract_host/_device        ORBextractor::operator()          include/ORBextractor.h:43-45, src/ORBextractor.cpp:1001-1065
 *   orbx_set_input_format            cv::cvtColor(..2GRAY) before the extractor   src/Tracking.cpp:57-70
 *   orbx_undistort_keypoints_device  Frame::UndistortedKeyPoints   src/Frame.cpp:80-109
 *   orbx_download_level              mvImagePyramid (public member)    include/ORBextractor.h:71, src/ORBextractor.cpp:1071-1096
 *   orbm_hamming_pairs_host          ORBmatcher::DescriptorDistance    include/ORBmatcher.h:19, src/ORBmatcher.cpp:128-144
 *   orbm_knn2_* / orbm_ratio_select  best-2 scan + acceptance          src/ORBmatcher.cpp:37-67
 *   orbm_search_init_device          ORBmatcher::SearchForInitialization + Frame grid   src/ORBmatcher.cpp:9-126, src/Frame.cpp:144-168,219-271
 *   orbm_search_window_device        ORBmatcher::SearchByProjection (empty in the reference) over Frame::GetFeaturesInArea   include/ORBmatcher.h:24, src/Frame.cpp:219-271
 *   orbm_search_groups_device        ORBmatcher::SearchByBoW (empty in the reference; group = vocabulary node)   include/ORBmatcher.h:22
 *   orbm_merge_shards_device         (database sharding; no reference counterpart)
 *
 * All functions return 0 on success or a negative ORBX_E_* code; nothing throws across the ABI.
 * Handles are NOT thread-safe (the reference extractor is stateful too: one call in flight).
 * There is no CPU fallback: every entry point needs a CUDA device of compute capability 10.0.
 */
#ifndef ORBX_H
#define ORBX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBX_OK 0
#define ORBX_E_INVALID (-1)   /* bad argument */
#define ORBX_E_CUDA (-2)      /* CUDA runtime error (see orbx_last_cuda_error) */
#define ORBX_E_NOMEM (-3)     /* allocation failed */
#define ORBX_E_CAPACITY (-4)  /* frame/batch larger than the handle was created for */
#define ORBX_E_NODEVICE (-5)  /* no usable sm_100 device */
#define ORBX_E_UNSUPPORTED (-6) /* outside the supported range: a level wider/taller than 4095 + 32 px, a FAST cell > 64 px, or a per-level
                                  feature quota above 65 531 (16-bit node ids).  Quotas above ~2 400 per level run with the octree's node
                                  tables in global instead of shared memory (slower, same results) */

#define ORBX_MAX_LEVELS 16
#define ORBX_EDGE_THRESHOLD 19 /* pyramid border, src/ORBextractor.cpp:24 */

typedef struct {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels;
    int32_t ini_th_fast;
    int32_t min_th_fast;
} orbx_params;

/* Binary layout identical to cv::KeyPoint (28 bytes): the C++ wrapper memcpy's these. */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbx_keypoint;

/* FAST candidate / octree survivor, coordinates relative to (minBorderX, minBorderY) = (16,16) of
 * its level, exactly the values held by vToDistributeKeys (src/ORBextractor.cpp:777-782). */
typedef struct {
    int16_t x, y;
    int32_t score;
} orbx_cand;

typedef struct orbx_extractor orbx_extractor;
typedef struct orbm_matcher orbm_matcher;

const char *orbx_strerror(int code);
const char *orbx_last_cuda_error(void);
int orbx_version(void);
/* 16 hex digits identifying the kernel sources this library was built from (sha256 of csrc + orbx.h); profiles record it so
 * that bench.py only quotes ncu counters captured on the same object code */
const char *orbx_build_id(void);
/* number of visible CUDA devices with compute capability 10.x (0 if none / no driver) */
int orbx_device_count(void);

/* ---------------- extractor ---------------- */
int orbx_create(const orbx_params *params, int max_width, int max_height, int max_batch, int device,
                orbx_extractor **out);
int orbx_destroy(orbx_extractor *ex);
int orbx_nlevels(const orbx_extractor *ex);
/* row stride (in keypoints) of the per-frame output arrays: sum over levels of the octree bound */
int orbx_capacity(const orbx_extractor *ex);
/* arrays of nlevels entries (umax: 16); any pointer may be NULL */
int orbx_tables(const orbx_extractor *ex, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2,
                int32_t *features_per_level, int32_t *umax);

/* operator() on a batch of same-sized 8-bit grayscale frames held in HOST memory.
 *   imgs: frame f, row y at imgs + f*frame_stride + y*row_pitch
 *   kps  [nframes][capacity], desc [nframes][capacity][32], counts [nframes] (host)
 * Copies in, runs the kernels, copies out and synchronises before returning.  Zero-sized images
 * return ORBX_OK with counts = 0 (the reference returns silently, src/ORBextractor.cpp:1004). */
int orbx_extract_host(orbx_extractor *ex, const uint8_t *imgs, size_t row_pitch, size_t frame_stride,
                      int width, int height, int nframes,
                      orbx_keypoint *kps, uint8_t *desc, int32_t *counts);

/* Split form of orbx_extract_host for pipelining across handles: _begin enqueues the uploads, kernels and downloads
 * of the batch and returns at once; _end waits for them (the output buffers are valid after _end).  With two handles,
 * begin(A, batch k+1) can be issued before end(B, batch k), so A's PCIe upload overlaps B's kernels.  The host
 * buffers must stay alive and unmodified between _begin and _end, and should be pinned (pageable memory makes the
 * copies synchronous).  One batch in flight per handle. */
int orbx_extract_host_begin(orbx_extractor *ex, const uint8_t *imgs, size_t row_pitch, size_t frame_stride,
                            int width, int height, int nframes,
                            orbx_keypoint *kps, uint8_t *desc, int32_t *counts);
int orbx_extract_host_end(orbx_extractor *ex);

/* Same with DEVICE pointers; asynchronous on `stream` (a cudaStream_t; NULL = the handle's stream). */
int orbx_extract_device(orbx_extractor *ex, const uint8_t *d_imgs, size_t row_pitch, size_t frame_stride,
                        int width, int height, int nframes,
                        orbx_keypoint *d_kps, uint8_t *d_desc, int32_t *d_counts, void *stream);

/* Small batches (<= 8 frames; the reference's caller passes ONE frame per call, src/Frame.cpp:75-78) take a low-latency
 * path by default: the call's kernels are replayed as a CUDA graph whose nodes follow the real dependencies (FAST + octree
 * of level l start as soon as level l of the resize chain exists), and the host entry points use one stream, pinned
 * staging for pageable caller buffers and one download.  Results are identical.  enabled = 0 restores the stream path. */
int orbx_set_low_latency(orbx_extractor *ex, int enabled);

/* Debug aid: a handle created while the environment has ORBX_GUARD=1 surrounds every device buffer it owns with a 4 KB canary
 * zone; this call synchronises the device and checks all of them.  ORBX_OK = intact, ORBX_E_CUDA = a kernel wrote outside a
 * buffer (*bad_buffer = its index in allocation order), ORBX_E_UNSUPPORTED = the handle has no guard zones. */
int orbx_debug_guard_check(orbx_extractor *ex, int *bad_buffer);

/* Input pixel format of the following extract calls: channels = 1 (gray, default), 3 or 4 interleaved 8-bit
 * channels; rgb_order = 0 for BGR(A), 1 for RGB(A).  Colour input is converted to gray inside the level-0 kernel with
 * OpenCV's 8U fixed point (R*9798 + G*19235 + B*3735 + 16384) >> 15, replacing the cv::cvtColor call in front of the
 * extractor (src/Tracking.cpp:57-70).  row_pitch / frame_stride of the extract calls stay in bytes. */
int orbx_set_input_format(orbx_extractor *ex, int channels, int rgb_order);

/* Pyramid border policy.  The extractor's own kernels read at most 4 px outside a level, and that much
 * reflect-101 border is always written.  The full 19-px border of the reference's mvImagePyramid
 * (src/ORBextractor.cpp:1086-1092) exists for outside consumers only: by default (0, lazy) it is
 * materialised by orbx_download_level(border = 19) for the level/frame that is downloaded; with 1 (eager)
 * every extract call writes it for all levels and frames on the device. */
int orbx_set_pyramid_border(orbx_extractor *ex, int enabled);

/* orbx_extract_device runs a batch of >= 32 frames as `nsplit` independent sub-batches alternating between the caller's
 * stream and an internal one (joined before it returns control of the stream): the sub-batches' different kernels fill
 * each other's idle issue slots.  Default 2 (environment ORBX_DEVICE_SPLIT overrides at creation).  Use 1 when the
 * caller already overlaps consecutive batches itself with several handles on several streams. */
int orbx_set_device_split(orbx_extractor *ex, int nsplit);

/* --- stage access of the LAST extract call (synchronous; for mvImagePyramid and parity tests) --- */
int orbx_level_dims(const orbx_extractor *ex, int level, int *width, int *height);
/* copies level pixels of `frame` to host.  border = 0 -> width x height; border = 19 -> the
 * (width+38) x (height+38) bordered image.  blurred != 0 -> the Gaussian-blurred level (border 0). */
int orbx_download_level(orbx_extractor *ex, int frame, int level, int blurred, int border,
                        uint8_t *dst, size_t dst_pitch);
/* device pointer to pixel (0,0) of the level interior of `frame` and its row pitch */
int orbx_level_device_ptr(orbx_extractor *ex, int frame, int level, const uint8_t **ptr, size_t *pitch);
/* FAST candidates in reference order (cell row-major, then row-major inside the cell) */
int orbx_download_candidates(orbx_extractor *ex, int frame, int level, orbx_cand *out, int cap, int *n);
/* octree survivors in reference list order */
int orbx_download_kept(orbx_extractor *ex, int frame, int level, orbx_cand *out, int cap, int *n);
/* upper bound of candidates a level can produce (sizing for orbx_download_candidates) */
int orbx_max_candidates(const orbx_extractor *ex, int level);
/* kernels launched by this handle since creation (for bench.py's gpu_launches) */
long long orbx_launch_count(const orbx_extractor *ex);
/* Per-stage device timing.  When enabled, every extract call brackets its stages with CUDA events
 * on the launch stream; orbx_stage_times returns the mean durations (ms) over the calls made since
 * profiling was enabled (at most the last 64) for
 * stages {level0, resize chain, FAST, octree, blur, describe} (ORBX_NUM_STAGES floats).
 * Synchronises the stream. */
#define ORBX_NUM_STAGES 6
int orbx_set_profiling(orbx_extractor *ex, int enabled);
int orbx_stage_times(orbx_extractor *ex, float *ms);

/* Frame::UndistortedKeyPoints (src/Frame.cpp:80-109): cv::undistortPoints(pts, K, dist, R = I, P = K) applied to
 * pt.x / pt.y of n keypoints (device pointers; in may equal out; all other fields are copied).
 * cam = {fx, fy, cx, cy}, dist = {k1, k2, p1, p2, k3} (host pointers).  dist[0] == 0 -> plain copy (:82-86).
 * literal_bug != 0 reproduces :106 as written (the undistorted x is stored into y as well). */
int orbx_undistort_keypoints_device(orbx_extractor *ex, const orbx_keypoint *d_in, orbx_keypoint *d_out, int n,
                                    const float *cam, const float *dist, int literal_bug, void *stream);

/* ---------------- matcher ---------------- */
int orbm_create(int max_queries, int max_db, int device, orbm_matcher **out);
int orbm_destroy(orbm_matcher *m);
long long orbm_launch_count(const orbm_matcher *m);

/* DescriptorDistance for n independent pairs of 32-byte rows (host pointers). */
int orbm_hamming_pairs_host(orbm_matcher *m, const uint8_t *a, const uint8_t *b, int n, int32_t *dist);

/* Brute-force best-2: for each query scan db rows in ascending index; strict '<' (first minimal
 * index wins); d2 = second smallest distance (may equal d1).  idx1 = index_base + row.
 * Empty db -> d1 = d2 = INT32_MAX, idx1 = -1.  Device pointers, async on stream. */
int orbm_knn2_device(orbm_matcher *m, const uint8_t *d_query, int nq, const uint8_t *d_db, int ndb,
                     int index_base, int32_t *d_d1, int32_t *d_idx1, int32_t *d_d2, void *stream);
/* Host-pointer convenience (copies in/out, synchronous). */
int orbm_knn2_host(orbm_matcher *m, const uint8_t *query, int nq, const uint8_t *db, int ndb,
                   int index_base, int32_t *d1, int32_t *idx1, int32_t *d2);
/* match[i] = idx1[i] if d1 <= th_low && (float)d1 < (float)d2 * ratio else -1 (ORBmatcher.cpp:65-67) */
int orbm_ratio_select_device(orbm_matcher *m, const int32_t *d_d1, const int32_t *d_idx1, const int32_t *d_d2,
                             int nq, int th_low, float ratio, int32_t *d_match, void *stream);
/* Merge `nshards` per-shard triples (shards in ascending index-range order) into the unsharded
 * result.  Shard s of each input array starts at element s * shard_stride (0 = nq, i.e. [shard][nq];
 * 3 * nq for an all-gathered [shard][3][nq] buffer). */
int orbm_merge_shards_device(orbm_matcher *m, const int32_t *d_d1, const int32_t *d_idx1, const int32_t *d_d2,
                             int nshards, int nq, size_t shard_stride,
                             int32_t *d_od1, int32_t *d_oidx1, int32_t *d_od2, void *stream);
/* device time of the last orbm_knn2_device call's scan kernel and merge kernel (ms); needs
 * orbm_set_profiling(m, 1) before the call.  Synchronises the stream. */
int orbm_set_profiling(orbm_matcher *m, int enabled);
int orbm_knn2_times(orbm_matcher *m, float *scan_ms, float *merge_ms);

/* --- database-sharded search with a peer-memory exchange (one process per GPU on one NVLink/NVSwitch box) ---
 * Setup: every rank calls orbm_exchange_create and publishes the returned 64-byte IPC handle (e.g. with
 * torch.distributed.all_gather); every rank then calls orbm_exchange_open with all handles in rank order.
 * Per search: orbm_knn2_sharded_device scans this rank's shard (rows index_base .. index_base + ndb_shard - 1 of the
 * global database; ranks own ascending ranges), publishes the result to the peers, and runs ONE fused kernel that waits for
 * all shards, reads their (d1, idx1, d2) over NVLink peer loads and writes the merged best-2 and the ratio-test matches
 * (d_match may be NULL).  Results are identical on every rank and identical to an unsharded scan.
 * orbm_exchange_status returns ORBX_E_CUDA if a previous fused kernel timed out waiting for a peer. */
#define ORBM_IPC_HANDLE_BYTES 64
int orbm_exchange_create(orbm_matcher *m, int max_queries, int rank, int world, unsigned char *handle_out);
int orbm_exchange_open(orbm_matcher *m, const unsigned char *handles /* world x ORBM_IPC_HANDLE_BYTES */);
int orbm_knn2_sharded_device(orbm_matcher *m, const uint8_t *d_query, int nq, const uint8_t *d_db_shard, int ndb_shard,
                             int index_base, int32_t *d_d1, int32_t *d_idx1, int32_t *d_d2,
                             int th_low, float ratio, int32_t *d_match, void *stream);
int orbm_exchange_status(orbm_matcher *m);

/* Many independent brute-force scans in one launch pair: descriptor rows of frame d_pair_a[p] (queries) against those of frame
 * d_pair_b[p] (database), both inside the extractor's output layout d_desc[nframes][capacity][32] with d_counts[nframes] on the
 * device -- e.g. left -> right matching of a batch of stereo pairs (BASELINE configs[2]; the reference has no stereo matcher,
 * this is its best-2 scan src/ORBmatcher.cpp:37-62 + acceptance :65-67 applied per pair).  Outputs are [npairs][capacity]:
 * d1 / idx1 (row inside frame b, -1 if it has none) / d2 and, if d_match != NULL, the accepted row or -1; entries past the
 * query frame's count are (INT_MAX, -1, INT_MAX, -1).  workspace: orbm_knn2_pairs_workspace_bytes(capacity, npairs). */
size_t orbm_knn2_pairs_workspace_bytes(int capacity, int npairs);
int orbm_knn2_pairs_device(orbm_matcher *m, const uint8_t *d_desc, const int32_t *d_counts, int capacity,
                           const int32_t *d_pair_a, const int32_t *d_pair_b, int npairs,
                           int32_t *d_d1, int32_t *d_idx1, int32_t *d_d2, int th_low, float ratio, int32_t *d_match,
                           void *d_workspace, size_t workspace_bytes, void *stream);

/* ORBmatcher::SearchForInitialization (src/ORBmatcher.cpp:9-126) on extractor outputs, including the Frame grid
 * it queries (Frame::AssignFeaturesToGrid / GetFeaturesInArea, src/Frame.cpp:144-168, 219-271): windowed best-2
 * with the sequential one-to-one gate, TH_LOW / ratio acceptance, rotation-histogram filter, prev-matched update.
 *   d_kps/d_desc/d_counts: [nframes][capacity] arrays as written by orbx_extract_device
 *   pair p matches frame d_pair_a[p] (reference frame F1) against frame d_pair_b[p] (F2)
 *   d_prev_matched [npairs][capacity][2] float, in/out (vbPrevMatched);  d_matches12 [npairs][capacity] (vnMatches12)
 *   d_nmatches [npairs]: the return value of the reference function, or -1 if the workspace was too small
 *   width/height: image size (grid bounds of an undistorted frame, src/Frame.cpp:113-118)
 *   literal_gridid_bug != 0 reproduces src/Frame.cpp:164 as written (y cell computed against miMaxY: no candidates)
 *   workspace: device scratch, orbm_search_init_workspace_bytes(capacity, npairs) always suffices.
 * capacity must be < 65536. */
size_t orbm_search_init_workspace_bytes(int capacity, int npairs);
int orbm_search_init_device(orbm_matcher *m, const orbx_keypoint *d_kps, const uint8_t *d_desc, const int32_t *d_counts,
                            int capacity, const int32_t *d_pair_a, const int32_t *d_pair_b, int npairs,
                            float *d_prev_matched, int32_t *d_matches12, int32_t *d_nmatches,
                            int window, float nnratio, int check_orientation, int width, int height,
                            int literal_gridid_bug, void *d_workspace, size_t workspace_bytes, void *stream);

/* Host-pointer convenience for ONE frame pair (what ORBSlam::ORBmatcher::SearchForInitialization forwards to):
 * copies both frames' keypoints/descriptors and prev_matched in, runs the kernel, copies matches12 (n1 ints),
 * the updated prev_matched (n1 x 2 floats) and the match count out; synchronous.  n1, n2 < 65536. */
int orbm_search_init_host(orbm_matcher *m, const orbx_keypoint *kp1, const uint8_t *desc1, int n1,
                          const orbx_keypoint *kp2, const uint8_t *desc2, int n2,
                          float *prev_matched, int32_t *matches12, int32_t *nmatches,
                          int window, float nnratio, int check_orientation, int width, int height, int literal_gridid_bug);

/* Windowed search with per-query windows: the generalisation of the kernel above that the reference's two empty
 * matcher entry points need (SearchByProjection include/ORBmatcher.h:24, SearchByBoW :22).
 * The window query is Frame::GetFeaturesInArea(x, y, r, minLevel, maxLevel) as the reference wrote it
 * (src/Frame.cpp:219-271, level filter :245-258); the loop around it follows upstream ORB-SLAM2's frame-to-frame
 * SearchByProjection, because the reference's body is empty -- parity for that loop is pinned only through the
 * SearchForInitialization instance of the same kernel.
 *   query q of frame F1 searches F2 around d_centers[q] (the projected position; a NaN x skips the query) with
 *   r = radius * level_scale[octave(q)], candidate octaves in [octave(q) - level_below, octave(q) + level_above]
 *   (level_below < 0: from level 0; level_above < 0: no upper bound), queries with an octave outside
 *   [query_level_min, query_level_max] are skipped.
 *   gate 0: SearchForInitialization's rule (a candidate already matched at distance <= dist is skipped; a better
 *           match displaces the earlier one, src/ORBmatcher.cpp:49-50,69-77)
 *   gate 1: first come, first served (a matched F2 keypoint is skipped by later queries; upstream SearchByProjection)
 *   accept: best <= th_dist, and best < (float)best2 * nnratio unless nnratio <= 0
 *   check_orientation: rotation-histogram filter (src/ORBmatcher.cpp:79-118);  update_centers: write the matched
 *   keypoint's position back into d_centers (:121-124). */
typedef struct {
    float radius;
    float level_scale[16];
    int32_t query_level_min, query_level_max;
    int32_t level_below, level_above;
    int32_t gate;
    int32_t th_dist;
    float nnratio;
    int32_t check_orientation;
    int32_t update_centers;
    int32_t width, height;
    int32_t literal_gridid_bug;
    /* Grid bounds of the frames (Frame::FindimageBound, src/Frame.cpp:111-142).  use_bounds == 0: the zero-distortion case
     * [0, width) x [0, height) (:113-118).  use_bounds != 0: the explicit float bounds below -- what FindimageBound
     * computes from the undistorted image corners when the lens is distorted (:121-141). */
    int32_t use_bounds;
    float min_x, max_x, min_y, max_y;
} orbm_window_params;

int orbm_search_window_device(orbm_matcher *m, const orbx_keypoint *d_kps, const uint8_t *d_desc, const int32_t *d_counts,
                              int capacity, const int32_t *d_pair_a, const int32_t *d_pair_b, int npairs,
                              float *d_centers, int32_t *d_matches12, int32_t *d_nmatches,
                              const orbm_window_params *params, void *d_workspace, size_t workspace_bytes, void *stream);
int orbm_search_window_host(orbm_matcher *m, const orbx_keypoint *kp1, const uint8_t *desc1, int n1,
                            const orbx_keypoint *kp2, const uint8_t *desc2, int n2,
                            float *centers, int32_t *matches12, int32_t *nmatches, const orbm_window_params *params);

/* Group-restricted search: the other empty matcher entry point of the reference, SearchByBoW (include/ORBmatcher.h:22), in the
 * form upstream ORB-SLAM2 gives it.  Every keypoint carries a group id (the vocabulary node DBoW2's FeatureVector assigns it;
 * 0xffff = none -- the vocabulary itself is not part of the reference and stays with the caller).  Groups are visited in
 * ascending id and, inside a group, F1's keypoints in ascending index; a query scans the F2 keypoints of its group in
 * ascending index, skipping those already matched, keeps the best two distances (both start at 256), accepts when
 * best <= th_dist and (float)best < nnratio * (float)second, then the rotation-histogram filter.  Same kernel, workspace
 * and output conventions as orbm_search_init_device; d_groups is [nframes][capacity] uint16.  PARITY UNPINNED by the
 * reference (empty body); pinned against this repo's oracle restatement of the upstream loop only. */
int orbm_search_groups_device(orbm_matcher *m, const orbx_keypoint *d_kps, const uint8_t *d_desc, const int32_t *d_counts,
                              const uint16_t *d_groups, int capacity, const int32_t *d_pair_a, const int32_t *d_pair_b, int npairs,
                              int32_t *d_matches12, int32_t *d_nmatches, int th_dist, float nnratio, int check_orientation,
                              void *d_workspace, size_t workspace_bytes, void *stream);
int orbm_search_groups_host(orbm_matcher *m, const orbx_keypoint *kp1, const uint8_t *desc1, const uint16_t *group1, int n1,
                            const orbx_keypoint *kp2, const uint8_t *desc2, const uint16_t *group2, int n2,
                            int32_t *matches12, int32_t *nmatches, int th_dist, float nnratio, int check_orientation);

/* Integer-pipe microbenchmark used for the kNN roofline: runs a dependent-free POPC loop on every
 * SM and returns measured 32-bit POPC results per second (device-event timed). */
int orbm_popc_peak(int device, double *popc_per_second, double *lop3_per_second);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H */

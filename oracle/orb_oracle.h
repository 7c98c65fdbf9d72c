/*
 * orb_oracle.h -- CPU ORACLE for the ORB front-end hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (orbslam_in_practice_b200/) never links or calls it.
 *
 * It restates, in plain C, the reference's algorithm for
 *   ORBextractor::operator()            /root/reference/src/ORBextractor.cpp:1001-1065
 *   ORBmatcher::DescriptorDistance etc. /root/reference/src/ORBmatcher.cpp:9-144
 * and the OpenCV primitives those call (resize, GaussianBlur, FAST, fastAtan2, cvRound), which
 * are NOT vendored in the reference (OpenCV is unpinned there; this oracle is pinned against
 * cv2 4.13.0 by oracle/pin_cv2.py and tests/test_oracle_cv2_pin.py).
 *
 * Parity pin status: the reference has no golden vectors or tests.  The
 * oracle is pinned (a) primitive-by-primitive against real OpenCV 4.13.0 through cv2 and (b)
 * end-to-end against the reference's own ORBextractor.cpp compiled from /root/reference against
 * a header shim (oracle/_ref, built by oracle/Makefile).
 */
#ifndef ORB_ORACLE_H
#define ORB_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBO_MAX_LEVELS 16

/* Same 28-byte layout as cv::KeyPoint (pt.x, pt.y, size, angle, response, octave, class_id). */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbo_keypoint;

/* A FAST candidate in level coordinates relative to (minBorderX, minBorderY) = (16,16). */
typedef struct {
    int16_t x, y;
    int32_t score;
} orbo_cand;

typedef struct orbo_extractor orbo_extractor;

/* ORBextractor ctor, ORBextractor.cpp:360-420 */
orbo_extractor *orbo_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void orbo_destroy(orbo_extractor *ex);

/* tables built by the ctor; each out array has nlevels entries (umax: 16) */
void orbo_tables(const orbo_extractor *ex, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2,
                 int32_t *features_per_level, int32_t *umax);

/* Octree tie-break rule for equal node sizes (ORBextractor.cpp:638 sorts by (size, pointer)):
 *   0 = DEFINED rule: later-created node first (= reference under a monotonic allocator)
 *   1 = opposite rule (earlier-created first), used only to measure sensitivity. */
void orbo_set_tiebreak(orbo_extractor *ex, int rule);

/* operator(), ORBextractor.cpp:1001-1065.  Returns the number of keypoints (<= cap) or -1 if
 * cap is too small.  kps/desc may be NULL to only run the stages. */
int orbo_extract(orbo_extractor *ex, const uint8_t *img, int width, int height, size_t pitch,
                 orbo_keypoint *kps, uint8_t *desc, int cap);

/* stage intermediates of the LAST orbo_extract call */
int orbo_level_dims(const orbo_extractor *ex, int level, int *w, int *h);
const uint8_t *orbo_level_pixels(const orbo_extractor *ex, int level);   /* tight pitch = w */
const uint8_t *orbo_level_blurred(const orbo_extractor *ex, int level);  /* tight pitch = w; NULL if level had no keypoints */
int orbo_level_candidates(const orbo_extractor *ex, int level, const orbo_cand **cands);
int orbo_level_kept(const orbo_extractor *ex, int level, const orbo_cand **kept); /* octree output order */
int orbo_level_retries(const orbo_extractor *ex, int level);             /* cells that fell back to minThFAST */

/* ---- OpenCV primitives restated ---- */
void orbo_resize_linear_u8(const uint8_t *src, int sw, int sh, size_t spitch,
                           uint8_t *dst, int dw, int dh, size_t dpitch);
void orbo_gaussian7_s2_u8(const uint8_t *src, int w, int h, size_t spitch, uint8_t *dst, size_t dpitch);
/* cv::FAST(TYPE_9_16, nonmaxSuppression=true) on a w x h image; output row-major order */
int orbo_fast9_nms(const uint8_t *img, int w, int h, size_t pitch, int threshold, orbo_cand *out, int cap);
/* corner score with threshold 0 (-1 for non corners) for every pixel; border (3px) = -1 */
void orbo_fast9_score0(const uint8_t *img, int w, int h, size_t pitch, int16_t *score, size_t score_pitch);
float orbo_fast_atan2(float y, float x);
int orbo_cv_round(double v);
void orbo_reflect101_border(const uint8_t *src, int w, int h, size_t spitch, uint8_t *dst, int border, size_t dpitch);

/* DistributeOctTree (ORBextractor.cpp:489-718) on candidates; region = [0,width) x [0,height)
 * (i.e. maxX-minX, maxY-minY).  Returns number kept, written in list order. */
int orbo_distribute_octree(const orbo_cand *cands, int n, int width, int height, int N, int tiebreak,
                           orbo_cand *out, int cap);

/* IC_Angle (ORBextractor.cpp:27-54) on a level image; also returns the raw moments */
float orbo_ic_angle(const uint8_t *img, size_t pitch, int x, int y, const int32_t *umax, int32_t *m10, int32_t *m01);
/* computeOrbDescriptor (ORBextractor.cpp:58-97) */
void orbo_orb_descriptor(const uint8_t *blurred, size_t pitch, int x, int y, float angle_deg, uint8_t *desc32);

/* cv::cvtColor(RGB/BGR/RGBA/BGRA -> GRAY) for 8U, the step before the extractor (src/Tracking.cpp:57-70).
 * OpenCV 4.13.0 fixed point: (R*9798 + G*19235 + B*3735 + 16384) >> 15 (pinned against cv2; the 2.4/3.x era used 14 bits). */
void orbo_cvt_gray_u8(const uint8_t *src, int w, int h, size_t spitch, int channels, int rgb_order, uint8_t *dst, size_t dpitch);

/* Frame::UndistortedKeyPoints (src/Frame.cpp:80-109): cv::undistortPoints(pts, K, dist, R = I, P = K) on the keypoint
 * coordinates, restated from OpenCV 4.13.0 (5 fixed iterations in double; pinned against cv2).  cam = {fx, fy, cx, cy},
 * dist = {k1, k2, p1, p2, k3} as floats (the reference stores them in CV_32F Mats, Frame.cpp:46-55).  dist[0] == 0 -> copy
 * (:82-86).  literal_bug != 0 reproduces :106 (the undistorted x is also written to y). */
void orbo_undistort_keypoints(const orbo_keypoint *in, orbo_keypoint *out, int n, const float *cam, const float *dist, int literal_bug);

/* ---- matcher ---- */
/* ORBmatcher::DescriptorDistance, ORBmatcher.cpp:128-144 (SWAR popcount on 8 x int32) */
int orbo_descriptor_distance(const uint8_t *a, const uint8_t *b);
/* best-2 scan of ORBmatcher.cpp:37-62 without the one-to-one gate: for each query scan the db in
 * ascending index; strict '<' so the first minimal index wins; d2 may equal d1.
 * idx1 = index_base + local index; empty db -> d1=d2=INT32_MAX, idx1=-1. */
void orbo_knn2(const uint8_t *q, int nq, const uint8_t *db, int ndb, int index_base,
               int32_t *d1, int32_t *idx1, int32_t *d2);
/* same, OpenMP-free multi-thread helper for the CPU baseline (pthreads over query ranges) */
void orbo_knn2_mt(const uint8_t *q, int nq, const uint8_t *db, int ndb, int index_base,
                  int32_t *d1, int32_t *idx1, int32_t *d2, int nthreads);
/* acceptance of ORBmatcher.cpp:65-67: d1 <= th_low && d1 < (float)d2 * ratio -> idx1 else -1 */
void orbo_ratio_select(const int32_t *d1, const int32_t *idx1, const int32_t *d2, int nq,
                       int th_low, float ratio, int32_t *match);
/* merge of per-shard best-2 triples (shards ordered by ascending index range), SURVEY 8e */
void orbo_merge_shards(const int32_t *d1, const int32_t *idx1, const int32_t *d2, int nshards, int nq,
                       int32_t *od1, int32_t *oidx1, int32_t *od2);

/* Frame grid + SearchForInitialization (Frame.cpp:144-168,219-271; ORBmatcher.cpp:9-126).
 * Grid bounds are the image bounds (no distortion): minX=minY=0, maxX=width, maxY=height.
 * literal_bug != 0 reproduces Frame.cpp:164 (y index computed against miMaxY). */
int orbo_search_for_initialization(const orbo_keypoint *kp1, const uint8_t *desc1, int n1,
                                   const orbo_keypoint *kp2, const uint8_t *desc2, int n2,
                                   float *prev_matched_xy /* n1 x 2, in/out */, int32_t *matches12 /* n1 */,
                                   int window, float nnratio, int check_orientation,
                                   int width, int height, int literal_bug);

/* Windowed search with per-query windows: SearchForInitialization's loop with the gate and acceptance as parameters
 * (same field meaning as orbm_window_params in include/orbx.h).  Gate 0 + SearchForInitialization parameters must equal
 * orbo_search_for_initialization; gate 1 follows upstream ORB-SLAM2's SearchByProjection (reference body empty,
 * include/ORBmatcher.h:24 -> parity unpinned for that instance). */
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
    int32_t use_bounds;                 /* 0: grid bounds [0,width) x [0,height) (Frame.cpp:113-118); else the four floats (:121-141) */
    float min_x, max_x, min_y, max_y;
} orbo_window_params;
int orbo_search_window(const orbo_keypoint *kp1, const uint8_t *desc1, int n1,
                       const orbo_keypoint *kp2, const uint8_t *desc2, int n2,
                       float *centers /* n1 x 2, NaN x = skip */, int32_t *matches12 /* n1 */, const orbo_window_params *params);

/* Group-restricted search (upstream SearchByBoW with the vocabulary node of each keypoint as a uint16 group id, 0xffff = none;
 * reference body empty, include/ORBmatcher.h:22 -> parity unpinned).  Nodes ascending, F1 features of a node ascending, F2
 * features of the node ascending; already matched F2 features are skipped; best two distances start at 256. */
int orbo_search_groups(const orbo_keypoint *kp1, const uint8_t *desc1, const uint16_t *group1, int n1,
                       const orbo_keypoint *kp2, const uint8_t *desc2, const uint16_t *group2, int n2,
                       int32_t *matches12, int th_dist, float nnratio, int check_orientation);

/* multi-thread helper for the CPU baseline: extract `nframes` frames with `nthreads` pthreads,
 * one extractor per thread.  counts[nframes] receives keypoint counts. */
int orbo_extract_many(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST,
                      const uint8_t *imgs, int width, int height, size_t frame_stride, int nframes,
                      int nthreads, int32_t *counts);

#ifdef __cplusplus
}
#endif
#endif

/*
 * ref_match_main.cpp -- driver for the reference's OWN src/ORBmatcher.cpp (SearchForInitialization,
 * DescriptorDistance, ComputeThreeMaxima) compiled unmodified against the shim headers in this directory.
 * TEST INFRASTRUCTURE ONLY; output binary oracle/_ref/ref_match.
 *
 * in.bin : int32 {n1, n2, width, height, window, checkOri, literalBug}, float nnratio,
 *          n1 x 28 B keypoints, n1 x 32 B descriptors, n2 x 28 B keypoints, n2 x 32 B descriptors, n1 x 2 float prevMatched
 * out.bin: int32 nmatches, n1 x int32 matches12, n1 x 2 float prevMatched (updated)
 * optional argv[3..6]: minX maxX minY maxY (grid bounds of a distorted lens); default {0, width, 0, height}
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ORBmatcher.h"

static void read_frame(FILE *f, int n, std::vector<cv::KeyPoint> &kps, cv::Mat &desc)
{
    kps.resize(n);
    std::vector<orbo_keypoint> raw(n);
    if (n && fread(raw.data(), sizeof(orbo_keypoint), n, f) != (size_t)n) exit(3);
    for (int i = 0; i < n; ++i) {
        kps[i].pt.x = raw[i].x; kps[i].pt.y = raw[i].y; kps[i].size = raw[i].size; kps[i].angle = raw[i].angle;
        kps[i].response = raw[i].response; kps[i].octave = raw[i].octave; kps[i].class_id = raw[i].class_id;
    }
    desc.create(n > 0 ? n : 1, 32, CV_8UC1);
    if (n && fread(desc.data, 32, n, f) != (size_t)n) exit(3);
}

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 3;
    int h[7]; float ratio;
    if (fread(h, 4, 7, f) != 7 || fread(&ratio, 4, 1, f) != 1) return 3;
    std::vector<cv::KeyPoint> k1, k2; cv::Mat d1, d2;
    read_frame(f, h[0], k1, d1); read_frame(f, h[1], k2, d2);
    std::vector<cv::Point2f> prev(h[0]);
    for (int i = 0; i < h[0]; ++i) { float xy[2]; if (fread(xy, 4, 2, f) != 2) return 3; prev[i] = cv::Point2f(xy[0], xy[1]); }
    fclose(f);
    float bounds[4] = { 0.f, (float)h[2], 0.f, (float)h[3] };
    if (argc >= 7) for (int i = 0; i < 4; ++i) bounds[i] = (float)atof(argv[3 + i]);
    ORBSlam::Frame F1(k1, d1, bounds, h[6] != 0), F2(k2, d2, bounds, h[6] != 0);
    ORBSlam::ORBmatcher matcher(ratio, h[5] != 0);
    std::vector<int> m12;
    int n = matcher.SearchForInitialization(F1, F2, prev, m12, h[4]);
    FILE *o = fopen(argv[2], "wb");
    fwrite(&n, 4, 1, o);
    fwrite(m12.data(), 4, m12.size(), o);
    for (size_t i = 0; i < prev.size(); ++i) { float xy[2] = { prev[i].x, prev[i].y }; fwrite(xy, 4, 2, o); }
    fclose(o);
    return 0;
}

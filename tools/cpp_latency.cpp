// cpp_latency.cpp -- wall-clock latency of ORBSlam::ORBextractor::operator() on ONE frame per call, the way the reference's
// Frame calls it (src/Frame.cpp:75-78): pageable cv::Mat in, std::vector<cv::KeyPoint> + cv::Mat out.
// usage: cpp_latency W H nfeatures calls device [frame.raw]   -> one JSON line {"p50":..,"p99":..,"mean":..,"keypoints":..}
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ORBextractor.h"

int main(int argc, char **argv)
{
    if (argc < 6) return 2;
    const int W = std::atoi(argv[1]), H = std::atoi(argv[2]), NF = std::atoi(argv[3]), calls = std::atoi(argv[4]), dev = std::atoi(argv[5]);
    std::vector<unsigned char> img((size_t)W * H);
    // textured test image (integer hash noise over blocks): enough corners to fill the quota
    unsigned s = 12345u;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const unsigned b = ((unsigned)(y / 12) * 73856093u) ^ ((unsigned)(x / 12) * 19349663u);
            s = s * 1664525u + 1013904223u;
            img[(size_t)y * W + x] = (unsigned char)(((b >> 7) & 0xff) * 3 / 4 + ((s >> 24) & 7));
        }
    if (argc > 6) {                                     // the bench's own synthetic frame (W*H raw bytes)
        FILE *fi = std::fopen(argv[6], "rb");
        if (!fi || std::fread(img.data(), 1, img.size(), fi) != img.size()) return 3;
        std::fclose(fi);
    }
    try {
        ORBSlam::ORBextractor ex(NF, 1.2f, 8, 20, 7);
        ex.SetDevice(dev);
        cv::Mat im(H, W, CV_8UC1, img.data());
        std::vector<cv::KeyPoint> kps; cv::Mat desc;
        for (int i = 0; i < 30; ++i) ex(im, cv::Mat(), kps, desc);
        std::vector<double> us((size_t)calls);
        for (int i = 0; i < calls; ++i) {
            const auto t0 = std::chrono::steady_clock::now();
            ex(im, cv::Mat(), kps, desc);
            us[i] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        }
        double mean = 0; for (double v : us) mean += v; mean /= calls;
        std::sort(us.begin(), us.end());
        std::printf("{\"p50\": %.1f, \"p99\": %.1f, \"mean\": %.1f, \"keypoints\": %d, \"api\": \"ORBSlam::ORBextractor::operator() (C++ class over the C ABI), pageable cv::Mat\"}\n",
                    us[calls / 2], us[std::min(calls - 1, calls * 99 / 100)], mean, (int)kps.size());
    } catch (const std::exception &e) {
        std::printf("{\"error\": \"%s\"}\n", e.what());
        return 1;
    }
    return 0;
}

/*
 * ref_main.cpp -- driver for the reference's OWN ORBextractor.cpp (compiled unmodified from
 * /root/reference/src against cvshim.h).  TEST INFRASTRUCTURE ONLY; output binary oracle/_ref/ref_orb.
 *
 * Usage: ref_orb <in.bin> <out.bin> [bump|malloc] [repeat]
 *   in.bin : int32 {W, H, nframes, nfeatures, nlevels, iniTh, minTh}, float32 scaleFactor, then frames (u8)
 *   out.bin: per frame int32 n, then n x 28 B keypoints (cv::KeyPoint layout), then n x 32 B descriptors
 *   stdout : "seconds_per_frame <t>" measured over `repeat` passes (extraction only)
 *
 * "bump": global operator new is a monotonic arena, so std::list node addresses increase with
 * creation order and the (size, pointer) sort at ORBextractor.cpp:638 realises the DEFINED
 * tie-break (later-created node first).  "malloc": glibc heap order (history dependent).
 */
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "ORBextractor.h"

static bool g_bump = false;
static char *g_arena = 0;
static size_t g_arena_size = 0, g_arena_off = 0;

void *operator new(size_t n)
{
    if (g_bump) {
        size_t a = (g_arena_off + 15) & ~(size_t)15;
        if (a + n > g_arena_size) { std::fprintf(stderr, "ref_orb: arena exhausted\n"); std::abort(); }
        g_arena_off = a + n;
        return g_arena + a;
    }
    void *p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void operator delete(void *p) noexcept
{
    if (!p) return;
    if (g_arena && (char *)p >= g_arena && (char *)p < g_arena + g_arena_size) return;
    std::free(p);
}
void operator delete(void *p, size_t) noexcept { operator delete(p); }

int main(int argc, char **argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: ref_orb in.bin out.bin [bump|malloc] [repeat]\n"); return 2; }
    bool bump = argc < 4 || std::string(argv[3]) == "bump";
    int repeat = argc > 4 ? std::atoi(argv[4]) : 1;
    FILE *fi = std::fopen(argv[1], "rb");
    if (!fi) return 3;
    int hdr[7]; float sf;
    if (std::fread(hdr, 4, 7, fi) != 7 || std::fread(&sf, 4, 1, fi) != 1) return 3;
    const int W = hdr[0], H = hdr[1], NF = hdr[2];
    std::vector<uchar> frames((size_t)W * H * NF);
    if (std::fread(frames.data(), 1, frames.size(), fi) != frames.size()) return 3;
    std::fclose(fi);
    g_arena_size = (size_t)1 << 30;
    g_arena = (char *)std::malloc(g_arena_size);

    FILE *fo = std::fopen(argv[2], "wb");
    double secs = 0;
    for (int rep = 0; rep < repeat; ++rep) {
        for (int f = 0; f < NF; ++f) {
            cv::Mat img(H, W, CV_8UC1);
            std::memcpy(img.data, frames.data() + (size_t)f * W * H, (size_t)W * H);
            std::vector<orbo_keypoint> okp; std::vector<uchar> odesc; int n = 0;
            {
                g_arena_off = 0; g_bump = bump;
                {
                    ORBSlam::ORBextractor ex(hdr[3], sf, hdr[4], hdr[5], hdr[6]);
                    std::vector<cv::KeyPoint> kps; cv::Mat desc;
                    auto t0 = std::chrono::steady_clock::now();
                    ex(img, cv::Mat(), kps, desc);
                    secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    g_bump = false;
                    n = (int)kps.size();
                    okp.resize(n); odesc.resize((size_t)n * 32);
                    for (int i = 0; i < n; ++i) {
                        orbo_keypoint k = { kps[i].pt.x, kps[i].pt.y, kps[i].size, kps[i].angle, kps[i].response, kps[i].octave, kps[i].class_id };
                        okp[i] = k;
                        std::memcpy(&odesc[(size_t)i * 32], desc.ptr(i), 32);
                    }
                    g_bump = bump; /* destructors of arena objects are no-ops either way */
                }
                g_bump = false;
            }
            if (rep == 0) {
                std::fwrite(&n, 4, 1, fo);
                std::fwrite(okp.data(), sizeof(orbo_keypoint), n, fo);
                std::fwrite(odesc.data(), 1, (size_t)n * 32, fo);
            }
        }
    }
    std::fclose(fo);
    std::printf("seconds_per_frame %.9f\n", secs / ((double)NF * repeat));
    return 0;
}

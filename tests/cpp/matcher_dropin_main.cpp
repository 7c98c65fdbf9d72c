// matcher_dropin_main.cpp -- TEST DRIVER: the reference's matcher call sites, verbatim, against the repo's ORBmatcher.h.
//   mode "init":  src/Tracking.cpp:181-189 -> out.bin in oracle/ref_shim_matcher/ref_match_main.cpp's format, so the
//                 Python test can compare it byte for byte with the reference's own ORBmatcher.cpp (oracle/_ref/ref_match)
//   mode "typed": src/Tracking.cpp:298 (SearchByBoW) and :344 (SearchByProjection) plus ComputeThreeMaxima / DescriptorDistance
// in.bin: int32 {n1, n2, width, height, window, checkOri, literalBug}, float nnratio, n1 x 28 B keypoints, n1 x 32 B
//         descriptors, n2 x 28 B keypoints, n2 x 32 B descriptors, n1 x 2 float prevMatched
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "frame_shim.h"
#include "ORBmatcher.h"

using namespace ORBSlam;

float Frame::sK[9] = { 500.f, 0.f, 320.f, 0.f, 500.f, 240.f, 0.f, 0.f, 1.f };
float Frame::sB[4] = { 0.f, 640.f, 0.f, 480.f };

static void read_frame(FILE *f, int n, std::vector<cv::KeyPoint> &kps, cv::Mat &desc)
{
    kps.resize((size_t)n);
    if (n && std::fread((void *)kps.data(), sizeof(cv::KeyPoint), (size_t)n, f) != (size_t)n) std::exit(3);
    desc.create(n > 0 ? n : 1, 32, CV_8UC1);
    if (n && std::fread(desc.data, 32, (size_t)n, f) != (size_t)n) std::exit(3);
    if (n == 0) desc = cv::Mat();
}

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    const std::string mode = argv[3];
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 3;
    int h[7]; float ratio;
    if (std::fread(h, 4, 7, f) != 7 || std::fread(&ratio, 4, 1, f) != 1) return 3;
    std::vector<cv::KeyPoint> k1, k2; cv::Mat d1, d2;
    read_frame(f, h[0], k1, d1); read_frame(f, h[1], k2, d2);
    std::vector<cv::Point2f> prevIn((size_t)h[0]);
    if (h[0] && std::fread((void *)prevIn.data(), 8, (size_t)h[0], f) != (size_t)h[0]) return 3;
    std::fclose(f);
    Frame::sB[0] = 0.f; Frame::sB[1] = (float)h[2]; Frame::sB[2] = 0.f; Frame::sB[3] = (float)h[3];
    Frame minitialRefFrame(k1, d1), mCurframe(k2, d2);
    FILE *o = std::fopen(argv[2], "wb");
    if (!o) return 3;

    if (mode == "init") {
        std::vector<int> mvIniMatches;
        // ---- src/Tracking.cpp:181-189, verbatim (the ratio / orientation flag come from the test file instead of 0.9, true) ----
        ORBmatcher matcher(ratio, h[5] != 0);
#ifndef FRAME_SHIM_WITH_BOUNDS_ACCESSOR
        matcher.SetImageBounds(0.f, (float)h[2], 0.f, (float)h[3]);
#endif
        matcher.SetLiteralGridIdBug(h[6] != 0);
        std::vector<cv::Point2f> vbPrevMatched;
        vbPrevMatched.resize(minitialRefFrame.GetUnKeyPts().size());
        for (int i=0;i<minitialRefFrame.GetUnKeyPts().size();i++)
        {
            vbPrevMatched[i] = minitialRefFrame.GetUnKeyPts()[i].pt;
        }
        vbPrevMatched = prevIn;                                  // the test's own prev-matched positions
        int matcherCounts = matcher.SearchForInitialization(minitialRefFrame, mCurframe, vbPrevMatched,mvIniMatches,h[4]);
        // ----
        std::fwrite(&matcherCounts, 4, 1, o);
        std::fwrite(mvIniMatches.data(), 4, mvIniMatches.size(), o);
        std::fwrite((const void *)vbPrevMatched.data(), 8, vbPrevMatched.size(), o);
    } else {
        // map points for the keypoints of frame 1: a point at depth 2 that projects back onto its keypoint, every third missing
        std::vector<MapPoint *> owned;
        for (size_t i = 0; i < k1.size(); ++i) {
            if (i % 3 == 2) continue;
            const double z = 2.0;
            const Vec3 p = { { (k1[i].pt.x - Frame::sK[2]) / Frame::sK[0] * z, (k1[i].pt.y - Frame::sK[5]) / Frame::sK[4] * z, z } };
            owned.push_back(new MapPoint(p));
            if (i % 17 == 0) owned.back()->SetBad();
            minitialRefFrame.mvpMappts[i] = owned.back();
        }
        if (k1.size() > 5) minitialRefFrame.mvbOutlier[5] = true;
        Frame &mLastFrame = minitialRefFrame;
        KeyFrame *mpReferenceKF = new KeyFrame(minitialRefFrame);
        std::vector<float> sf(8, 1.f);
        for (int i = 1; i < 8; ++i) sf[i] = (float)(sf[i - 1] * (double)1.2f);
        float th = 7.f;
        int nmatchers;
        {
            // ---- src/Tracking.cpp:333, :344 ----
            ORBmatcher matcher(0.9, true);
            matcher.SetImageBounds(0.f, (float)h[2], 0.f, (float)h[3]);
            matcher.SetScaleFactors(sf);
            nmatchers=matcher.SearchByProjection(mCurframe,mLastFrame,th,true);
        }
        std::vector<int> typedProj(k2.size(), -1);               // current keypoint -> last-frame keypoint owning the stored map point
        for (size_t i2 = 0; i2 < k2.size(); ++i2)
            for (size_t i1 = 0; i1 < k1.size() && mCurframe.mvpMappts[i2]; ++i1)
                if (mLastFrame.mvpMappts[i1] == mCurframe.mvpMappts[i2]) { typedProj[i2] = (int)i1; break; }
        int nmatches;
        std::vector<MapPoint*> vpMapPointMatches;
        {
            // ---- src/Tracking.cpp:293, :298 ----
            ORBmatcher matcher(0.7, true);
            nmatches = matcher.SearchByBoW(mpReferenceKF, mCurframe, vpMapPointMatches);
        }
        std::vector<int> typedBow(k2.size(), -1);
        for (size_t i2 = 0; i2 < vpMapPointMatches.size(); ++i2)
            for (size_t i1 = 0; i1 < k1.size() && vpMapPointMatches[i2]; ++i1)
                if (mLastFrame.mvpMappts[i1] == vpMapPointMatches[i2]) { typedBow[i2] = (int)i1; break; }
        // ComputeThreeMaxima (src/ORBmatcher.cpp:147-188) on a histogram of the keypoint angles, DescriptorDistance on rows 0/1
        ORBmatcher m2;
        std::vector<int> histo[30];
        for (size_t i = 0; i < k1.size(); ++i) histo[((int)(k1[i].angle / 12.f)) % 30].push_back((int)i);
        int ind1 = -1, ind2 = -1, ind3 = -1;
        m2.ComputeThreeMaxima(histo, 30, ind1, ind2, ind3);
        int dd = d1.rows > 1 ? m2.DescriptorDistance(d1.row(0), d1.row(1)) : -1;
        int sizes[30];
        for (int i = 0; i < 30; ++i) sizes[i] = (int)histo[i].size();
        const int n2 = (int)k2.size();
        std::fwrite(&nmatchers, 4, 1, o); std::fwrite(&n2, 4, 1, o); std::fwrite(typedProj.data(), 4, typedProj.size(), o);
        std::fwrite(&nmatches, 4, 1, o); std::fwrite(typedBow.data(), 4, typedBow.size(), o);
        const int tm[4] = { ind1, ind2, ind3, dd };
        std::fwrite(tm, 4, 4, o); std::fwrite(sizes, 4, 30, o);
        for (MapPoint *p : owned) delete p;
        delete mpReferenceKF;
    }
    std::fclose(o);
    return 0;
}

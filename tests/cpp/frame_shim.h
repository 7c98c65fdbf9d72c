// frame_shim.h -- TEST INFRASTRUCTURE.  Stand-ins for the reference's Frame / KeyFrame / MapPoint with exactly the
// PUBLIC members the reference declares (include/Frame.h:30-65, include/KeyFrame.h:13-63, include/MapPoint.h:16-37), so
// that the reference's matcher call sites (src/Tracking.cpp:189, :298, :344) compile verbatim against the repo's
// ORBmatcher.h.  The real headers need Eigen / g2o / the rest of the tree, none of which builds here;
// Vec3 / Mat3 offer Eigen's element access `v(i)` / `m(i, j)`, which is all the matcher templates use.
#pragma once

#include <vector>

#include "cv_compat.h"

namespace ORBSlam {

struct Vec3 { double d[3]; double operator()(int i) const { return d[i]; } };
struct Mat3 { double d[9]; double operator()(int r, int c) const { return d[3 * r + c]; } };

class KeyFrame;

class MapPoint {
public:
    explicit MapPoint(const Vec3 &pos) : mWorldPos(pos), mbBad(false) {}
    bool IsBad() const { return mbBad; }                 // include/MapPoint.h:26
    Vec3 Getpos() const { return mWorldPos; }            // include/MapPoint.h:28-31
    void SetBad() { mbBad = true; }
private:
    Vec3 mWorldPos;
    bool mbBad;
};

class Frame {
public:
    Frame() {}
    Frame(const std::vector<cv::KeyPoint> &kps, const cv::Mat &desc) : mvUnKeypts(kps), mcvDescriptors(desc)
    {
        mvpMappts.assign(kps.size(), static_cast<MapPoint *>(nullptr));
        mvbOutlier.assign(kps.size(), false);
        const Mat3 I = { { 1, 0, 0, 0, 1, 0, 0, 0, 1 } }; const Vec3 z = { { 0, 0, 0 } };
        mRwc = I; mtwc = z;
    }
    std::vector<cv::KeyPoint> GetUnKeyPts() const { return mvUnKeypts; }      // include/Frame.h:30-32 (by value)
    std::vector<cv::KeyPoint> &GetUnKeyPts() { return mvUnKeypts; }            // include/Frame.h:34-37
    cv::Mat GetDescriptors() const { return mcvDescriptors.clone(); }          // include/Frame.h:39-42 (clones)
    Mat3 GetR() const { return mRwc; }                                         // include/Frame.h:44
    Vec3 GetT() const { return mtwc; }                                         // include/Frame.h:48
    void SetPose(const Mat3 &R, const Vec3 &t) { mRwc = R; mtwc = t; }
    static void GetCameraPara(cv::Mat &camk)                                   // include/Frame.h:52
    {
        camk.create(3, 3 * (int)sizeof(float), CV_8UC1);                       // 3 x 3 floats in the byte-typed compat Mat
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) camk.at<float>(r, c) = sK[3 * r + c];
    }
    std::vector<MapPoint *> mvpMappts;                                         // include/Frame.h:64
    std::vector<bool> mvbOutlier;                                              // include/Frame.h:65
#ifdef FRAME_SHIM_WITH_BOUNDS_ACCESSOR
    // the one-line accessor INTEGRATION.md asks a maintainer to add (the bounds are private statics in the reference)
    static void GetImageBounds(float &minX, float &maxX, float &minY, float &maxY) { minX = sB[0]; maxX = sB[1]; minY = sB[2]; maxY = sB[3]; }
#endif
    static float sK[9];
    static float sB[4];
private:
    std::vector<cv::KeyPoint> mvUnKeypts;
    cv::Mat mcvDescriptors;
    Mat3 mRwc; Vec3 mtwc;
};

class KeyFrame {
public:
    explicit KeyFrame(const Frame &f) : mvUnKeypts(f.GetUnKeyPts()), mcvDescriptors(f.GetDescriptors()), mvpMappts(f.mvpMappts) {}
    std::vector<MapPoint *> GetMapPoints() const { return mvpMappts; }         // include/KeyFrame.h:54-57
    std::vector<cv::KeyPoint> mvUnKeypts;                                      // include/KeyFrame.h:62
    cv::Mat mcvDescriptors;                                                    // include/KeyFrame.h:63
private:
    std::vector<MapPoint *> mvpMappts;
};

} // namespace ORBSlam

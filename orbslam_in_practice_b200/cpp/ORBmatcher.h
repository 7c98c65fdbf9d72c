// ORBmatcher.h -- drop-in ORBSlam::ORBmatcher backed by the sm_100a CUDA library (liborbx.so).
//
// Public surface = the reference's include/ORBmatcher.h:8-32, same names and argument meaning:
//   ORBmatcher(float nnratio = 0.6, bool checkOri = true)                                              :11
//   int  SearchForInitialization(Frame&, Frame&, vector<Point2f>&, vector<int>&, int windowSize)       :16-17
//   int  DescriptorDistance(const cv::Mat&, const cv::Mat&)                                            :19
//   void ComputeThreeMaxima(std::vector<int>* histo, const int L, int&, int&, int&)                    :20
//   int  SearchByBoW(KeyFrame*, Frame, std::vector<MapPoint*>&)                                        :22 (empty body there)
//   int  SearchByProjection(Frame&, const Frame&, const float th, const bool bMono)                    :24 (empty body there)
// The Frame / KeyFrame / MapPoint typed entry points are member TEMPLATES over the caller's own classes, so this header
// needs none of the reference's other headers (Eigen, Map, ...) and the call sites src/Tracking.cpp:189, :298, :344, :348
// compile unchanged.  They read the frames only through the reference's public interface (Frame::GetUnKeyPts,
// GetDescriptors, GetR, GetT, GetCameraPara, mvpMappts, mvbOutlier; KeyFrame::mvUnKeypts, mcvDescriptors, GetMapPoints;
// MapPoint::Getpos, IsBad) and forward to the plain-container members below, which call the C ABI (include/orbx.h).
// The one thing the reference's Frame keeps private is the grid bounds (static miMinX.. include/Frame.h:76): give the
// matcher the bounds once with SetImageBounds(), or add the one-line accessor shown in INTEGRATION.md to Frame.h
// (`GetImageBounds(float&, float&, float&, float&)`), which the templates pick up automatically.
#pragma once

#include <cmath>
#include <limits>
#include <stdexcept>
#include <vector>

#include "cv_compat.h"

struct orbm_matcher;

namespace ORBSlam {

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    ~ORBmatcher();
    ORBmatcher(const ORBmatcher &) = delete;
    ORBmatcher &operator=(const ORBmatcher &) = delete;

    // ---------------------------------------------------------------- reference signatures ----
    // include/ORBmatcher.h:16-17, src/ORBmatcher.cpp:9-126.  Call site: src/Tracking.cpp:189.
    template <class FrameT>
    int SearchForInitialization(FrameT &F1, FrameT &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12,
                                int windowSize)
    {
        float b[4];
        FrameBounds(F1, b);
        const std::vector<cv::KeyPoint> &k1 = F1.GetUnKeyPts(), &k2 = F2.GetUnKeyPts();
        return SearchForInitializationBounds(k1, F1.GetDescriptors(), k2, F2.GetDescriptors(), vbPrevMatched, vnMatches12,
                                             windowSize, b[0], b[1], b[2], b[3]);
    }

    // Hamming distance of two 32-byte descriptor rows, src/ORBmatcher.cpp:128-144.  One pair is 8 XOR + popcount on the
    // host: a device round trip per pair would cost ~10 us for 5 ns of work (batches go through BestTwo / the searches).
    int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);

    // src/ORBmatcher.cpp:147-188: indices of the three largest histogram bins (ties: first bin wins; second / third
    // dropped when below 10 % of the largest).  ind1..ind3 are updated in place like the reference (callers pass -1).
    void ComputeThreeMaxima(std::vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3);

    // include/ORBmatcher.h:24 (call sites src/Tracking.cpp:344,348).  The reference body is empty; this follows upstream
    // ORB-SLAM2's frame-to-frame search (PARITY UNPINNED by the reference): every map point of LastFrame that is not an
    // outlier is projected into CurrentFrame with CurrentFrame's pose (x_c = R x_w + t; u = fx x/z + cx), searched in
    // r = th * scaleFactor[octave] among keypoints of octave +-1, best distance <= TH_HIGH, then the rotation check;
    // a match stores the map point in CurrentFrame.mvpMappts.  bMono is accepted and unused (the reference is monocular).
    template <class FrameT>
    int SearchByProjection(FrameT &CurrentFrame, const FrameT &LastFrame, const float th, const bool /*bMono*/)
    {
        float b[4];
        FrameBounds(CurrentFrame, b);
        cv::Mat K;
        FrameT::GetCameraPara(K);
        const float fx = K.template at<float>(0, 0), fy = K.template at<float>(1, 1), cx = K.template at<float>(0, 2), cy = K.template at<float>(1, 2);
        const auto R = CurrentFrame.GetR();
        const auto t = CurrentFrame.GetT();
        const std::vector<cv::KeyPoint> lastKeys = LastFrame.GetUnKeyPts();
        const std::vector<cv::KeyPoint> &curKeys = CurrentFrame.GetUnKeyPts();
        const float nan = std::numeric_limits<float>::quiet_NaN();
        std::vector<cv::Point2f> proj(lastKeys.size(), cv::Point2f(nan, nan));
        for (size_t i = 0; i < lastKeys.size() && i < LastFrame.mvpMappts.size(); ++i) {
            if (!LastFrame.mvpMappts[i] || (i < LastFrame.mvbOutlier.size() && LastFrame.mvbOutlier[i])) continue;
            const auto p = LastFrame.mvpMappts[i]->Getpos();
            const double xc = R(0, 0) * p(0) + R(0, 1) * p(1) + R(0, 2) * p(2) + t(0);
            const double yc = R(1, 0) * p(0) + R(1, 1) * p(1) + R(1, 2) * p(2) + t(1);
            const double zc = R(2, 0) * p(0) + R(2, 1) * p(1) + R(2, 2) * p(2) + t(2);
            if (zc <= 0) continue;
            const float u = (float)(fx * xc / zc + cx), v = (float)(fy * yc / zc + cy);
            if (u < b[0] || u > b[1] || v < b[2] || v > b[3]) continue;
            proj[i] = cv::Point2f(u, v);
        }
        std::vector<int> m;
        const int n = SearchByProjectionBounds(lastKeys, LastFrame.GetDescriptors(), proj, curKeys, CurrentFrame.GetDescriptors(),
                                               mvScaleFactors, m, th, b[0], b[1], b[2], b[3]);
        if (CurrentFrame.mvpMappts.size() < curKeys.size()) CurrentFrame.mvpMappts.resize(curKeys.size(), nullptr);
        for (size_t i = 0; i < m.size(); ++i) if (m[i] >= 0) CurrentFrame.mvpMappts[(size_t)m[i]] = LastFrame.mvpMappts[i];
        return n;
    }

    // include/ORBmatcher.h:22 (call site src/Tracking.cpp:298).  Empty in the reference, which also has no vocabulary
    // (KeyFrame::ComputeBoW / Frame::ComputeBOW are empty, src/KeyFrame.cpp:19, src/Frame.cpp:274-277), so the node of a
    // keypoint comes from SetVocabularyNodes(); without them every keypoint sits in ONE node, i.e. upstream's loop over a
    // single-node vocabulary.  Keyframe keypoints without a good map point do not take part (upstream).
    // vpMatches12 gets one entry per keypoint of F2: the matched map point or NULL.  PARITY UNPINNED by the reference.
    template <class KeyFrameT, class FrameT, class MapPointT>
    int SearchByBoW(KeyFrameT *pKF1, FrameT F2, std::vector<MapPointT *> &vpMatches12)
    {
        const std::vector<cv::KeyPoint> &k1 = pKF1->mvUnKeypts;
        const std::vector<cv::KeyPoint> &k2 = F2.GetUnKeyPts();
        const std::vector<MapPointT *> mps = pKF1->GetMapPoints();
        std::vector<unsigned short> n1(k1.size(), 0xffff), n2(k2.size(), 0);
        for (size_t i = 0; i < k1.size(); ++i)
            if (i < mps.size() && mps[i] && !mps[i]->IsBad()) n1[i] = i < mvNodes1.size() ? mvNodes1[i] : (unsigned short)0;
        for (size_t i = 0; i < k2.size() && i < mvNodes2.size(); ++i) n2[i] = mvNodes2[i];
        std::vector<int> m12;
        const int n = SearchByBoW(k1, pKF1->mcvDescriptors, n1, k2, F2.GetDescriptors(), n2, m12);
        vpMatches12.assign(k2.size(), static_cast<MapPointT *>(nullptr));
        for (size_t i = 0; i < m12.size(); ++i) if (m12[i] >= 0) vpMatches12[(size_t)m12[i]] = mps[i];
        return n;
    }

    // ---------------------------------------------------------------- settings the reference keeps elsewhere ----
    // Grid bounds of the frames = Frame::miMinX/miMaxX/miMinY/miMaxY (private statics, src/Frame.cpp:111-142).
    void SetImageBounds(float minX, float maxX, float minY, float maxY) { mBounds[0] = minX; mBounds[1] = maxX; mBounds[2] = minY; mBounds[3] = maxY; mbHaveBounds = true; }
    // src/Frame.cpp:164 computes the grid row against miMaxY, which leaves the reference's grid empty.
    // Default false = the intended miMinY; true reproduces the reference as written (no candidates, zero matches).
    void SetLiteralGridIdBug(bool on) { mbLiteralGridIdBug = on; }
    // mvScaleFactor of the extractor (ORBextractor::GetScaleFactors()), used by SearchByProjection's window radius
    void SetScaleFactors(const std::vector<float> &v) { mvScaleFactors = v; }
    // vocabulary node per keypoint of the keyframe / the frame for the typed SearchByBoW (0xffff = none)
    void SetVocabularyNodes(const std::vector<unsigned short> &kf, const std::vector<unsigned short> &frame) { mvNodes1 = kf; mvNodes2 = frame; }
    // CUDA device of this matcher (default 0); takes effect when the device handle is (re)created
    void SetDevice(int device);

    // ---------------------------------------------------------------- plain-container forms (what the templates forward to) ----
    // best-2 scan of every query row against every database row, in ascending database order with strict '<'
    // (first minimal index wins).  Outputs have one entry per query; empty database -> (INT_MAX, -1, INT_MAX).
    void BestTwo(const cv::Mat &queries, const cv::Mat &database, std::vector<int> &bestDist, std::vector<int> &bestIdx,
                 std::vector<int> &bestDist2);

    // BestTwo + acceptance: bestDist <= TH_LOW && bestDist < (float)bestDist2 * mfNNratio; returns #matches,
    // vnMatches12[i] = database row or -1
    int SearchBruteForce(const cv::Mat &queries, const cv::Mat &database, std::vector<int> &vnMatches12);

    // SearchForInitialization on keypoints + descriptors; grid bounds [0,width) x [0,height) (src/Frame.cpp:113-118) or explicit
    int SearchForInitialization(const std::vector<cv::KeyPoint> &vKeys1, const cv::Mat &Descriptors1,
                                const std::vector<cv::KeyPoint> &vKeys2, const cv::Mat &Descriptors2,
                                std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize,
                                int imageWidth, int imageHeight);
    int SearchForInitializationBounds(const std::vector<cv::KeyPoint> &vKeys1, const cv::Mat &Descriptors1,
                                      const std::vector<cv::KeyPoint> &vKeys2, const cv::Mat &Descriptors2,
                                      std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize,
                                      float minX, float maxX, float minY, float maxY);

    // upstream frame-to-frame SearchByProjection: last-frame point i, projected to vProjected[i] in the current frame
    // (x = NaN: not visible), searches r = th * vScaleFactors[octave] around it among current keypoints of octave +-1
    // with the reference's GetFeaturesInArea (src/Frame.cpp:219-271); best distance <= TH_HIGH wins, an already matched
    // current keypoint is skipped, then the rotation-histogram filter.  vnMatches[i] = current keypoint index or -1.
    int SearchByProjection(const std::vector<cv::KeyPoint> &vLastKeys, const cv::Mat &LastDescriptors,
                           const std::vector<cv::Point2f> &vProjected,
                           const std::vector<cv::KeyPoint> &vCurrentKeys, const cv::Mat &CurrentDescriptors,
                           const std::vector<float> &vScaleFactors, std::vector<int> &vnMatches, float th,
                           int imageWidth, int imageHeight);
    int SearchByProjectionBounds(const std::vector<cv::KeyPoint> &vLastKeys, const cv::Mat &LastDescriptors,
                                 const std::vector<cv::Point2f> &vProjected,
                                 const std::vector<cv::KeyPoint> &vCurrentKeys, const cv::Mat &CurrentDescriptors,
                                 const std::vector<float> &vScaleFactors, std::vector<int> &vnMatches, float th,
                                 float minX, float maxX, float minY, float maxY);

    // upstream SearchByBoW with the vocabulary node of every keypoint as a uint16 id (0xffff = none): nodes ascending, a
    // node's F1 features ascending, candidates = the node's F2 features not matched yet, best two distances from 256,
    // accept on best <= TH_LOW and best < mfNNratio * second, rotation histogram.  vnMatches12[i] = F2 index or -1.
    int SearchByBoW(const std::vector<cv::KeyPoint> &vKeys1, const cv::Mat &Descriptors1, const std::vector<unsigned short> &vNodes1,
                    const std::vector<cv::KeyPoint> &vKeys2, const cv::Mat &Descriptors2, const std::vector<unsigned short> &vNodes2,
                    std::vector<int> &vnMatches12);

    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

private:
    // bounds of a frame: FrameT::GetImageBounds(minX, maxX, minY, maxY) when the class has it, else SetImageBounds()
    template <class FrameT>
    auto FrameBoundsImpl(FrameT &F, float *b, int) -> decltype(F.GetImageBounds(b[0], b[1], b[2], b[3]), void()) { F.GetImageBounds(b[0], b[1], b[2], b[3]); }
    template <class FrameT>
    void FrameBoundsImpl(FrameT &, float *b, long)
    {
        if (!mbHaveBounds)
            throw std::runtime_error("ORBmatcher: the frame's grid bounds are private in the reference (Frame::miMinX..); call "
                                     "SetImageBounds() or add Frame::GetImageBounds (see INTEGRATION.md)");
        for (int i = 0; i < 4; ++i) b[i] = mBounds[i];
    }
    template <class FrameT> void FrameBounds(FrameT &F, float *b) { FrameBoundsImpl(F, b, 0); }

    void Ensure(int nq, int ndb);
    float mfNNratio;
    bool mbCheckOrientation;
    orbm_matcher *mHandle;
    int mMaxQ, mMaxDb, mDevice;
    bool mbHaveBounds, mbLiteralGridIdBug;
    float mBounds[4];
    std::vector<float> mvScaleFactors;
    std::vector<unsigned short> mvNodes1, mvNodes2;
};

} // namespace ORBSlam

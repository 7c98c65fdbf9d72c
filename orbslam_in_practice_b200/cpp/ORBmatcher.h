// ORBmatcher.h -- ORBSlam::ORBmatcher's distance and candidate-search core on the GPU.
//
// Mirrors include/ORBmatcher.h:8-32 for the hot path: ctor(nnratio, checkOri), the non-static
// DescriptorDistance(a, b) (src/ORBmatcher.cpp:128-144), and the brute-force best-2 candidate scan with
// TH_LOW / ratio acceptance (src/ORBmatcher.cpp:37-67) that SearchForInitialization, SearchByProjection
// and SearchByBoW are built on.  The Frame/KeyFrame-typed entry points stay in the caller's code base
// (SearchByBoW / SearchByProjection have empty bodies in the reference, ORBmatcher.h:22,24; the plain-container forms
// below follow upstream ORB-SLAM2).
#pragma once

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

    // Hamming distance of two 32-byte descriptor rows (one pair per call: kept for API fidelity, slow by nature)
    int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);

    // best-2 scan of every query row against every database row, in ascending database order with strict '<'
    // (first minimal index wins).  Outputs have one entry per query; empty database -> (INT_MAX, -1, INT_MAX).
    void BestTwo(const cv::Mat &queries, const cv::Mat &database, std::vector<int> &bestDist, std::vector<int> &bestIdx,
                 std::vector<int> &bestDist2);

    // BestTwo + acceptance: bestDist <= TH_LOW && bestDist < (float)bestDist2 * mfNNratio; returns #matches,
    // vnMatches12[i] = database row or -1
    int SearchBruteForce(const cv::Mat &queries, const cv::Mat &database, std::vector<int> &vnMatches12);

    // SearchForInitialization (src/ORBmatcher.cpp:9-126) on plain containers: F1 = (vKeys1, Descriptors1),
    // F2 = (vKeys2, Descriptors2), undistorted keypoints, grid bounds [0,width) x [0,height) (src/Frame.cpp:113-118).
    // The reference's Frame-typed overload forwards here with F.GetUnKeyPts() / F.GetDescriptors().
    int SearchForInitialization(const std::vector<cv::KeyPoint> &vKeys1, const cv::Mat &Descriptors1,
                                const std::vector<cv::KeyPoint> &vKeys2, const cv::Mat &Descriptors2,
                                std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize,
                                int imageWidth, int imageHeight);

    // SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, float th, bool bMono) (include/ORBmatcher.h:24) on
    // plain containers.  The reference's body is empty, so the loop follows upstream ORB-SLAM2's frame-to-frame variant
    // (SURVEY.md 8f row 4): last-frame point i, projected to vProjected[i] in the current frame (x = NaN: not visible),
    // searches r = th * vScaleFactors[octave] around it among current keypoints of octave +-1 with the reference's
    // GetFeaturesInArea (src/Frame.cpp:219-271); best distance <= TH_HIGH wins, an already matched current keypoint is
    // skipped, then the rotation-histogram filter.  vnMatches[i] = current keypoint index or -1; returns #matches.
    int SearchByProjection(const std::vector<cv::KeyPoint> &vLastKeys, const cv::Mat &LastDescriptors,
                           const std::vector<cv::Point2f> &vProjected,
                           const std::vector<cv::KeyPoint> &vCurrentKeys, const cv::Mat &CurrentDescriptors,
                           const std::vector<float> &vScaleFactors, std::vector<int> &vnMatches, float th,
                           int imageWidth, int imageHeight);

    // SearchByBoW(KeyFrame* pKF1, Frame F2, std::vector<MapPoint*>& vpMatches12) (include/ORBmatcher.h:22) on plain
    // containers.  The reference's body is empty and it has no vocabulary, so the caller supplies the vocabulary node of
    // every keypoint (DBoW2's FeatureVector, as a uint16 id; 0xffff = none) and the loop follows upstream ORB-SLAM2:
    // nodes ascending, a node's F1 features ascending, candidates = the node's F2 features not matched yet, best two
    // distances from 256, accept on best <= TH_LOW and best < mfNNratio * second, rotation histogram.
    // vnMatches12[i] = F2 keypoint index or -1; returns #matches.
    int SearchByBoW(const std::vector<cv::KeyPoint> &vKeys1, const cv::Mat &Descriptors1, const std::vector<unsigned short> &vNodes1,
                    const std::vector<cv::KeyPoint> &vKeys2, const cv::Mat &Descriptors2, const std::vector<unsigned short> &vNodes2,
                    std::vector<int> &vnMatches12);

    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;

private:
    void Ensure(int nq, int ndb);
    float mfNNratio;
    bool mbCheckOrientation;
    orbm_matcher *mHandle;
    int mMaxQ, mMaxDb;
};

} // namespace ORBSlam

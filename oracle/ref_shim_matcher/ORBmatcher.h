/*
 * ORBmatcher.h (shim) -- same class declaration as the reference's include/ORBmatcher.h:8-32 minus the two
 * KeyFrame/MapPoint-typed stubs with empty bodies (:22,:24), on top of the Frame shim.  Lets the reference's
 * src/ORBmatcher.cpp compile unmodified.  TEST INFRASTRUCTURE ONLY.
 */
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include "Frame.h"

namespace ORBSlam {
class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    int SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize);
    int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
    void ComputeThreeMaxima(std::vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3);

private:
    static const int HISTO_LENGTH;
    static const int TH_LOW;
    float mfNNratio;
    bool mbCheckOrientation;
};
} // namespace ORBSlam
#endif

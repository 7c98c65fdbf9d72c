/*
 * Declaration shim for the matcher translation unit.  TEST INFRASTRUCTURE ONLY.
 *
 * oracle/Makefile compiles the reference's src/ORBmatcher.cpp where it lies; that file defines four members
 * of ORBSlam::ORBmatcher (constructor inline in the header, SearchForInitialization, DescriptorDistance,
 * ComputeThreeMaxima) and two static constants.  This header declares exactly those so the definitions have
 * something to attach to.  The KeyFrame / MapPoint entry points of the reference header (include/ORBmatcher.h:22,24)
 * have empty bodies there and are not on the measured path, so they are not declared.
 *
 * Member order and the two data members' types must match what the .cpp's definitions use; nothing else here
 * is taken from the reference header.
 */
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include "Frame.h"

namespace ORBSlam {

class ORBmatcher {
    /* state the definitions read: ratio of the best-2 test and the rotation-histogram switch */
    float mfNNratio;
    bool mbCheckOrientation;

    /* defined at the top of the .cpp */
    static const int TH_LOW;
    static const int HISTO_LENGTH;

public:
    typedef std::vector<int> IntVec;
    typedef std::vector<cv::Point2f> PointVec;

    explicit ORBmatcher(float ratio = 0.6, bool useOrientation = true)
    {
        mfNNratio = ratio;
        mbCheckOrientation = useOrientation;
    }

    /* 32-byte Hamming distance */
    int DescriptorDistance(const cv::Mat &lhs, const cv::Mat &rhs);

    /* three largest bins of a rotation histogram with `bins` entries */
    void ComputeThreeMaxima(IntVec *histogram, const int bins, int &first, int &second, int &third);

    /* windowed best-2 search between the initial and the current frame */
    int SearchForInitialization(Frame &initial, Frame &current, PointVec &prevMatched, IntVec &matches12, int window);
};

} // namespace ORBSlam
#endif

// ORBextractor.h -- drop-in ORBSlam::ORBextractor backed by the sm_100a CUDA library (liborbx.so).
//
// Public surface = the reference's include/ORBextractor.h:29-71 (same names, argument meaning and
// error behaviour): ctor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST), operator()(image,
// mask, keypoints, descriptors) with the mask ignored, Get* accessors returning by value, and the
// public mvImagePyramid member.  Everything per-pixel happens on the GPU through the C ABI in
// include/orbx.h; there is no CPU fallback (construction throws std::runtime_error without a B200).
// mvImagePyramid has no reader anywhere in the reference, so the pyramid stays on the device and the
// member is materialised on request: DownloadPyramid() after a call, or SetPyramidDownload(true) for every call.
// Additions (not in the reference): ExtractBatch() for many same-sized frames per call, SetDevice().
#pragma once

#include <vector>

#include "cv_compat.h"

struct orbx_extractor;

namespace ORBSlam {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor &) = delete;
    ORBextractor &operator=(const ORBextractor &) = delete;

    // Compute the ORB features and descriptors on an image (8-bit, single channel).
    // ORB are dispersed on the image using an octree.  Mask is ignored (as in the reference).
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint> &keypoints,
                    cv::OutputArray descriptors);

    int GetLevels() { return nlevels; }
    float GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // level l is a view of (cols x rows) pixels inside a buffer carrying the 19-px reflect-101 border.  Filled by
    // DownloadPyramid() (the last operator() call's pyramid stays on the device until the next call) or, with
    // SetPyramidDownload(true), by every operator() call as in the reference.
    std::vector<cv::Mat> mvImagePyramid;
    void DownloadPyramid();

    // ---- extensions ----
    // nframes same-sized frames: frame f, row y at imgs + f*frameStride + y*rowPitch
    void ExtractBatch(const unsigned char *imgs, int width, int height, size_t rowPitch, size_t frameStride, int nframes,
                      std::vector<std::vector<cv::KeyPoint> > &keypoints, std::vector<cv::Mat> &descriptors);
    void SetPyramidDownload(bool on) { mbDownloadPyramid = on; }
    void SetDevice(int device);

protected:
    void EnsureHandle(int width, int height, int batch);

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<int> umax;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

    orbx_extractor *mHandle;
    int mDevice, mMaxW, mMaxH, mMaxBatch;
    bool mbDownloadPyramid;
    std::vector<cv::Mat> mvBordered;   // owners of the bordered level buffers
    // host staging of one call's outputs, sized once per handle: no allocation on the per-frame path
    std::vector<unsigned char> mKpsBuf, mDescBuf;
};

} // namespace ORBSlam

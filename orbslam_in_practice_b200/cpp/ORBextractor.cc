// ORBextractor.cc -- host-side mirror of the reference class over the C ABI (see ORBextractor.h).
#include "ORBextractor.h"

#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/orbx.h"

namespace ORBSlam {

static void check(int rc, const char *what)
{
    if (rc != ORBX_OK)
        throw std::runtime_error(std::string("ORBextractor: ") + what + ": " + orbx_strerror(rc) +
                                 (rc == ORBX_E_CUDA ? std::string(" [") + orbx_last_cuda_error() + "]" : std::string()));
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
      mHandle(nullptr), mDevice(0), mMaxW(0), mMaxH(0), mMaxBatch(0), mbDownloadPyramid(false)
{
    // the tables come from the library (it restates ORBextractor.cpp:360-420); a 64x64 handle is enough
    // to query them and verifies early that a device exists (no CPU fallback)
    EnsureHandle(64, 64, 1);
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels); umax.resize(16);
    check(orbx_tables(mHandle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                      mnFeaturesPerLevel.data(), umax.data()), "tables");
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() { if (mHandle) orbx_destroy(mHandle); }

void ORBextractor::EnsureHandle(int width, int height, int batch)
{
    if (mHandle && width <= mMaxW && height <= mMaxH && batch <= mMaxBatch) return;
    if (mHandle) { orbx_destroy(mHandle); mHandle = nullptr; }
    orbx_params p = { nfeatures, (float)scaleFactor, nlevels, iniThFAST, minThFAST };
    mMaxW = width > mMaxW ? width : mMaxW; mMaxH = height > mMaxH ? height : mMaxH; mMaxBatch = batch > mMaxBatch ? batch : mMaxBatch;
    check(orbx_create(&p, mMaxW, mMaxH, mMaxBatch, mDevice, &mHandle), "create");
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*mask*/, std::vector<cv::KeyPoint> &_keypoints,
                              cv::OutputArray _descriptors)
{
    if (_image.empty()) return;                                   // reference: silent return, outputs untouched
    cv::Mat image = _image.getMat();
    if (image.type() != CV_8UC1) throw std::runtime_error("ORBextractor: image must be CV_8UC1");
    EnsureHandle(image.cols, image.rows, 1);
    const size_t cap = (size_t)orbx_capacity(mHandle);
    if (mKpsBuf.size() < cap * sizeof(orbx_keypoint)) { mKpsBuf.resize(cap * sizeof(orbx_keypoint)); mDescBuf.resize(cap * 32); }
    int count = 0;
    check(orbx_extract_host(mHandle, image.data, image.step, image.step * (size_t)image.rows, image.cols, image.rows, 1,
                            (orbx_keypoint *)mKpsBuf.data(), mDescBuf.data(), &count), "extract");
    if (count == 0) _descriptors.release();                       // src/ORBextractor.cpp:1024-1030
    else {
        _descriptors.create(count, 32, CV_8U);
        cv::Mat d = _descriptors.getMat();
        if (d.isContinuous()) std::memcpy(d.ptr(0), mDescBuf.data(), (size_t)count * 32);
        else for (int i = 0; i < count; ++i) std::memcpy(d.ptr(i), &mDescBuf[(size_t)i * 32], 32);
    }
    _keypoints.resize((size_t)count);                             // :1032-1033 (clear + reserve, then filled)
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "KeyPoint layout");
    if (count) std::memcpy((void *)_keypoints.data(), mKpsBuf.data(), (size_t)count * sizeof(orbx_keypoint));
    if (mbDownloadPyramid) DownloadPyramid();
}

void ORBextractor::DownloadPyramid()
{
    if (!mHandle) return;
    mvBordered.resize(nlevels);
    for (int l = 0; l < nlevels; ++l) {
        int w = 0, h = 0;
        check(orbx_level_dims(mHandle, l, &w, &h), "level_dims");
        const int B = ORBX_EDGE_THRESHOLD;
        mvBordered[l].create(h + 2 * B, w + 2 * B, CV_8UC1);
        check(orbx_download_level(mHandle, 0, l, 0, B, mvBordered[l].data, mvBordered[l].step), "download_level");
        mvImagePyramid[l] = mvBordered[l](cv::Rect(B, B, w, h));
    }
}

void ORBextractor::SetDevice(int device)
{
    if (device == mDevice) return;
    if (mHandle) { orbx_destroy(mHandle); mHandle = nullptr; mMaxW = mMaxH = mMaxBatch = 0; }   // re-created on the new device by the next call
    mDevice = device;
}

void ORBextractor::ExtractBatch(const unsigned char *imgs, int width, int height, size_t rowPitch, size_t frameStride, int nframes,
                                std::vector<std::vector<cv::KeyPoint> > &keypoints, std::vector<cv::Mat> &descriptors)
{
    keypoints.assign((size_t)nframes, std::vector<cv::KeyPoint>());
    descriptors.assign((size_t)nframes, cv::Mat());
    if (nframes <= 0 || !imgs || width <= 0 || height <= 0) return;
    EnsureHandle(width, height, nframes);
    const size_t cap = (size_t)orbx_capacity(mHandle);
    std::vector<orbx_keypoint> kps(cap * nframes);
    std::vector<unsigned char> desc(cap * 32 * nframes);
    std::vector<int> counts((size_t)nframes);
    check(orbx_extract_host(mHandle, imgs, rowPitch, frameStride, width, height, nframes, kps.data(), desc.data(), counts.data()), "extract");
    for (int f = 0; f < nframes; ++f) {
        const int n = counts[f];
        keypoints[f].resize((size_t)n);
        if (n) {
            std::memcpy((void *)keypoints[f].data(), &kps[cap * f], (size_t)n * sizeof(orbx_keypoint));
            descriptors[f].create(n, 32, CV_8U);
            for (int i = 0; i < n; ++i) std::memcpy(descriptors[f].ptr(i), &desc[(cap * f + i) * 32], 32);
        }
    }
}

} // namespace ORBSlam

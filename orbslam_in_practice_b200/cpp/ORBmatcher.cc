// ORBmatcher.cc -- host-side mirror of the reference matcher over the C ABI (see ORBmatcher.h).
#include "ORBmatcher.h"

#include <climits>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/orbx.h"

namespace ORBSlam {

const int ORBmatcher::TH_LOW = 50;          // src/ORBmatcher.cpp:7
const int ORBmatcher::TH_HIGH = 100;        // upstream ORB-SLAM2 (the reference defines TH_LOW only)
const int ORBmatcher::HISTO_LENGTH = 30;    // src/ORBmatcher.cpp:6

static void check(int rc, const char *what)
{
    if (rc != ORBX_OK)
        throw std::runtime_error(std::string("ORBmatcher: ") + what + ": " + orbx_strerror(rc) +
                                 ((rc == ORBX_E_CUDA || rc == ORBX_E_UNSUPPORTED) ? std::string(" [") + orbx_last_cuda_error() + "]" : std::string()));
}

ORBmatcher::ORBmatcher(float nnratio, bool checkOri)
    : mfNNratio(nnratio), mbCheckOrientation(checkOri), mHandle(nullptr), mMaxQ(0), mMaxDb(0), mDevice(0),
      mbHaveBounds(false), mbLiteralGridIdBug(false)
{
    mBounds[0] = mBounds[1] = mBounds[2] = mBounds[3] = 0.f;
}

ORBmatcher::~ORBmatcher() { if (mHandle) orbm_destroy(mHandle); }

void ORBmatcher::SetDevice(int device)
{
    if (device == mDevice) return;
    if (mHandle) { orbm_destroy(mHandle); mHandle = nullptr; }   // re-created on the new device by the next call
    mDevice = device;
}

void ORBmatcher::Ensure(int nq, int ndb)
{
    if (mHandle && nq <= mMaxQ && ndb <= mMaxDb) return;
    if (mHandle) { orbm_destroy(mHandle); mHandle = nullptr; }
    mMaxQ = nq > mMaxQ ? nq : mMaxQ; mMaxDb = ndb > mMaxDb ? ndb : mMaxDb;
    if (mMaxQ < 1) mMaxQ = 1;
    check(orbm_create(mMaxQ, mMaxDb, mDevice, &mHandle), "create");
}

// n x 32 descriptor rows of a Mat as one dense byte array (a Mat may be a row-strided view).  Every entry point comes
// through here, so a descriptor matrix that does not fit the keypoint list is refused before anything reaches the GPU.
static std::vector<unsigned char> pack_rows(const cv::Mat &m, int expect_rows, const char *what)
{
    if (expect_rows == 0 && m.rows == 0) return std::vector<unsigned char>();
    if (m.type() != CV_8UC1 || m.cols != 32)
        throw std::runtime_error(std::string("ORBmatcher: ") + what + " must be an N x 32 CV_8UC1 descriptor matrix");
    if (expect_rows >= 0 && m.rows != expect_rows)
        throw std::runtime_error(std::string("ORBmatcher: ") + what + " must have one row per keypoint");
    std::vector<unsigned char> v((size_t)m.rows * 32);
    for (int r = 0; r < m.rows; ++r) std::memcpy(&v[(size_t)r * 32], m.ptr(r), 32);
    return v;
}

int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    if (a.cols < 32 || b.cols < 32 || a.empty() || b.empty()) throw std::runtime_error("ORBmatcher: DescriptorDistance needs two 32-byte rows");
    uint64_t x[4], y[4];                       // memcpy: rows of a strided Mat need not be 8-byte aligned
    std::memcpy(x, a.ptr(0), 32); std::memcpy(y, b.ptr(0), 32);
    int dist = 0;
    for (int i = 0; i < 4; ++i) dist += __builtin_popcountll(x[i] ^ y[i]);
    return dist;
}

void ORBmatcher::ComputeThreeMaxima(std::vector<int> *histo, const int L, int &ind1, int &ind2, int &ind3)
{
    // three running maxima over the bin sizes; a strict '>' keeps the first of equal bins in front
    int top[3] = { 0, 0, 0 };
    int *ind[3] = { &ind1, &ind2, &ind3 };
    for (int i = 0; i < L; ++i) {
        const int s = (int)histo[i].size();
        int k = 0;
        while (k < 3 && s <= top[k]) ++k;
        if (k == 3) continue;
        for (int j = 2; j > k; --j) { top[j] = top[j - 1]; *ind[j] = *ind[j - 1]; }
        top[k] = s; *ind[k] = i;
    }
    if ((float)top[1] < 0.1f * (float)top[0]) { ind2 = -1; ind3 = -1; }
    else if ((float)top[2] < 0.1f * (float)top[0]) ind3 = -1;
}

void ORBmatcher::BestTwo(const cv::Mat &queries, const cv::Mat &database, std::vector<int> &bestDist, std::vector<int> &bestIdx,
                         std::vector<int> &bestDist2)
{
    const int nq = queries.rows, ndb = database.rows;
    bestDist.assign((size_t)nq, INT_MAX); bestIdx.assign((size_t)nq, -1); bestDist2.assign((size_t)nq, INT_MAX);
    if (nq == 0) return;
    const std::vector<unsigned char> q = pack_rows(queries, -1, "queries"), d = pack_rows(database, ndb ? -1 : 0, "database");
    Ensure(nq, ndb);
    check(orbm_knn2_host(mHandle, q.data(), nq, d.data(), ndb, 0, bestDist.data(), bestIdx.data(), bestDist2.data()), "knn2");
}

int ORBmatcher::SearchForInitialization(const std::vector<cv::KeyPoint> &vKeys1, const cv::Mat &Descriptors1,
                                        const std::vector<cv::KeyPoint> &vKeys2, const cv::Mat &Descriptors2,
                                        std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize,
                                        int imageWidth, int imageHeight)
{
    // zero-distortion bounds, src/Frame.cpp:113-118
    return SearchForInitializationBounds(vKeys1, Descriptors1, vKeys2, Descriptors2, vbPrevMatched, vnMatches12, windowSize,
                                         0.f, (float)imageWidth, 0.f, (float)imageHeight);
}

static void set_bounds(orbm_window_params &p, float minX, float maxX, float minY, float maxY)
{
    p.use_bounds = 1; p.min_x = minX; p.max_x = maxX; p.min_y = minY; p.max_y = maxY;
    p.width = 1; p.height = 1;
}

int ORBmatcher::SearchForInitializationBounds(const std::vector<cv::KeyPoint> &vKeys1, const cv::Mat &Descriptors1,
                                              const std::vector<cv::KeyPoint> &vKeys2, const cv::Mat &Descriptors2,
                                              std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize,
                                              float minX, float maxX, float minY, float maxY)
{
    const int n1 = (int)vKeys1.size(), n2 = (int)vKeys2.size();
    vnMatches12.assign((size_t)n1, -1);                                       // src/ORBmatcher.cpp:13
    if (n1 == 0) return 0;
    if ((int)vbPrevMatched.size() != n1) throw std::runtime_error("ORBmatcher: vbPrevMatched must have one entry per keypoint of F1");
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint) && sizeof(cv::Point2f) == 8, "layout");
    const std::vector<unsigned char> d1 = pack_rows(Descriptors1, n1, "Descriptors1"), d2 = pack_rows(Descriptors2, n2, "Descriptors2");
    Ensure(n1 > n2 ? n1 : n2, n1 > n2 ? n1 : n2);
    // SearchForInitialization as an instance of the windowed search (src/ORBmatcher.cpp:24-33,49-50,65-77): octave-0
    // queries and candidates, fixed window, TH_LOW, displacement gate, prev-matched update
    orbm_window_params p = orbm_window_params();
    p.radius = (float)windowSize;
    for (int i = 0; i < 16; ++i) p.level_scale[i] = 1.0f;
    p.query_level_min = 0; p.query_level_max = 0; p.level_below = 0; p.level_above = 0;
    p.gate = 0; p.th_dist = TH_LOW; p.nnratio = mfNNratio; p.check_orientation = mbCheckOrientation ? 1 : 0; p.update_centers = 1;
    p.literal_gridid_bug = mbLiteralGridIdBug ? 1 : 0;
    set_bounds(p, minX, maxX, minY, maxY);
    int nmatches = 0;
    check(orbm_search_window_host(mHandle, (const orbx_keypoint *)vKeys1.data(), d1.data(), n1,
                                  (const orbx_keypoint *)vKeys2.data(), d2.data(), n2, (float *)vbPrevMatched.data(),
                                  vnMatches12.data(), &nmatches, &p), "search_init");
    if (nmatches < 0) throw std::runtime_error("ORBmatcher: search workspace too small");
    return nmatches;
}

int ORBmatcher::SearchByProjection(const std::vector<cv::KeyPoint> &vLastKeys, const cv::Mat &LastDescriptors,
                                   const std::vector<cv::Point2f> &vProjected,
                                   const std::vector<cv::KeyPoint> &vCurrentKeys, const cv::Mat &CurrentDescriptors,
                                   const std::vector<float> &vScaleFactors, std::vector<int> &vnMatches, float th,
                                   int imageWidth, int imageHeight)
{
    return SearchByProjectionBounds(vLastKeys, LastDescriptors, vProjected, vCurrentKeys, CurrentDescriptors, vScaleFactors, vnMatches, th,
                                    0.f, (float)imageWidth, 0.f, (float)imageHeight);
}

int ORBmatcher::SearchByProjectionBounds(const std::vector<cv::KeyPoint> &vLastKeys, const cv::Mat &LastDescriptors,
                                         const std::vector<cv::Point2f> &vProjected,
                                         const std::vector<cv::KeyPoint> &vCurrentKeys, const cv::Mat &CurrentDescriptors,
                                         const std::vector<float> &vScaleFactors, std::vector<int> &vnMatches, float th,
                                         float minX, float maxX, float minY, float maxY)
{
    const int n1 = (int)vLastKeys.size(), n2 = (int)vCurrentKeys.size();
    vnMatches.assign((size_t)n1, -1);
    if (n1 == 0) return 0;
    if ((int)vProjected.size() != n1) throw std::runtime_error("ORBmatcher: vProjected must have one entry per last-frame keypoint");
    const std::vector<unsigned char> d1 = pack_rows(LastDescriptors, n1, "LastDescriptors"), d2 = pack_rows(CurrentDescriptors, n2, "CurrentDescriptors");
    Ensure(n1 > n2 ? n1 : n2, n1 > n2 ? n1 : n2);
    orbm_window_params p = orbm_window_params();
    p.radius = th;
    for (int i = 0; i < 16; ++i) p.level_scale[i] = i < (int)vScaleFactors.size() ? vScaleFactors[(size_t)i] : 1.0f;
    p.query_level_min = 0; p.query_level_max = 15; p.level_below = 1; p.level_above = 1;
    p.gate = 1; p.th_dist = TH_HIGH; p.nnratio = 0.0f; p.check_orientation = mbCheckOrientation ? 1 : 0; p.update_centers = 0;
    p.literal_gridid_bug = mbLiteralGridIdBug ? 1 : 0;
    set_bounds(p, minX, maxX, minY, maxY);
    std::vector<cv::Point2f> centers(vProjected);
    int nmatches = 0;
    check(orbm_search_window_host(mHandle, (const orbx_keypoint *)vLastKeys.data(), d1.data(), n1,
                                  (const orbx_keypoint *)vCurrentKeys.data(), d2.data(), n2, (float *)centers.data(),
                                  vnMatches.data(), &nmatches, &p), "search_window");
    if (nmatches < 0) throw std::runtime_error("ORBmatcher: search workspace too small");
    return nmatches;
}

int ORBmatcher::SearchByBoW(const std::vector<cv::KeyPoint> &vKeys1, const cv::Mat &Descriptors1, const std::vector<unsigned short> &vNodes1,
                            const std::vector<cv::KeyPoint> &vKeys2, const cv::Mat &Descriptors2, const std::vector<unsigned short> &vNodes2,
                            std::vector<int> &vnMatches12)
{
    const int n1 = (int)vKeys1.size(), n2 = (int)vKeys2.size();
    vnMatches12.assign((size_t)n1, -1);
    if (n1 == 0) return 0;
    if ((int)vNodes1.size() != n1 || (int)vNodes2.size() != n2) throw std::runtime_error("ORBmatcher: one vocabulary node per keypoint expected");
    const std::vector<unsigned char> d1 = pack_rows(Descriptors1, n1, "Descriptors1"), d2 = pack_rows(Descriptors2, n2, "Descriptors2");
    Ensure(n1 > n2 ? n1 : n2, n1 > n2 ? n1 : n2);
    int nmatches = 0;
    check(orbm_search_groups_host(mHandle, (const orbx_keypoint *)vKeys1.data(), d1.data(), vNodes1.data(), n1,
                                  (const orbx_keypoint *)vKeys2.data(), d2.data(), vNodes2.data(), n2,
                                  vnMatches12.data(), &nmatches, TH_LOW, mfNNratio, mbCheckOrientation ? 1 : 0), "search_groups");
    if (nmatches < 0) throw std::runtime_error("ORBmatcher: search workspace too small");
    return nmatches;
}

int ORBmatcher::SearchBruteForce(const cv::Mat &queries, const cv::Mat &database, std::vector<int> &vnMatches12)
{
    std::vector<int> d1, i1, d2;
    BestTwo(queries, database, d1, i1, d2);
    vnMatches12.assign(d1.size(), -1);
    int nmatches = 0;
    for (size_t i = 0; i < d1.size(); ++i)                           // src/ORBmatcher.cpp:65-67
        if (i1[i] >= 0 && d1[i] <= TH_LOW && d1[i] < (float)d2[i] * mfNNratio) { vnMatches12[i] = i1[i]; ++nmatches; }
    return nmatches;
}

} // namespace ORBSlam

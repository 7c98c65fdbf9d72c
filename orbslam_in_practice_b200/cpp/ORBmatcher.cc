// ORBmatcher.cc -- host-side mirror of the reference matcher core over the C ABI (see ORBmatcher.h).
#include "ORBmatcher.h"

#include <climits>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/orbx.h"

namespace ORBSlam {

const int ORBmatcher::TH_LOW = 50;          // src/ORBmatcher.cpp:7
const int ORBmatcher::TH_HIGH = 100;        // upstream ORB-SLAM2 (the reference defines TH_LOW only)
const int ORBmatcher::HISTO_LENGTH = 30;    // src/ORBmatcher.cpp:6

static void check(int rc, const char *what)
{
    if (rc != ORBX_OK) throw std::runtime_error(std::string("ORBmatcher: ") + what + ": " + orbx_strerror(rc));
}

ORBmatcher::ORBmatcher(float nnratio, bool checkOri)
    : mfNNratio(nnratio), mbCheckOrientation(checkOri), mHandle(nullptr), mMaxQ(0), mMaxDb(0) {}

ORBmatcher::~ORBmatcher() { if (mHandle) orbm_destroy(mHandle); }

void ORBmatcher::Ensure(int nq, int ndb)
{
    if (mHandle && nq <= mMaxQ && ndb <= mMaxDb) return;
    if (mHandle) { orbm_destroy(mHandle); mHandle = nullptr; }
    mMaxQ = nq > mMaxQ ? nq : mMaxQ; mMaxDb = ndb > mMaxDb ? ndb : mMaxDb;
    if (mMaxQ < 1) mMaxQ = 1;
    check(orbm_create(mMaxQ, mMaxDb, 0, &mHandle), "create");
}

static std::vector<unsigned char> pack_rows(const cv::Mat &m)
{
    std::vector<unsigned char> v((size_t)m.rows * 32);
    for (int r = 0; r < m.rows; ++r) std::memcpy(&v[(size_t)r * 32], m.ptr(r), 32);
    return v;
}

int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
{
    Ensure(1, 1);
    int dist = 0;
    check(orbm_hamming_pairs_host(mHandle, a.ptr(0), b.ptr(0), 1, &dist), "hamming");
    return dist;
}

void ORBmatcher::BestTwo(const cv::Mat &queries, const cv::Mat &database, std::vector<int> &bestDist, std::vector<int> &bestIdx,
                         std::vector<int> &bestDist2)
{
    const int nq = queries.rows, ndb = database.rows;
    bestDist.assign((size_t)nq, INT_MAX); bestIdx.assign((size_t)nq, -1); bestDist2.assign((size_t)nq, INT_MAX);
    if (nq == 0) return;
    if (queries.cols != 32 || (ndb && database.cols != 32)) throw std::runtime_error("ORBmatcher: descriptors must be N x 32 bytes");
    Ensure(nq, ndb);
    const std::vector<unsigned char> q = pack_rows(queries), d = pack_rows(database);
    check(orbm_knn2_host(mHandle, q.data(), nq, d.data(), ndb, 0, bestDist.data(), bestIdx.data(), bestDist2.data()), "knn2");
}

int ORBmatcher::SearchForInitialization(const std::vector<cv::KeyPoint> &vKeys1, const cv::Mat &Descriptors1,
                                        const std::vector<cv::KeyPoint> &vKeys2, const cv::Mat &Descriptors2,
                                        std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize,
                                        int imageWidth, int imageHeight)
{
    const int n1 = (int)vKeys1.size(), n2 = (int)vKeys2.size();
    vnMatches12.assign((size_t)n1, -1);                                       // src/ORBmatcher.cpp:13
    if (n1 == 0) return 0;
    if ((int)vbPrevMatched.size() != n1) throw std::runtime_error("ORBmatcher: vbPrevMatched must have one entry per keypoint of F1");
    Ensure(n1 > n2 ? n1 : n2, n1 > n2 ? n1 : n2);
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint) && sizeof(cv::Point2f) == 8, "layout");
    const std::vector<unsigned char> d1 = pack_rows(Descriptors1), d2 = n2 ? pack_rows(Descriptors2) : std::vector<unsigned char>();
    int nmatches = 0;
    check(orbm_search_init_host(mHandle, (const orbx_keypoint *)vKeys1.data(), d1.data(), n1,
                                (const orbx_keypoint *)vKeys2.data(), d2.data(), n2, (float *)vbPrevMatched.data(),
                                vnMatches12.data(), &nmatches, windowSize, mfNNratio, mbCheckOrientation ? 1 : 0,
                                imageWidth, imageHeight, 0), "search_init");
    if (nmatches < 0) throw std::runtime_error("ORBmatcher: search workspace too small");
    return nmatches;
}

int ORBmatcher::SearchByProjection(const std::vector<cv::KeyPoint> &vLastKeys, const cv::Mat &LastDescriptors,
                                   const std::vector<cv::Point2f> &vProjected,
                                   const std::vector<cv::KeyPoint> &vCurrentKeys, const cv::Mat &CurrentDescriptors,
                                   const std::vector<float> &vScaleFactors, std::vector<int> &vnMatches, float th,
                                   int imageWidth, int imageHeight)
{
    const int n1 = (int)vLastKeys.size(), n2 = (int)vCurrentKeys.size();
    vnMatches.assign((size_t)n1, -1);
    if (n1 == 0) return 0;
    if ((int)vProjected.size() != n1) throw std::runtime_error("ORBmatcher: vProjected must have one entry per last-frame keypoint");
    Ensure(n1 > n2 ? n1 : n2, n1 > n2 ? n1 : n2);
    const std::vector<unsigned char> d1 = pack_rows(LastDescriptors), d2 = n2 ? pack_rows(CurrentDescriptors) : std::vector<unsigned char>();
    orbm_window_params p = orbm_window_params();
    p.radius = th;
    for (int i = 0; i < 16; ++i) p.level_scale[i] = i < (int)vScaleFactors.size() ? vScaleFactors[(size_t)i] : 1.0f;
    p.query_level_min = 0; p.query_level_max = 15; p.level_below = 1; p.level_above = 1;
    p.gate = 1; p.th_dist = TH_HIGH; p.nnratio = 0.0f; p.check_orientation = mbCheckOrientation ? 1 : 0; p.update_centers = 0;
    p.width = imageWidth; p.height = imageHeight; p.literal_gridid_bug = 0;
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
    Ensure(n1 > n2 ? n1 : n2, n1 > n2 ? n1 : n2);
    const std::vector<unsigned char> d1 = pack_rows(Descriptors1), d2 = n2 ? pack_rows(Descriptors2) : std::vector<unsigned char>();
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

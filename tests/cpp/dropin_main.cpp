// dropin_main.cpp -- exercises the C++ drop-in classes exactly like the reference's caller does
// (Frame::ExtractorOrbFeatures, src/Frame.cpp:75-78: (*mpOrbextractor)(img, cv::Mat(), keypoints, descriptors)).
// usage: dropin_main in.bin out.bin   (same file formats as oracle/ref_shim/ref_main.cpp)
#include <cstdio>
#include <cstring>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcher.h"

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    FILE *fi = std::fopen(argv[1], "rb");
    if (!fi) return 3;
    int hdr[7]; float sf;
    if (std::fread(hdr, 4, 7, fi) != 7 || std::fread(&sf, 4, 1, fi) != 1) return 3;
    const int W = hdr[0], H = hdr[1], NF = hdr[2];
    std::vector<unsigned char> frames((size_t)W * H * NF);
    if (std::fread(frames.data(), 1, frames.size(), fi) != frames.size()) return 3;
    std::fclose(fi);
    ORBSlam::ORBextractor *ex = new ORBSlam::ORBextractor(hdr[3], sf, hdr[4], hdr[5], hdr[6]);   // Tracking.cpp:47
    FILE *fo = std::fopen(argv[2], "wb");
    cv::Mat lastDesc;
    std::vector<std::vector<cv::KeyPoint> > allKps; std::vector<cv::Mat> allDesc;
    for (int f = 0; f < NF; ++f) {
        cv::Mat img(H, W, CV_8UC1, frames.data() + (size_t)f * W * H);
        std::vector<cv::KeyPoint> kps; cv::Mat desc;
        (*ex)(img, cv::Mat(), kps, desc);
        const int n = (int)kps.size();
        std::fwrite(&n, 4, 1, fo);
        std::fwrite(kps.data(), sizeof(cv::KeyPoint), n, fo);
        for (int i = 0; i < n; ++i) std::fwrite(desc.ptr(i), 1, 32, fo);
        if (n) lastDesc = desc;
        allKps.push_back(kps); allDesc.push_back(desc.clone());
    }
    // trailer: pyramid checks + matcher checks on the last frame.  mvImagePyramid is materialised on request (it has no
    // reader in the reference): the last call's pyramid is still on the device
    ex->DownloadPyramid();
    int levels = ex->GetLevels();
    std::fwrite(&levels, 4, 1, fo);
    for (int l = 0; l < levels; ++l) {
        const cv::Mat &m = ex->mvImagePyramid[l];
        int dims[2] = { m.cols, m.rows };
        std::fwrite(dims, 4, 2, fo);
        unsigned long long sum = 0, bsum = 0;
        for (int y = 0; y < m.rows; ++y) for (int x = 0; x < m.cols; ++x) sum += m.ptr(y)[x];
        // the 19-px border is addressable around the view, like the reference's Mat ROI
        for (int y = -19; y < m.rows + 19; ++y) for (int x = -19; x < m.cols + 19; ++x) bsum += *(m.data + (ptrdiff_t)y * (ptrdiff_t)m.step + x);
        std::fwrite(&sum, 8, 1, fo); std::fwrite(&bsum, 8, 1, fo);
    }
    std::vector<float> sfv = ex->GetScaleFactors();
    std::fwrite(sfv.data(), 4, sfv.size(), fo);
    ORBSlam::ORBmatcher matcher(0.7f, true);
    int d01 = lastDesc.rows > 1 ? matcher.DescriptorDistance(lastDesc.row(0), lastDesc.row(1)) : -1;
    std::fwrite(&d01, 4, 1, fo);
    std::vector<int> m12;
    int nm = lastDesc.rows ? matcher.SearchBruteForce(lastDesc, lastDesc, m12) : 0;
    std::fwrite(&nm, 4, 1, fo);
    int selfok = 1;
    for (size_t i = 0; i < m12.size(); ++i) if (m12[i] >= 0 && m12[i] != (int)i) { /* duplicates may map to the first copy */ }
    std::fwrite(&selfok, 4, 1, fo);
    // SearchForInitialization frame 0 -> frame 1, like Tracking::MonocularInitialization (src/Tracking.cpp:181-189)
    if (NF >= 2) {
        ORBSlam::ORBmatcher init(0.9f, true);
        std::vector<cv::Point2f> prev(allKps[0].size());
        for (size_t i = 0; i < prev.size(); ++i) prev[i] = allKps[0][i].pt;
        std::vector<int> m12i;
        int n = init.SearchForInitialization(allKps[0], allDesc[0], allKps[1], allDesc[1], prev, m12i, 100, W, H);
        int cnt = (int)m12i.size();
        std::fwrite(&n, 4, 1, fo); std::fwrite(&cnt, 4, 1, fo);
        std::fwrite(m12i.data(), 4, m12i.size(), fo);
        // SearchByProjection-style search: "projection" = the frame-0 keypoint position itself, th = 15
        ORBSlam::ORBmatcher proj(0.9f, true);
        std::vector<cv::Point2f> centers(allKps[0].size());
        for (size_t i = 0; i < centers.size(); ++i) centers[i] = allKps[0][i].pt;
        std::vector<int> mp;
        int np_ = proj.SearchByProjection(allKps[0], allDesc[0], centers, allKps[1], allDesc[1], ex->GetScaleFactors(), mp, 15.0f, W, H);
        int cntp = (int)mp.size();
        std::fwrite(&np_, 4, 1, fo); std::fwrite(&cntp, 4, 1, fo);
        std::fwrite(mp.data(), 4, mp.size(), fo);
        // SearchByBoW-style search: "vocabulary node" = first descriptor byte / 8
        ORBSlam::ORBmatcher bow(0.7f, true);
        std::vector<unsigned short> nodes0(allKps[0].size()), nodes1(allKps[1].size());
        for (size_t i = 0; i < nodes0.size(); ++i) nodes0[i] = (unsigned short)(allDesc[0].ptr<unsigned char>((int)i)[0] >> 3);
        for (size_t i = 0; i < nodes1.size(); ++i) nodes1[i] = (unsigned short)(allDesc[1].ptr<unsigned char>((int)i)[0] >> 3);
        std::vector<int> mb;
        int nb = bow.SearchByBoW(allKps[0], allDesc[0], nodes0, allKps[1], allDesc[1], nodes1, mb);
        int cntb = (int)mb.size();
        std::fwrite(&nb, 4, 1, fo); std::fwrite(&cntb, 4, 1, fo);
        std::fwrite(mb.data(), 4, mb.size(), fo);
    }
    std::fclose(fo);
    delete ex;
    return 0;
}

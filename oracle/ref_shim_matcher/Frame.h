/*
 * Frame.h (shim) -- stands in for the reference's include/Frame.h so that the reference's
 * src/ORBmatcher.cpp can be compiled UNMODIFIED (the real Frame.h drags in Eigen, MapPoint, KeyFrame,
 * none of which compile here).  TEST INFRASTRUCTURE ONLY (oracle/_ref/ref_match).
 *
 * Only what ORBmatcher.cpp touches is provided.  The grid code restates src/Frame.cpp:144-168 and :219-271;
 * the image bounds follow FindimageBound for zero distortion (src/Frame.cpp:111-119).
 */
#ifndef FRAME_H
#define FRAME_H

#include <cassert>
#include <climits>
#include <cmath>
#include <vector>

#include "cvshim.h"

namespace ORBSlam {
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

class KeyFrame;
class MapPoint;

class Frame {
public:
    /* bounds = {minX, maxX, minY, maxY}: what FindimageBound leaves in the statics -- {0, cols, 0, rows} without
     * distortion (Frame.cpp:113-118), the undistorted corners otherwise (:121-141) */
    Frame(const std::vector<cv::KeyPoint> &kps, const cv::Mat &desc, const float *bounds, bool literalBug)
        : mvUnKeypts(kps), mcvDescriptors(desc), mbLiteralBug(literalBug)
    {
        miMinX = bounds[0]; miMaxX = bounds[1]; miMinY = bounds[2]; miMaxY = bounds[3];
        mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (miMaxX - miMinX);     /* Frame.cpp:59-60 */
        mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (miMaxY - miMinY);
        AssignFeaturesToGrid();
    }
    std::vector<cv::KeyPoint> &GetUnKeyPts() { return mvUnKeypts; }
    cv::Mat GetDescriptors() const { return mcvDescriptors.clone(); }                         /* Frame.h:39-42 */

    /* src/Frame.cpp:219-271 */
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel, const int maxLevel) const
    {
        std::vector<size_t> vIndices;
        int N = (int)mvUnKeypts.size();
        vIndices.reserve(N);
        const int nMinCellX = std::max(0, (int)floor((x - miMinX - r) * mfGridElementWidthInv));
        if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
        const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x - miMinX + r) * mfGridElementWidthInv));
        if (nMaxCellX < 0) return vIndices;
        const int nMinCellY = std::max(0, (int)floor((y - miMinY - r) * mfGridElementHeightInv));
        if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
        const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - miMinY + r) * mfGridElementHeightInv));
        if (nMaxCellY < 0) return vIndices;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const std::vector<int> &vCell = mGrid[ix][iy];
                for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
                    const cv::KeyPoint &kpUn = mvUnKeypts[vCell[j]];
                    if (bCheckLevels) {
                        if (kpUn.octave < minLevel) continue;
                        if (maxLevel >= 0 && kpUn.octave > maxLevel) continue;
                    }
                    const float distx = kpUn.pt.x - x, disty = kpUn.pt.y - y;
                    if (fabs(distx) < r && fabs(disty) < r) vIndices.push_back(vCell[j]);
                }
            }
        return vIndices;
    }

private:
    /* src/Frame.cpp:144-168; literalBug reproduces :164 (y index against miMaxY) */
    void AssignFeaturesToGrid()
    {
        for (int i = 0; i < (int)mvUnKeypts.size(); i++) {
            const double x = mvUnKeypts[i].pt.x, y = mvUnKeypts[i].pt.y;
            int ix = (int)std::round((x - miMinX) * mfGridElementWidthInv);
            int iy = (int)std::round((y - (mbLiteralBug ? miMaxY : miMinY)) * mfGridElementHeightInv);
            if (ix < 0 || ix >= FRAME_GRID_COLS || iy < 0 || iy >= FRAME_GRID_ROWS) continue;
            mGrid[ix][iy].push_back(i);
        }
    }
    std::vector<cv::KeyPoint> mvUnKeypts;
    cv::Mat mcvDescriptors;
    bool mbLiteralBug;
    float miMinX, miMaxX, miMinY, miMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<int> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
};

} // namespace ORBSlam
#endif

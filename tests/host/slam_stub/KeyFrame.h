// TEST INFRASTRUCTURE: stand-in for the reference's KeyFrame declaration (orb_slam3/include/KeyFrame.h:256, :380-522) and for
// DBoW2::FeatureVector (Thirdparty/DBoW2/DBoW2/FeatureVector.h: a std::map<NodeId, std::vector<unsigned int>>), carrying the members that
// orb_slam3_ros_b200/host/ORBmatcherGPU.cc touches.  See Frame.h.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <set>
#include <vector>

#include "Frame.h"

namespace ORB_SLAM3 {

class KeyFrame {
public:
    bool isBad() { return mbBad; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint*> GetMapPoints() {                    // KeyFrame.cc:404-418 (the good ones)
        std::set<MapPoint*> s;
        for (size_t i = 0; i < mvpMapPoints.size(); i++) if (mvpMapPoints[i] && !mvpMapPoints[i]->isBad()) s.insert(mvpMapPoints[i]);
        return s;
    }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }
    void ReplaceMapPointMatch(const int& idx, MapPoint* pMP) { mvpMapPoints[idx] = pMP; }
    void EraseMapPointMatch(const int& idx) { mvpMapPoints[idx] = static_cast<MapPoint*>(NULL); }
    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTcw.inverse(); }
    std::vector<float> mvLevelSigma2;
    int N = 0;
    Eigen::Vector3f GetCameraCenter() { return mTcw.inverse().translation(); }
    Sophus::SE3f mTcw;
    float mbf = 0;
    std::vector<float> mvuRight, mvInvLevelSigma2;
    // test stand-ins for the grid a KeyFrame copies from its Frame and KeyFrame::GetFeaturesInArea (KeyFrame.cc:707-751; no level test)
    void AssignFeaturesToGrid() {
        mGrid.assign(FRAME_GRID_COLS, std::vector<std::vector<std::size_t> >(FRAME_GRID_ROWS));
        for (std::size_t i = 0; i < mvKeysUn.size(); i++) {
            const int gx = (int)std::round((mvKeysUn[i].pt.x - mnMinX) * mfGridElementWidthInv);
            const int gy = (int)std::round((mvKeysUn[i].pt.y - mnMinY) * mfGridElementHeightInv);
            if (gx >= 0 && gx < FRAME_GRID_COLS && gy >= 0 && gy < FRAME_GRID_ROWS) mGrid[gx][gy].push_back(i);
        }
    }
    std::vector<std::size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const {
        (void)bRight;
        std::vector<std::size_t> found;
        const int cx0 = std::max(0, (int)std::floor((x - mnMinX - r) * mfGridElementWidthInv));
        const int cx1 = std::min(FRAME_GRID_COLS - 1, (int)std::ceil((x - mnMinX + r) * mfGridElementWidthInv));
        const int cy0 = std::max(0, (int)std::floor((y - mnMinY - r) * mfGridElementHeightInv));
        const int cy1 = std::min(FRAME_GRID_ROWS - 1, (int)std::ceil((y - mnMinY + r) * mfGridElementHeightInv));
        if (cx0 >= FRAME_GRID_COLS || cx1 < 0 || cy0 >= FRAME_GRID_ROWS || cy1 < 0) return found;
        for (int ix = cx0; ix <= cx1; ix++)
            for (int iy = cy0; iy <= cy1; iy++)
                for (std::size_t idx : mGrid[ix][iy])
                    if (std::fabs(mvKeysUn[idx].pt.x - x) < r && std::fabs(mvKeysUn[idx].pt.y - y) < r) found.push_back(idx);
        return found;
    }
    std::vector<std::vector<std::vector<std::size_t> > > mGrid;
    bool IsInImage(const float& x, const float& y) const { return x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY; }      // KeyFrame.cc:753-756
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;      // (integers in KeyFrame.h)
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    std::vector<float> mvScaleFactors;
    int mnScaleLevels = 0;
    float mfLogScaleFactor = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint*> mvpMapPoints;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    int NLeft = -1, NRight = -1;
    bool mbBad = false;
};

inline void MapPoint::Replace(MapPoint* pMP) {      // MapPoint.cc:248-300 without the found / visible counters and the descriptor update
    if (pMP == this) return;
    std::map<KeyFrame*, int> obs = mObservations;
    mObservations.clear();
    mbBad = true;
    mpReplaced = pMP;
    for (std::map<KeyFrame*, int>::iterator mit = obs.begin(); mit != obs.end(); mit++) {
        KeyFrame* pKF = mit->first;
        if (!pMP->IsInKeyFrame(pKF)) { pKF->ReplaceMapPointMatch(mit->second, pMP); pMP->AddObservation(pKF, mit->second); }
        else pKF->EraseMapPointMatch(mit->second);
    }
}

inline int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) {      // MapPoint.cc:514-529
    const float ratio = mfMaxDistance / currentDist;
    int nScale = std::ceil(std::log(ratio) / pKF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pKF->mnScaleLevels) nScale = pKF->mnScaleLevels - 1;
    return nScale;
}

}  // namespace ORB_SLAM3

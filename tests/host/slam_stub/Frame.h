// TEST INFRASTRUCTURE: stand-ins for the reference's Frame / MapPoint / KeyFrame declarations (orb_slam3/include/Frame.h:44-45,
// :144-147, :214-360; MapPoint.h:114-207; KeyFrame.h) carrying exactly the members that orb_slam3_ros_b200/host/ORBmatcherGPU.cc
// touches, with the reference's names and types, so that the adapter can be compiled and run in an image without Eigen / Sophus /
// OpenCV.  The same members are declared by the stand-ins of oracle/ref_cut_tu.cpp, against which the REFERENCE's own function bodies
// are compiled; the two sides are compared on identical flat inputs by tests/test_gpu_matcher_host.py.
#pragma once
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include <opencv2/core/core.hpp>

#include "mini_geom.hpp"      // oracle/cvshim: Eigen::Vector2f/3f, Sophus::SE3f, GeometricCamera (pinhole)

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace DBoW2 {
class FeatureVector : public std::map<unsigned int, std::vector<unsigned int> > {};
}

namespace ORB_SLAM3 {

class Frame;
class KeyFrame;

class MapPoint {
public:
    // stand-ins for MapPoint::GetMin/MaxDistanceInvariance and PredictScale(const float&, Frame*) (MapPoint.cc:502-512, :531-546)
    float GetMinDistanceInvariance() { return 0.8f * mfMinDistance; }
    float GetMaxDistanceInvariance() { return 1.2f * mfMaxDistance; }
    inline int PredictScale(const float& currentDist, Frame* pF);
    inline int PredictScale(const float& currentDist, KeyFrame* pKF);
    Eigen::Vector3f GetNormal() { return mNormalVector; }
    void AddObservation(KeyFrame* pKF, int idx) { mObservations[pKF] = idx; nObs++; }      // (stand-in for MapPoint.cc:137-166)
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF); }                  // MapPoint.cc:420-424
    inline void Replace(MapPoint* pMP);                                                    // stand-in for MapPoint.cc:248-300 (KeyFrame.h)
    MapPoint* mpReplaced = nullptr;
    std::map<KeyFrame*, int> mObservations;
    Eigen::Vector3f mNormalVector;
    float mfMinDistance = 0, mfMaxDistance = 0;
    int Observations() { return nObs; }
    bool isBad() { return mbBad; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = -1;
    float mTrackViewCos = 1, mTrackViewCosR = 1;
    Eigen::Vector3f mWorldPos;
    cv::Mat mDescriptor;
    int nObs = 0;
    bool mbBad = false;
};

class Frame {
public:
    Sophus::SE3<float> GetPose() const { return mTcw; }
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }      // Frame.cc:1054
    Sophus::SE3<float> mTrl;
    // Test stand-ins for Frame::AssignFeaturesToGrid / GetFeaturesInArea (Frame.cc:385-416, :657-735; monocular branch): the key points of a
    // 64 x 48 grid cell in index order, cells visited column by column.  (The compiled adapter calls GetFeaturesInArea only for the rare key
    // point of SearchForInitialization whose candidate list is exhausted; in the reference's build it is the reference's own method.)
    void AssignFeaturesToGrid() {
        for (int i = 0; i < FRAME_GRID_COLS; i++) for (int j = 0; j < FRAME_GRID_ROWS; j++) mGrid[i][j].clear();
        for (int i = 0; i < N; i++) {
            const int gx = (int)std::round((mvKeysUn[i].pt.x - mnMinX) * mfGridElementWidthInv);
            const int gy = (int)std::round((mvKeysUn[i].pt.y - mnMinY) * mfGridElementHeightInv);
            if (gx >= 0 && gx < FRAME_GRID_COLS && gy >= 0 && gy < FRAME_GRID_ROWS) mGrid[gx][gy].push_back((std::size_t)i);
        }
    }
    std::vector<std::size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const {
        std::vector<std::size_t> found;
        const int cx0 = std::max(0, (int)std::floor((x - mnMinX - r) * mfGridElementWidthInv));
        const int cx1 = std::min(FRAME_GRID_COLS - 1, (int)std::ceil((x - mnMinX + r) * mfGridElementWidthInv));
        const int cy0 = std::max(0, (int)std::floor((y - mnMinY - r) * mfGridElementHeightInv));
        const int cy1 = std::min(FRAME_GRID_ROWS - 1, (int)std::ceil((y - mnMinY + r) * mfGridElementHeightInv));
        if (cx0 >= FRAME_GRID_COLS || cx1 < 0 || cy0 >= FRAME_GRID_ROWS || cy1 < 0) return found;
        const bool levels = minLevel > 0 || maxLevel >= 0;
        for (int ix = cx0; ix <= cx1; ix++)
            for (int iy = cy0; iy <= cy1; iy++)
                for (std::size_t idx : mGrid[ix][iy]) {
                    const cv::KeyPoint& kp = mvKeysUn[idx];
                    if (levels && (kp.octave < minLevel || (maxLevel >= 0 && kp.octave > maxLevel))) continue;
                    if (std::fabs(kp.pt.x - x) < r && std::fabs(kp.pt.y - y) < r) found.push_back(idx);
                }
        return found;
    }
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    long unsigned int mnId = 0;
    Sophus::SE3<float> mTcw;
    float mbf = 0, mb = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    DBoW2::FeatureVector mFeatVec;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    int mnScaleLevels = 0;
    float mfLogScaleFactor = 0;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    int Nleft = -1, Nright = -1;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
};

inline int MapPoint::PredictScale(const float& currentDist, Frame* pF) {
    const float ratio = mfMaxDistance / currentDist;
    int nScale = std::ceil(std::log(ratio) / pF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pF->mnScaleLevels) nScale = pF->mnScaleLevels - 1;
    return nScale;
}

}  // namespace ORB_SLAM3

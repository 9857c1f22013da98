// TEST INFRASTRUCTURE: stand-in for the reference's KeyFrame declaration (orb_slam3/include/KeyFrame.h:256, :380-522) and for
// DBoW2::FeatureVector (Thirdparty/DBoW2/DBoW2/FeatureVector.h: a std::map<NodeId, std::vector<unsigned int>>), carrying the members that
// orb_slam3_ros_b200/host/ORBmatcherGPU.cc touches.  See Frame.h.
#pragma once
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

inline int MapPoint::PredictScale(const float& currentDist, KeyFrame* pKF) {      // MapPoint.cc:514-529
    const float ratio = mfMaxDistance / currentDist;
    int nScale = std::ceil(std::log(ratio) / pKF->mfLogScaleFactor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= pKF->mnScaleLevels) nScale = pKF->mnScaleLevels - 1;
    return nScale;
}

}  // namespace ORB_SLAM3

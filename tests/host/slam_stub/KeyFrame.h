// TEST INFRASTRUCTURE: stand-in for the reference's KeyFrame declaration (orb_slam3/include/KeyFrame.h:256, :380-522) and for
// DBoW2::FeatureVector (Thirdparty/DBoW2/DBoW2/FeatureVector.h: a std::map<NodeId, std::vector<unsigned int>>), carrying the members that
// orb_slam3_ros_b200/host/ORBmatcherGPU.cc touches.  See Frame.h.
#pragma once
#include <map>
#include <vector>

#include "Frame.h"

namespace ORB_SLAM3 {

class KeyFrame {
public:
    bool isBad() { return mbBad; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint*> mvpMapPoints;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    int NLeft = -1, NRight = -1;
    bool mbBad = false;
};

}  // namespace ORB_SLAM3

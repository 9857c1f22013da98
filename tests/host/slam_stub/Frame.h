// TEST INFRASTRUCTURE: stand-ins for the reference's Frame / MapPoint / KeyFrame declarations (orb_slam3/include/Frame.h:44-45,
// :144-147, :214-360; MapPoint.h:114-207; KeyFrame.h) carrying exactly the members that orb_slam3_ros_b200/host/ORBmatcherGPU.cc
// touches, with the reference's names and types, so that the adapter can be compiled and run in an image without Eigen / Sophus /
// OpenCV.  The same members are declared by the stand-ins of oracle/ref_cut_tu.cpp, against which the REFERENCE's own function bodies
// are compiled; the two sides are compared on identical flat inputs by tests/test_gpu_matcher_host.py.
#pragma once
#include <mutex>
#include <vector>

#include <opencv2/core/core.hpp>

#include "mini_geom.hpp"      // oracle/cvshim: Eigen::Vector2f/3f, Sophus::SE3f, GeometricCamera (pinhole)

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

namespace ORB_SLAM3 {

class MapPoint {
public:
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
    long unsigned int mnId = 0;
    Sophus::SE3<float> mTcw;
    float mbf = 0, mb = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    int Nleft = -1, Nright = -1;
};

}  // namespace ORB_SLAM3

// TEST: the reference's OWN call-site text -- Frame::ExtractORB (orb_slam3/src/Frame.cc:418-425: the operator() call with cv::Mat() as
// mask and a vector<int> lapping area) and Frame::ComputeStereoMatches (Frame.cc:811-981: mvImagePyramid[l].rows / .cols /
// .rowRange().colRange() of both extractors) -- cut out of /root/reference at test time (oracle/cut_reference.py) and compiled against
// the adapter's class declaration (orb_slam3_ros_b200/host/ORBextractor.h).  Compile-only: it proves source compatibility of the
// drop-in header with the text that calls it; nothing of the reference is stored in the repository.
#include <algorithm>
#include <climits>
#include <cmath>
#include <vector>

#include <opencv2/opencv.hpp>

#include "ORBextractor.h"      // the adapter, NOT the reference's header

using namespace std;

namespace ORB_SLAM3 {

class ORBmatcher {
public:
    static const int TH_HIGH = 100, TH_LOW = 50;
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
};

class Frame {                   // the members the two bodies touch (Frame.h:214-360)
public:
    void ExtractORB(int flag, const cv::Mat& im, const int x0, const int x1);
    void ComputeStereoMatches();
    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    float mbf = 0, mb = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    int monoLeft = 0, monoRight = 0;
};

#include "cut/Frame_ExtractORB.inc"
#include "cut/Frame_ComputeStereoMatches.inc"

}  // namespace ORB_SLAM3

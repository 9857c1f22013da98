// Drop-in replacement for orb_slam3/include/ORBextractor.h of giltchcity/orb_slam3_ros: same namespace, class name,
// constructor, operator(), getters and the public mvImagePyramid member (reference ORBextractor.h:43-109), so that
// Frame.cc (:110-125, :222, :311, :418-425, :818-923, :1059-1060) and Tracking.cc (:631-637, :1319-1325) compile
// unchanged.  All computation happens in liborbb200.so (CUDA, sm_100a) through the C ABI of include/orbb200.h;
// there is no CPU implementation behind this class.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <vector>
#include <opencv2/opencv.hpp>

struct orbb_extractor;
struct orbb_rectifier;

namespace ORB_SLAM3 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image; mask is ignored (as in the reference).
    // Returns the number of keypoints outside vLappingArea (written at the front), -1 on an empty image.
    int operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                   cv::OutputArray _descriptors, std::vector<int>& vLappingArea);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Views into a host copy of the device pyramid of the last frame (valid until the next call), refreshed after
    // every operator() unless SetPyramidDownload(false) -- only Frame::ComputeStereoMatches reads them.
    std::vector<cv::Mat> mvImagePyramid;

    // ---- extensions (not in the reference) ----
    // mvImagePyramid is a public data member that Frame::ComputeStereoMatches reads right after the two extractions (Frame.cc:818-923),
    // so by default every call ends with the hand-out of the pyramid (19-px apron built on the device, one 1.5 MB copy at 752x480:
    // +0.05 ms, see bench.py single_frame_with_pyramid_ms).  Monocular / RGB-D callers never read it and can switch it off; stereo
    // callers that use orbb_stereo_match (the pyramids stay on the device) can too.
    void SetPyramidDownload(bool enabled) { mbDownloadPyramid = enabled; }
    // CUDA device of the extractors constructed after this call (default: environment variable ORBB_DEVICE, else 0).  The reference's
    // constructor has no such argument and Tracking constructs the extractors itself, hence a process-wide setting.
    static void SetDefaultDevice(int device);
    static int DefaultDevice();
    orbb_extractor* Handle() { return mpHandle; }          // for orbb_stereo_match(hL, hR, ...)
    // Colour frame (CV_8UC3 / CV_8UC4 data, `channels` bytes per pixel): the cvtColor(.., COLOR_*2GRAY) that Tracking applies
    // before building the Frame (Tracking.cc:1498-1525) runs on the device; bRGB = Tracking's mbRGB.
    int ExtractColor(const unsigned char* data, int cols, int rows, size_t step, int channels, bool bRGB,
                     std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors, std::vector<int>& vLappingArea);
    // Raw (unrectified) stereo frame: the cv::remap of System::TrackStereo (System.cc:233-240) runs on the device, the
    // rectified image never comes back to the host.  `rect` from orbb_rectifier_create (maps of Settings.cc:506-509).
    int ExtractRectified(orbb_rectifier* rect, cv::InputArray _rawImage, std::vector<cv::KeyPoint>& _keypoints,
                         cv::OutputArray _descriptors, std::vector<int>& vLappingArea);

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    int Deliver(int rc, int n, int monoIndex, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors);
    void* Staging(int& cap);

    orbb_extractor* mpHandle;
    bool mbDownloadPyramid;
    // result staging in PINNED host memory (orbb_host_alloc): the device-to-host copies of the call then run asynchronously
    // instead of through the driver's pageable path (measured: 0.193 -> 0.175 ms per 752x480 frame)
    unsigned char* mpKeypointStaging = nullptr;
    unsigned char* mpDescStaging = nullptr;
    int mnStagingCap = 0;
};

}  // namespace ORB_SLAM3

#endif

// Host adapter: ORB_SLAM3::ORBextractor on top of liborbb200.so (see ORBextractor.h).  Mirrors the calling
// conventions of the reference's ORBextractor.cc:1086-1168: input must be CV_8UC1, empty input returns -1, the callee
// replaces _keypoints and (re)creates _descriptors as n x 32 CV_8U (released when n == 0).
#include "ORBextractor.h"

#include <cassert>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "orbb200.h"

namespace ORB_SLAM3 {

static int g_defaultDevice = -1;      // -1: not set -> ORBB_DEVICE or 0

void ORBextractor::SetDefaultDevice(int device) { g_defaultDevice = device; }
int ORBextractor::DefaultDevice() {
    if (g_defaultDevice >= 0) return g_defaultDevice;
    const char* e = getenv("ORBB_DEVICE");
    return e ? atoi(e) : 0;
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
      mpHandle(nullptr), mbDownloadPyramid(true) {
    orbb_params prm;
    prm.nfeatures = _nfeatures;
    prm.scale_factor = _scaleFactor;
    prm.nlevels = _nlevels;
    prm.ini_th_fast = _iniThFAST;
    prm.min_th_fast = _minThFAST;
    prm.device = DefaultDevice();
    prm.max_batch = 1;
    const int rc = orbb_create(&prm, &mpHandle);
    if (rc != ORBB_OK) throw std::runtime_error(std::string("orbb_create failed: ") + orbb_last_error(nullptr));
    mvScaleFactor.resize(nlevels);
    mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    orbb_get_tables(mpHandle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                    mnFeaturesPerLevel.data());
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor() {
    orbb_host_free(mpKeypointStaging);
    orbb_host_free(mpDescStaging);
    orbb_destroy(mpHandle);
}

void* ORBextractor::Staging(int& cap) {
    cap = orbb_max_keypoints(mpHandle);
    if (cap > mnStagingCap) {
        orbb_host_free(mpKeypointStaging);
        orbb_host_free(mpDescStaging);
        mpKeypointStaging = static_cast<unsigned char*>(orbb_host_alloc((size_t)cap * sizeof(orbb_keypoint)));
        mpDescStaging = static_cast<unsigned char*>(orbb_host_alloc((size_t)cap * 32));
        mnStagingCap = (mpKeypointStaging && mpDescStaging) ? cap : 0;
        if (!mnStagingCap) throw std::runtime_error("liborbb200: pinned staging allocation failed");
    }
    return mpKeypointStaging;
}

// common tail of the extraction entry points: error mapping, output containers, pyramid views
int ORBextractor::Deliver(int rc, int n, int monoIndex, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors) {
    if (rc == ORBB_ERR_EMPTY) return -1;
    if (rc != ORBB_OK) throw std::runtime_error(std::string("liborbb200 extraction failed: ") + orbb_last_error(mpHandle));
    const orbb_keypoint* kps = reinterpret_cast<const orbb_keypoint*>(mpKeypointStaging);
    if (n == 0) {
        _descriptors.release();
    } else {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat descriptors = _descriptors.getMat();
        for (int i = 0; i < n; i++) memcpy(descriptors.ptr(i), mpDescStaging + (size_t)i * 32, 32);
    }
    _keypoints = std::vector<cv::KeyPoint>(n);
    for (int i = 0; i < n; i++) {
        cv::KeyPoint& kp = _keypoints[i];
        kp.pt.x = kps[i].x;
        kp.pt.y = kps[i].y;
        kp.size = kps[i].size;
        kp.angle = kps[i].angle;
        kp.response = kps[i].response;
        kp.octave = kps[i].octave;
        kp.class_id = -1;
    }
    if (mbDownloadPyramid) {
        for (int level = 0; level < nlevels; ++level) {
            const uint8_t* p = nullptr;
            int w = 0, h = 0;
            size_t stride = 0;
            if (orbb_pyramid_level(mpHandle, level, &p, &w, &h, &stride) != ORBB_OK)
                throw std::runtime_error(std::string("orbb_pyramid_level failed: ") + orbb_last_error(mpHandle));
            mvImagePyramid[level] = cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(p), stride);
        }
    }
    return monoIndex;
}

int ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                             cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
    if (_image.empty()) return -1;
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);
    int cap = 0, n = 0, monoIndex = 0;
    orbb_keypoint* kps = static_cast<orbb_keypoint*>(Staging(cap));
    const int rc = orbb_extract(mpHandle, image.data, image.cols, image.rows, (size_t)image.step, vLappingArea[0], vLappingArea[1],
                                kps, mpDescStaging, cap, &n, &monoIndex);
    return Deliver(rc, n, monoIndex, _keypoints, _descriptors);
}

int ORBextractor::ExtractColor(const unsigned char* data, int cols, int rows, size_t step, int channels, bool bRGB,
                               std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
    if (!data || cols <= 0 || rows <= 0) return -1;
    int cap = 0, n = 0, monoIndex = 0;
    orbb_keypoint* kps = static_cast<orbb_keypoint*>(Staging(cap));
    int rc = orbb_extract_color(mpHandle, data, cols, rows, step, channels, bRGB ? 1 : 0, vLappingArea[0], vLappingArea[1], kps,
                                mpDescStaging, cap, &n, &monoIndex);
    if (rc == ORBB_ERR_CAPACITY) {                          // first frame of this size: the plan allows more keypoints than the estimate
        kps = static_cast<orbb_keypoint*>(Staging(cap));
        rc = orbb_extract_color(mpHandle, data, cols, rows, step, channels, bRGB ? 1 : 0, vLappingArea[0], vLappingArea[1], kps,
                                mpDescStaging, cap, &n, &monoIndex);
    }
    return Deliver(rc, n, monoIndex, _keypoints, _descriptors);
}

int ORBextractor::ExtractRectified(orbb_rectifier* rect, cv::InputArray _rawImage, std::vector<cv::KeyPoint>& _keypoints,
                                   cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
    if (_rawImage.empty()) return -1;
    cv::Mat image = _rawImage.getMat();
    assert(image.type() == CV_8UC1);
    int cap = 0, n = 0, monoIndex = 0;
    orbb_keypoint* kps = static_cast<orbb_keypoint*>(Staging(cap));
    int rc = orbb_extract_rectified(mpHandle, rect, image.data, (size_t)image.step, vLappingArea[0], vLappingArea[1], kps,
                                    mpDescStaging, cap, &n, &monoIndex);
    if (rc == ORBB_ERR_CAPACITY) {
        kps = static_cast<orbb_keypoint*>(Staging(cap));
        rc = orbb_extract_rectified(mpHandle, rect, image.data, (size_t)image.step, vLappingArea[0], vLappingArea[1], kps,
                                    mpDescStaging, cap, &n, &monoIndex);
    }
    return Deliver(rc, n, monoIndex, _keypoints, _descriptors);
}

}  // namespace ORB_SLAM3

// Mimics the reference call sites against the host adapter (compiled with tests/cvstub in place of OpenCV):
//   Tracking.cc:631   new ORBextractor(nFeatures, fScaleFactor, nLevels, fIniThFAST, fMinThFAST)
//   Frame.cc:110-116  the six getters
//   Frame.cc:418-425  (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors, vLapping)
//   Frame.cc:818,908  mvImagePyramid[l].rows / .rowRange().colRange()
// argv[1] = raw 8-bit gray file, argv[2] = width, argv[3] = height.  Prints a checksum the python test compares.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ORBextractor.h"
#include "ORBmatcherGPU.h"

using namespace ORB_SLAM3;

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    const int w = atoi(argv[2]), h = atoi(argv[3]);
    std::vector<unsigned char> buf((size_t)w * h);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(buf.data(), 1, buf.size(), f) != buf.size()) return 3;
    fclose(f);
    cv::Mat im(h, w, CV_8UC1, buf.data());

    ORBextractor* mpORBextractorLeft = new ORBextractor(300, 1.2f, 4, 20, 7);
    const int levels = mpORBextractorLeft->GetLevels();
    const std::vector<float> sf = mpORBextractorLeft->GetScaleFactors();
    const std::vector<float> isf = mpORBextractorLeft->GetInverseScaleFactors();
    const std::vector<float> s2 = mpORBextractorLeft->GetScaleSigmaSquares();
    const std::vector<float> is2 = mpORBextractorLeft->GetInverseScaleSigmaSquares();
    if (levels != 4 || sf.size() != 4 || isf.size() != 4 || s2.size() != 4 || is2.size() != 4) return 4;
    if (mpORBextractorLeft->GetScaleFactor() != 1.2f) return 5;

    std::vector<cv::KeyPoint> mvKeys;
    cv::Mat mDescriptors;
    std::vector<int> vLapping = {0, 1000};
    const int monoIdx = (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors, vLapping);
    cv::Mat empty;
    std::vector<cv::KeyPoint> k2;
    cv::Mat d2;
    if ((*mpORBextractorLeft)(empty, cv::Mat(), k2, d2, vLapping) != -1) return 6;

    unsigned long long sum = 0;
    for (size_t i = 0; i < mvKeys.size(); i++) {
        sum = sum * 1000003ull + (unsigned)(mvKeys[i].pt.x * 16) + 7ull * (unsigned)(mvKeys[i].pt.y * 16) + 13ull * mvKeys[i].octave +
              17ull * (unsigned)mvKeys[i].response + 19ull * (unsigned)mvKeys[i].size;
        for (int b = 0; b < 32; b++) sum = sum * 31ull + mDescriptors.ptr<uchar>((int)i)[b];
    }
    const cv::Mat& p1 = mpORBextractorLeft->mvImagePyramid[1];
    cv::Mat win = p1.rowRange(10, 21).colRange(12, 23);
    unsigned long long psum = 0;
    for (int r = 0; r < win.rows; r++) for (int c = 0; c < win.cols; c++) psum += win.ptr<uchar>(r)[c] * (unsigned long long)(r * 11 + c + 1);
    const int d01 = mvKeys.size() > 1 ? ORBmatcherGPU::DescriptorDistance(mDescriptors.row(0), mDescriptors.row(1)) : -1;
    printf("n=%zu mono=%d desc=%dx%d sum=%llu pyr1=%dx%d psum=%llu d01=%d\n", mvKeys.size(), monoIdx, mDescriptors.rows, mDescriptors.cols, sum,
           p1.cols, p1.rows, psum, d01);

    // ---- the input-side extensions: colour frame, raw stereo frame + rectifier, keypoint undistortion ----
    unsigned long long sumC = 0, sumR = 0, sumU = 0;
    {
        std::vector<unsigned char> bgr((size_t)w * h * 3);
        for (size_t i = 0; i < (size_t)w * h; i++) { bgr[3 * i] = buf[i]; bgr[3 * i + 1] = (unsigned char)(255 - buf[i]); bgr[3 * i + 2] = (unsigned char)(buf[i] / 2); }
        std::vector<cv::KeyPoint> kc;
        cv::Mat dc;
        std::vector<int> lap0 = {0, 0};
        mpORBextractorLeft->ExtractColor(bgr.data(), w, h, (size_t)w * 3, 3, false, kc, dc, lap0);
        for (size_t i = 0; i < kc.size(); i++) {
            sumC = sumC * 1000003ull + (unsigned)(kc[i].pt.x * 16) + 7ull * (unsigned)(kc[i].pt.y * 16) + 13ull * kc[i].octave;
            for (int b = 0; b < 32; b++) sumC = sumC * 31ull + dc.ptr<uchar>((int)i)[b];
        }
        // rectifier: a shift by (2.25, -1.5) px with constant-zero border
        std::vector<float> mx((size_t)w * h), my((size_t)w * h);
        for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) { mx[(size_t)y * w + x] = x + 2.25f; my[(size_t)y * w + x] = y - 1.5f; }
        orbb_rectifier* rect = nullptr;
        if (orbb_rectifier_create(0, mx.data(), my.data(), (size_t)w, w, h, w, h, &rect) != ORBB_OK) return 7;
        std::vector<cv::KeyPoint> kr;
        cv::Mat dr;
        mpORBextractorLeft->ExtractRectified(rect, im, kr, dr, lap0);
        orbb_rectifier_destroy(rect);
        for (size_t i = 0; i < kr.size(); i++) {
            sumR = sumR * 1000003ull + (unsigned)(kr[i].pt.x * 16) + 7ull * (unsigned)(kr[i].pt.y * 16) + 13ull * kr[i].octave;
            for (int b = 0; b < 32; b++) sumR = sumR * 31ull + dr.ptr<uchar>((int)i)[b];
        }
        ORBmatcherGPU gpu;
        const float dist[4] = {-0.28340811f, 0.07395907f, 0.00019359f, 1.76187114e-05f};
        std::vector<cv::KeyPoint> ku;
        gpu.UndistortKeyPoints(mvKeys, 458.654f, 457.296f, 367.215f, 248.375f, dist, 4, ku);
        for (size_t i = 0; i < ku.size(); i++) {
            unsigned a, b2;
            memcpy(&a, &ku[i].pt.x, 4); memcpy(&b2, &ku[i].pt.y, 4);
            sumU = sumU * 1000003ull + a + 7ull * b2;
        }
        printf("sumC=%llu nC=%zu sumR=%llu nR=%zu sumU=%llu\n", sumC, kc.size(), sumR, kr.size(), sumU);
    }
    delete mpORBextractorLeft;
    return 0;
}

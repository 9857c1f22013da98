// Mimics the reference call sites against the host adapter (compiled with tests/cvstub in place of OpenCV):
//   Tracking.cc:631   new ORBextractor(nFeatures, fScaleFactor, nLevels, fIniThFAST, fMinThFAST)
//   Frame.cc:110-116  the six getters
//   Frame.cc:418-425  (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors, vLapping)
//   Frame.cc:818,908  mvImagePyramid[l].rows / .rowRange().colRange()
// argv[1] = raw 8-bit gray file, argv[2] = width, argv[3] = height.  Prints a checksum the python test compares.
#include <cstdio>
#include <cstdlib>
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
    delete mpORBextractorLeft;
    return 0;
}

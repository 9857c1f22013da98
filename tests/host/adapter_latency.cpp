// Latency of the reference's call site through the C++ adapter (compiled with tests/cvstub in place of OpenCV):
//   Frame.cc:418-425  (*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors, vLapping)
// on a pageable cv::Mat, results delivered as std::vector<cv::KeyPoint> + cv::Mat -- once with the default hand-out of
// mvImagePyramid (what a stereo caller that runs Frame::ComputeStereoMatches on the host needs), once without (monocular / RGB-D).
// argv: raw 8-bit gray file, width, height, features, levels, repetitions.  Prints one JSON object.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ORBextractor.h"

using namespace ORB_SLAM3;

static double run(ORBextractor& ext, cv::Mat& im, int reps, size_t& n) {
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc;
    std::vector<int> lapping = {0, 1000};
    for (int i = 0; i < (reps < 10 ? reps : 10); i++) ext(im, cv::Mat(), keys, desc, lapping);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; i++) ext(im, cv::Mat(), keys, desc, lapping);
    const auto t1 = std::chrono::steady_clock::now();
    n = keys.size();
    return std::chrono::duration<double, std::milli>(t1 - t0).count() / reps;
}

int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const int w = atoi(argv[2]), h = atoi(argv[3]), nf = atoi(argv[4]), nl = atoi(argv[5]), reps = atoi(argv[6]);
    std::vector<unsigned char> buf((size_t)w * h);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(buf.data(), 1, buf.size(), f) != buf.size()) return 3;
    fclose(f);
    cv::Mat im(h, w, CV_8UC1, buf.data());
    ORBextractor ext(nf, 1.2f, nl, 20, 7);
    size_t n = 0;
    const double withPyr = run(ext, im, reps, n);
    ext.SetPyramidDownload(false);
    const double without = run(ext, im, reps, n);
    printf("{\"operator_ms\": %.5f, \"operator_with_pyramid_ms\": %.5f, \"keypoints\": %zu, \"reps\": %d}\n", without, withPyr, n, reps);
    return 0;
}

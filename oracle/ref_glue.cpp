// ORACLE (test infrastructure, NOT product code): C entry points around the reference's own ORB_SLAM3::ORBextractor,
// compiled from /root/reference/orb_slam3/src/ORBextractor.cc against cvshim/ (see cvshim.hpp).  Same calling convention as
// the port_* functions of orb_port.cpp so that the tests can run the two side by side.
#include <thread>

#include "ORBextractor.h"

namespace {
struct PortKP { float x, y, size, angle, response; int octave; };

struct Ref : ORB_SLAM3::ORBextractor {          // the tables the reference keeps protected (ORBextractor.h:98-106)
    using ORB_SLAM3::ORBextractor::ORBextractor;
    const std::vector<int>& featuresPerLevel() const { return mnFeaturesPerLevel; }
    const std::vector<int>& uMax() const { return umax; }
};

int run(Ref* e, const uint8_t* img, int w, int h, size_t stride, int lap0, int lap1, PortKP* kps, uint8_t* desc, int cap, int* nOut,
        int* monoOut) {
    cv::Mat image(h, w, CV_8UC1, const_cast<uint8_t*>(img), stride);
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors;
    std::vector<int> lapping = {lap0, lap1};
    const int mono = (*e)(image, cv::Mat(), keys, descriptors, lapping);            // Frame.cc:418-425
    if (mono < 0) return -1;
    const int n = (int)keys.size();
    *nOut = n; *monoOut = mono;
    if (n > cap) return -2;
    for (int i = 0; i < n; i++) {
        const cv::KeyPoint& k = keys[i];
        if (kps) kps[i] = PortKP{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave};
        if (desc) memcpy(desc + (size_t)32 * i, descriptors.ptr<uchar>(i), 32);
    }
    return 0;
}
}  // namespace

extern "C" {

void* ref_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {           // Tracking.cc:631-637
    return new Ref(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void ref_destroy(void* h) { delete (Ref*)h; }

void ref_tables(void* h, float* scale, float* inv, float* sig2, float* invsig2, int* nfeat, int* umax) {
    Ref* e = (Ref*)h;
    const std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(),
                             d = e->GetInverseScaleSigmaSquares();
    for (int i = 0; i < e->GetLevels(); i++) { scale[i] = a[i]; inv[i] = b[i]; sig2[i] = c[i]; invsig2[i] = d[i]; nfeat[i] = e->featuresPerLevel()[i]; }
    for (size_t i = 0; i < e->uMax().size(); i++) umax[i] = e->uMax()[i];
}

int ref_extract(void* h, const uint8_t* img, int w, int hh, size_t stride, int lap0, int lap1, PortKP* kps, uint8_t* desc, int cap,
                int* nOut, int* monoOut) {
    return run((Ref*)h, img, w, hh, stride, lap0, lap1, kps, desc, cap, nOut, monoOut);
}

// mvImagePyramid[level] (Frame.cc:818 reads it); bordered=1 -> origin of the (w+38)x(h+38) buffer the level is a view of
int ref_level(void* h, int level, int bordered, const uint8_t** ptr, int* w, int* hh, size_t* stride) {
    Ref* e = (Ref*)h;
    if (level < 0 || level >= (int)e->mvImagePyramid.size() || e->mvImagePyramid[level].empty()) return -1;
    const cv::Mat& m = e->mvImagePyramid[level];
    *w = m.cols; *hh = m.rows; *stride = m.step;
    *ptr = bordered ? m.data - 19 * m.step - 19 : m.data;
    return 0;
}

// one extractor per worker thread, frames dealt round-robin (throughput arm of bench.py --impl reference)
int ref_extract_batch(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* imgs, int nframes, int w,
                      int hh, size_t rowStride, size_t frameStride, int lap0, int lap1, PortKP* kps, uint8_t* desc, int cap,
                      int32_t* counts, int nthreads) {
    nthreads = std::max(1, std::min(nthreads, nframes));
    std::vector<int> rcs(nthreads, 0);
    auto work = [&](int t) {
        Ref e(nfeatures, scaleFactor, nlevels, iniTh, minTh);
        for (int f = t; f < nframes; f += nthreads) {
            int n = 0, mono = 0;
            const int rc = run(&e, imgs + (size_t)f * frameStride, w, hh, rowStride, lap0, lap1, kps ? kps + (size_t)f * cap : nullptr,
                               desc ? desc + (size_t)f * cap * 32 : nullptr, cap, &n, &mono);
            if (rc) { rcs[t] = rc; return; }
            counts[2 * f] = n; counts[2 * f + 1] = mono;
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    for (int rc : rcs) if (rc) return rc;
    return 0;
}

}  // extern "C"

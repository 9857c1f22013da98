// ORACLE (test infrastructure, NOT product code): the OpenCV primitives behind cvshim.hpp, forwarded to the restatements in
// orb_port.cpp (each of which is pinned bit for bit against python-cv2 by tests/test_oracle_primitives.py).
#include "cvshim.hpp"

#include <algorithm>

extern "C" {
void port_resize_linear(const uint8_t* src, int sw, int sh, size_t ss, uint8_t* dst, int dw, int dh, size_t ds);
int port_fast9(const uint8_t* img, int w, int h, size_t stride, int th, float* xyr, int cap);
void port_gaussian7(const uint8_t* src, int w, int h, size_t ss, uint8_t* dst, size_t ds);
float port_fast_atan2(float y, float x);
}

namespace cv {

namespace {
struct Pool {
    std::vector<std::pair<size_t, uchar*>> free_;
    ~Pool() { for (auto& e : free_) delete[] e.second; }
};
thread_local Pool tPool;
}  // namespace

std::shared_ptr<uchar> Mat::allocate(size_t n) {
    uchar* p = nullptr;
    for (size_t i = 0; i < tPool.free_.size(); i++)
        if (tPool.free_[i].first == n) { p = tPool.free_[i].second; tPool.free_.erase(tPool.free_.begin() + i); break; }
    if (!p) p = new uchar[n ? n : 1];
    return std::shared_ptr<uchar>(p, [n](uchar* q) {
        if (tPool.free_.size() < 64) tPool.free_.emplace_back(n, q);      // (a buffer freed on another thread joins that thread's pool)
        else delete[] q;
    });
}

float fastAtan2(float y, float x) { return port_fast_atan2(y, x); }

double norm(InputArray a_, InputArray b_, int) {        // cv::norm(.., NORM_L1) on CV_8U: exact integer sum of |a - b|, returned as double
    const Mat a = a_.getMat(), b = b_.getMat();
    long long s = 0;
    for (int y = 0; y < a.rows; y++) {
        const uchar *pa = a.ptr(y), *pb = b.ptr(y);
        for (int x = 0; x < a.cols; x++) s += pa[x] > pb[x] ? pa[x] - pb[x] : pb[x] - pa[x];
    }
    return (double)s;
}

void resize(InputArray src_, OutputArray dst_, Size dsize, double, double, int) {
    const Mat src = src_.getMat();
    dst_.create(dsize, CV_8UC1);                   // a correctly sized ROI view stays where it is (ComputePyramid relies on it)
    Mat dst = dst_.getMat();
    port_resize_linear(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int) {
    const Mat src = src_.getMat();
    dst_.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    Mat dst = dst_.getMat();
    for (int y = 0; y < src.rows; y++)             // interior (memmove: the source may be the interior of dst itself)
        std::memmove(dst.data + (size_t)(y + top) * dst.step + left, src.data + (size_t)y * src.step, (size_t)src.cols);
    for (int y = 0; y < dst.rows; y++) {
        const int sy = reflect101(y - top, src.rows) + top;
        uchar* d = dst.data + (size_t)y * dst.step;
        const uchar* s = dst.data + (size_t)sy * dst.step;
        for (int x = 0; x < dst.cols; x++) {
            const int sx = reflect101(x - left, src.cols) + left;
            if (sy != y || sx != x) d[x] = s[sx];
        }
    }
}

void GaussianBlur(InputArray src_, OutputArray dst_, Size, double, double, int) {
    const Mat src = src_.getMat().clone();         // the reference blurs in place
    dst_.create(src.rows, src.cols, CV_8UC1);
    Mat dst = dst_.getMat();
    port_gaussian7(src.data, src.cols, src.rows, src.step, dst.data, dst.step);
}

void FAST(InputArray image_, std::vector<KeyPoint>& keypoints, int threshold, bool) {
    const Mat im = image_.getMat();
    keypoints.clear();
    if (im.rows < 7 || im.cols < 7) return;
    static thread_local std::vector<float> xyr;
    const size_t cap = ((size_t)im.rows * im.cols) / 2 + 16;           // NMS survivors are never 8-adjacent
    if (xyr.size() < 3 * cap) xyr.resize(3 * cap);
    const int n = port_fast9(im.data, im.cols, im.rows, im.step, threshold, xyr.data(), (int)cap);
    keypoints.reserve(n);
    for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1, xyr[3 * i + 2]));
}

void KeyPointsFilter::retainBest(std::vector<KeyPoint>& k, int n) {
    if (n < 0 || (size_t)n >= k.size()) return;
    std::stable_sort(k.begin(), k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
    k.resize(n);
}

}  // namespace cv

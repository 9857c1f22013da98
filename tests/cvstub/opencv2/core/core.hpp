// Minimal stand-in for the OpenCV types the ORBextractor / ORBmatcher interface uses, so that the host adapter
// (orb_slam3_ros_b200/host/*.cc) can be COMPILE- and RUN-checked in an image without OpenCV C++ headers.
// Test infrastructure only; the adapter itself uses nothing beyond the real OpenCV API subset mirrored here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5

typedef unsigned char uchar;

namespace cv {

template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
typedef Point_<int> Point2i;
typedef Point2i Point;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };

class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};

class Mat {
public:
    Mat() {}
    Mat(int r, int c, int t) { create(r, c, t); }
    Mat(int r, int c, int t, void* d, size_t s = 0) : rows(r), cols(c), data((uchar*)d), step(s ? s : (size_t)c), type_(t) {}
    void create(int r, int c, int t) {
        if (r == rows && c == cols && t == type_ && data && owner_) return;
        owner_.reset(new std::vector<uchar>((size_t)r * c));
        rows = r; cols = c; type_ = t; step = (size_t)c; data = owner_->data();
    }
    void release() { owner_.reset(); data = nullptr; rows = cols = 0; step = 0; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    bool isContinuous() const { return step == (size_t)cols; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    Mat row(int r) const { Mat m(*this); m.data = data + (size_t)r * step; m.rows = 1; return m; }
    Mat rowRange(int a, int b) const { Mat m(*this); m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.data = data + a; m.cols = b - a; return m; }
    Mat clone() const { Mat m(rows, cols, type_); for (int r = 0; r < rows; r++) memcpy(m.ptr(r), ptr(r), cols); return m; }
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;
private:
    int type_ = 0;
    std::shared_ptr<std::vector<uchar>> owner_;
};

// cv::InputArray / cv::OutputArray are "const _InputArray&" / "const _OutputArray&" proxies in OpenCV
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray() {}
    _OutputArray(Mat& m) { m_ = &m; }
    void create(int r, int c, int t) const { m_->create(r, c, t); }
    void release() const { if (m_) m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline const _InputArray& noArray() { static _InputArray a; return a; }

}  // namespace cv

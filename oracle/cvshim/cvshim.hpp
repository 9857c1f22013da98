// ORACLE (test infrastructure, NOT product code).
//
// A stand-in for the handful of OpenCV C++ declarations that the reference's ORBextractor.{h,cc} use, so that those two
// files can be compiled UNMODIFIED, from where they lie under /root/reference, into oracle/_ref/liborbref.so (OpenCV's
// C++ headers and libraries are not in this image; python-cv2 ships no headers).  Everything ORB-SLAM3-specific -- the
// cell loop, the iniTh/minTh fallback, DistributeOctTree with its std::list / std::sort, IC_Angle, the steered BRIEF
// sampling, the lapping-area assembly -- is then the reference's own object code.  The OpenCV primitives it calls
// (cv::resize INTER_LINEAR, copyMakeBorder REFLECT_101, cv::FAST with NMS, GaussianBlur 7x7 s=2, fastAtan2, cvRound)
// are implemented in cvshim.cpp by the restatements of orb_port.cpp, which the CPU tests pin bit for bit against the real
// OpenCV (python cv2) -- see tests/test_oracle_primitives.py.  Matrices are single-channel, 8-bit (32-bit float storage exists
// only so that the vendored DBoW2's FORB::toMat32F compiles).  `make ref` also compiles the vendored Thirdparty/DBoW2
// (FORB.cpp, BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp, TemplatedVocabulary.h) against this header; its YAML
// persistence (cv::FileStorage) and boost::serialization hooks are inert stubs -- vocabularies are loaded with DBoW2's own
// loadFromTextFile, as ORB-SLAM3 does (System.cc:116).
#pragma once
#include <algorithm>      // OpenCV's own headers pull these in; the reference relies on that (std::sort, assert)
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795

inline int cvRound(double v) { return (int)std::lrint(v); }        // round-half-even, as OpenCV's SSE2 / lrint paths
inline int cvRound(float v) { return (int)std::lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
    int x, y, width, height;
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

class _OutputArray;

class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;

    Mat() : rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, int type) : Mat() { create(r, c, type); }
    Mat(Size sz, int type) : Mat() { create(sz.height, sz.width, type); }
    Mat(int r, int c, int type, void* ext, size_t step_ = 0) : rows(r), cols(c), data((uchar*)ext), step(step_ ? step_ : (size_t)c * (type == CV_32F ? 4 : 1)), type_(type) {}

    void create(int r, int c, int type) {                        // keeps the buffer (and a ROI view) when the size already fits
        if (data && r == rows && c == cols && type == type_) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * elemSize();
        buf_ = allocate((size_t)r * step);
        data = buf_.get();
    }
    void release() { *this = Mat(); }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); if (m.data) std::memset(m.data, 0, (size_t)r * m.step); return m; }

    int type() const { return type_; }
    size_t elemSize() const { return type_ == CV_32F ? 4 : 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t step1() const { return step; }
    template <typename T> T& at(int r, int c) { return *(T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *(const T*)(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }

    Mat operator()(const Rect& r) const { Mat m(*this); m.data = data + (size_t)r.y * step + r.x; m.rows = r.height; m.cols = r.width; return m; }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat clone() const {
        Mat m(rows, cols, CV_8UC1);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols);
        return m;
    }
    void copyTo(const _OutputArray& dst) const;

private:
    // Uninitialised storage, recycled per thread: a fresh pyramid level per frame otherwise means an mmap + page faults +
    // munmap per level, which serialises the worker threads of the throughput baseline on the process's memory map.
    static std::shared_ptr<uchar> allocate(size_t n);
    std::shared_ptr<uchar> buf_;
    int type_ = CV_8UC1;
};

class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(&m) {}
    Mat getMat() const { return m_ ? *m_ : Mat(); }
    bool empty() const { return !m_ || m_->empty(); }
private:
    const Mat* m_;
};
class _OutputArray {
public:
    _OutputArray(Mat& m) : m_(&m) {}
    _OutputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}      // OpenCV has the same overload (temporaries such as m.row(i))
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void create(Size sz, int type) const { m_->create(sz.height, sz.width, type); }
    void release() const { *m_ = Mat(); }
    Mat getMat() const { return *m_; }
    Mat& getMatRef() const { return *m_; }
private:
    Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

inline void Mat::copyTo(const _OutputArray& dst) const {
    dst.create(rows, cols, CV_8UC1);
    Mat d = dst.getMat();
    for (int r = 0; r < rows; r++) std::memmove(d.data + (size_t)r * d.step, data + (size_t)r * step, (size_t)cols);
}

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };
enum { NORM_L1 = 2 };

float fastAtan2(float y, float x);
double norm(InputArray a, InputArray b, int normType);      // NORM_L1 of two 8-bit matrices of one size (Frame.cc:930)
void resize(InputArray src, OutputArray dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(InputArray src, OutputArray dst, int top, int bottom, int left, int right, int borderType);
void GaussianBlur(InputArray src, OutputArray dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
void FAST(InputArray image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);

// YAML persistence used by DBoW2's save()/load(): never opens here (DBoW2 then throws, nothing calls it)
class FileNode {
public:
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const std::string&) const { return FileNode(); }
    template <typename T> FileStorage& operator<<(const T&) { return *this; }
};

struct KeyPointsFilter {       // only ORBextractor::ComputeKeyPointsOld (dead code in the reference, :1101) calls this
    static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

}  // namespace cv

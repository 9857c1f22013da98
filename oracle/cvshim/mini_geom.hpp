// ORACLE / TEST INFRASTRUCTURE (not product code): the handful of Eigen / Sophus declarations that the motion-model
// ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) and the relocalisation overload (reference orb_slam3/src/ORBmatcher.cc:1676-2010) touch --
// Eigen::Vector2f / Vector3f with operator()(int), operator-, operator/, norm(), dot(), Matrix3f, Sophus::Sim3f with rotationMatrix() / translation() / scale(), SE3f composition, the pinhole epipolar test, Sophus::SE3f with inverse(), translation() and operator*(Vector3f).  Neither
// library exists in this image.  The geometry is host code on both sides of the comparison (the reference's cut-out body in
// oracle/_ref and the GPU-backed replacement in tests/host/ are compiled against THIS header with the same flags), so the floats
// they feed into the candidate scan are identical; what is compared is the scan and its decisions.
#pragma once
#include <cmath>

namespace Eigen {
struct Vector2f {
    float v[2] = {0, 0};
    Vector2f() {}
    Vector2f(float a, float b) { v[0] = a; v[1] = b; }
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
};
struct Vector3f {
    float v[3] = {0, 0, 0};
    Vector3f() {}
    Vector3f(float a, float b, float c) { v[0] = a; v[1] = b; v[2] = c; }
    float& operator()(int i) { return v[i]; }
    float operator()(int i) const { return v[i]; }
    Vector3f operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3f operator/(float d) const { return Vector3f(v[0] / d, v[1] / d, v[2] / d); }
    float norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    float dot(const Vector3f& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
};
struct Matrix3f {               // row major
    float m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float& operator()(int r, int c) { return m[3 * r + c]; }
    float operator()(int r, int c) const { return m[3 * r + c]; }
    Matrix3f operator*(const Matrix3f& o) const {
        Matrix3f r;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) r.m[3 * i + j] = m[3 * i] * o.m[j] + m[3 * i + 1] * o.m[3 + j] + m[3 * i + 2] * o.m[6 + j];
        return r;
    }
};
}  // namespace Eigen

namespace Sophus {
template <class T>
class SE3 {                     // rotation matrix (row major) + translation: x -> R x + t
public:
    T R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    T t[3] = {0, 0, 0};
    SE3() {}
    SE3(const T* r9, const T* t3) { for (int i = 0; i < 9; i++) R[i] = r9[i]; for (int i = 0; i < 3; i++) t[i] = t3[i]; }
    SE3(const Eigen::Matrix3f& r, const Eigen::Vector3f& tv) { for (int i = 0; i < 9; i++) R[i] = r.m[i]; for (int i = 0; i < 3; i++) t[i] = tv(i); }
    SE3 inverse() const {       // (R^T, -R^T t)
        SE3 o;
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) o.R[3 * i + j] = R[3 * j + i];
        for (int i = 0; i < 3; i++) o.t[i] = -(o.R[3 * i] * t[0] + o.R[3 * i + 1] * t[1] + o.R[3 * i + 2] * t[2]);
        return o;
    }
    Eigen::Vector3f translation() const { return Eigen::Vector3f(t[0], t[1], t[2]); }
    Eigen::Matrix3f rotationMatrix() const { Eigen::Matrix3f r; for (int i = 0; i < 9; i++) r.m[i] = R[i]; return r; }
    SE3 operator*(const SE3& o) const {      // x -> R (Ro x + to) + t
        SE3 r;
        for (int i = 0; i < 3; i++) {
            for (int j = 0; j < 3; j++) r.R[3 * i + j] = R[3 * i] * o.R[j] + R[3 * i + 1] * o.R[3 + j] + R[3 * i + 2] * o.R[6 + j];
            r.t[i] = R[3 * i] * o.t[0] + R[3 * i + 1] * o.t[1] + R[3 * i + 2] * o.t[2] + t[i];
        }
        return r;
    }
    Eigen::Vector3f operator*(const Eigen::Vector3f& p) const {
        return Eigen::Vector3f(R[0] * p(0) + R[1] * p(1) + R[2] * p(2) + t[0], R[3] * p(0) + R[4] * p(1) + R[5] * p(2) + t[1],
                               R[6] * p(0) + R[7] * p(1) + R[8] * p(2) + t[2]);
    }
};
typedef SE3<float> SE3f;
template <class T>
class Sim3 {                    // x -> s R x + t: the three accessors the Sim3 projection search uses (ORBmatcher.cc:435)
public:
    Eigen::Matrix3f R;
    Eigen::Vector3f t;
    T s = 1;
    Eigen::Matrix3f rotationMatrix() const { return R; }
    Eigen::Vector3f translation() const { return t; }
    T scale() const { return s; }
};
typedef Sim3<float> Sim3f;
}  // namespace Sophus

namespace ORB_SLAM3 {
class GeometricCamera {         // CameraModels/GeometricCamera.h:61-63: the one overload the cut function calls; pinhole (Pinhole.cpp:40-47)
public:
    float fx = 1, fy = 1, cx = 0, cy = 0;
    virtual ~GeometricCamera() {}
    virtual Eigen::Vector2f project(const Eigen::Vector3f& v3D) { return Eigen::Vector2f(fx * v3D(0) / v3D(2) + cx, fy * v3D(1) / v3D(2) + cy); }
    // Pinhole::epipolarConstrain (CameraModels/Pinhole.cpp:107-129): distance of kp2 to the epipolar line of kp1, F12 = K1^-T [t12]x R12 K2^-1
    // (a template on the key point type only so that this header needs no OpenCV declaration; the geometry is outside the matcher)
    template <class KP>
    bool epipolarConstrain(GeometricCamera* pCamera2, const KP& kp1, const KP& kp2, const Eigen::Matrix3f& R12, const Eigen::Vector3f& t12,
                           const float sigmaLevel, const float unc) {
        (void)sigmaLevel;
        Eigen::Matrix3f tx, K1tinv, K2inv;
        tx.m[0] = 0; tx.m[1] = -t12(2); tx.m[2] = t12(1); tx.m[3] = t12(2); tx.m[4] = 0; tx.m[5] = -t12(0); tx.m[6] = -t12(1); tx.m[7] = t12(0); tx.m[8] = 0;
        K1tinv.m[0] = 1 / fx; K1tinv.m[1] = 0; K1tinv.m[2] = 0; K1tinv.m[3] = 0; K1tinv.m[4] = 1 / fy; K1tinv.m[5] = 0;
        K1tinv.m[6] = -cx / fx; K1tinv.m[7] = -cy / fy; K1tinv.m[8] = 1;
        K2inv.m[0] = 1 / pCamera2->fx; K2inv.m[1] = 0; K2inv.m[2] = -pCamera2->cx / pCamera2->fx; K2inv.m[3] = 0; K2inv.m[4] = 1 / pCamera2->fy;
        K2inv.m[5] = -pCamera2->cy / pCamera2->fy; K2inv.m[6] = 0; K2inv.m[7] = 0; K2inv.m[8] = 1;
        const Eigen::Matrix3f F12 = K1tinv * tx * R12 * K2inv;
        const float a = kp1.pt.x * F12(0, 0) + kp1.pt.y * F12(1, 0) + F12(2, 0);
        const float b = kp1.pt.x * F12(0, 1) + kp1.pt.y * F12(1, 1) + F12(2, 1);
        const float c = kp1.pt.x * F12(0, 2) + kp1.pt.y * F12(1, 2) + F12(2, 2);
        const float num = a * kp2.pt.x + b * kp2.pt.y + c;
        const float den = a * a + b * b;
        if (den == 0) return false;
        const float dsqr = num * num / den;
        return dsqr < 3.84 * unc;
    }
};
}  // namespace ORB_SLAM3

// ORACLE (test infrastructure, NOT product code).
//
// CPU restatement ("port") of the ORB-SLAM3 front-end hot path of giltchcity/orb_slam3_ros, written
// in plain C++ with NO OpenCV dependency.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library.  The product (liborbb200.so) never
// links, loads or calls it.
//
// PARITY PINNING: the reference has no tests or golden vectors for this path.
// The arithmetic that lives in un-vendored OpenCV (resize, FAST, GaussianBlur, fastAtan2, BFMatcher) is
// restated here as closed integer / float32 formulas; oracle/orb_ref.py runs the SAME control flow through
// the real OpenCV primitives (python cv2) and tests/test_oracle_*.py require the two to agree bit-for-bit.
// EXTRACTOR rows: pinned to the reference itself -- oracle/_ref/liborbref.so is the reference's own
// ORBextractor.cc compiled unmodified (`make ref`, OpenCV declarations from cvshim/, primitives from this
// file) and tests/test_reference_source.py requires this port to reproduce it bit for bit.
// STEREO / GRID / DISTANCE / BoW rows: pinned the same way -- the vendored DBoW2 is compiled unmodified, and the
// definitions of Frame::ComputeStereoMatches, ComputeStereoFromRGBD, AssignFeaturesToGrid, PosInGrid,
// GetFeaturesInArea, ORBmatcher::DescriptorDistance, ComputeThreeMaxima and MapPoint::ComputeDistinctiveDescriptors
// are cut out of Frame.cc / ORBmatcher.cc / MapPoint.cc at build time and compiled inside stand-in classes
// (ref_cut_tu.cpp), as are SearchByProjection(Frame&, vector<MapPoint*>&, ..) and SearchByBoW(KeyFrame*, Frame&, ..)
// for end-to-end checks of the batched scans.  The other Search* loops stay line-cited restatements; OpenCV-defined results are pinned to cv2.
//
// Build: g++ -O3 -march=native -ffp-contract=off -shared -fPIC (see oracle/Makefile).
// -ffp-contract=off makes the un-fused float32 result the truth ("sin/cos and FMA").
//
// Every function cites the reference lines it follows (paths relative to /root/reference/).

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <thread>
#include <utility>
#include <vector>

namespace {

const int kPatchSize = 31;      // orb_slam3/src/ORBextractor.cc:71
const int kHalfPatch = 15;      // :72
const int kEdge = 19;           // :73  EDGE_THRESHOLD

// cvRound(float/double) on x86 = cvtss2si / cvtsd2si = round-half-to-even in the default mode.
inline int cv_round(double v) { return (int)std::lrint(v); }
inline int cv_round(float v) { return (int)std::lrintf(v); }
inline int cv_floor(float v) { int i = (int)v; return i - (i > v); }
inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }
inline int cv_ceil(float v) { int i = (int)v; return i + (i < v); }

// cv::borderInterpolate(p, len, BORDER_REFLECT_101)
inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

#include "pattern_table.inc"   // generated from oracle/bit_pattern_31.txt by oracle/Makefile

struct Image {
    int w = 0, h = 0;
    size_t stride = 0;
    std::vector<uint8_t> buf;   // bordered storage
    uint8_t* roi = nullptr;     // first pixel of the un-bordered region
    const uint8_t* row(int y) const { return roi + (ptrdiff_t)y * (ptrdiff_t)stride; }
};

// ---------------------------------------------------------------------------------------------
// cv::resize(src, dst, INTER_LINEAR) for CV_8UC1  (called at ORBextractor.cc:1183)
// OpenCV imgproc/resize.cpp: resizeGeneric_ with HResizeLinear<uchar,int,short,2048> and
// VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>.
// ---------------------------------------------------------------------------------------------
void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t ss, uint8_t* dst, int dw, int dh, size_t ds) {
    const double inv_sx = (double)dw / sw, inv_sy = (double)dh / sh;
    const double scale_x = 1. / inv_sx, scale_y = 1. / inv_sy;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> alpha(2 * dw), beta(2 * dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = (short)cv_round((1.f - fx) * 2048.f);
        alpha[2 * dx + 1] = (short)cv_round(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        beta[2 * dy] = (short)cv_round((1.f - fy) * 2048.f);
        beta[2 * dy + 1] = (short)cv_round(fy * 2048.f);
    }
    std::vector<int> r0(dw), r1(dw);
    auto hrow = [&](int sy, std::vector<int>& out) {
        sy = std::min(std::max(sy, 0), sh - 1);
        const uint8_t* S = src + (size_t)sy * ss;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            int s1 = S[std::min(sx + 1, sw - 1)];
            out[dx] = S[sx] * alpha[2 * dx] + s1 * alpha[2 * dx + 1];
        }
    };
    for (int dy = 0; dy < dh; dy++) {
        hrow(yofs[dy], r0);
        hrow(yofs[dy] + 1, r1);
        const int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * ds;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)std::min(std::max(v, 0), 255);
        }
    }
}

// cv::copyMakeBorder(..., BORDER_REFLECT_101) in place around the ROI (ORBextractor.cc:1185,1190)
void fill_border(Image& im) {
    const int B = kEdge;
    for (int y = -B; y < im.h + B; y++) {
        uint8_t* d = im.roi + (ptrdiff_t)y * (ptrdiff_t)im.stride;
        const uint8_t* s = im.row(reflect101(y, im.h));
        if (y < 0 || y >= im.h)
            for (int x = 0; x < im.w; x++) d[x] = s[x];
        for (int x = 1; x <= B; x++) {
            d[-x] = s[reflect101(-x, im.w)];
            d[im.w - 1 + x] = s[reflect101(im.w - 1 + x, im.w)];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cv::FAST(img, kps, threshold, true) == FAST_t<16> (features2d/fast.cpp, fast_score.cpp)
// called per cell at ORBextractor.cc:826,845.  Returns (x, y, response) in raster order.
// ---------------------------------------------------------------------------------------------
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// M = max over both polarities and all 16 start positions of the minimum over a 9-long arc.
inline int fast9_arc_score(const uint8_t* p, ptrdiff_t stride) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[kRingDy[k] * stride + kRingDx[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = -256;
    for (int s = 0; s < 16; s++) {
        int mn = d[s], mx = d[s];
        for (int k = 1; k < 9; k++) { mn = std::min(mn, d[s + k]); mx = std::max(mx, d[s + k]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best;
}

struct RawKey { float x, y, response; };

// true iff some 9-long arc of the ring is entirely brighter than v+t or entirely darker than v-t
inline bool fast9_is_corner(const uint8_t* p, ptrdiff_t stride, int t) {
    const int v = p[0], hi = v + t, lo = v - t;
    // any 9-arc contains ring pixel 0 or 8, and 4 or 12 (cv::FAST's high-speed test)
    const int p0 = p[3 * stride], p8 = p[-3 * stride];
    bool dark = (p0 < lo) | (p8 < lo), bright = (p0 > hi) | (p8 > hi);
    if (!(dark | bright)) return false;
    const int p4 = p[3], p12 = p[-3];
    dark &= (p4 < lo) | (p12 < lo);
    bright &= (p4 > hi) | (p12 > hi);
    if (!(dark | bright)) return false;
    unsigned md = 0, mb = 0;
    for (int k = 0; k < 16; k++) {
        const int q = p[kRingDy[k] * stride + kRingDx[k]];
        md |= (unsigned)(q < lo) << k;
        mb |= (unsigned)(q > hi) << k;
    }
    auto arc9 = [](unsigned m) {
        m |= m << 16;
        unsigned r = m & (m >> 1);
        r &= r >> 2;
        r &= r >> 4;
        r &= m >> 8;
        return (r & 0xffffu) != 0;
    };
    return (dark && arc9(md)) || (bright && arc9(mb));
}

void fast9_nms(const uint8_t* img, int w, int h, size_t stride, int threshold, std::vector<RawKey>& out) {
    out.clear();
    if (w < 7 || h < 7) return;
    threshold = std::min(std::max(threshold, 0), 255);
    // three rolling score rows would do; cells are tiny (<= 76x76) so a full map is simplest
    static thread_local std::vector<int> score;
    score.assign((size_t)w * h, 0);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const uint8_t* p = img + (size_t)y * stride + x;
            if (!fast9_is_corner(p, (ptrdiff_t)stride, threshold)) continue;
            // corner <=> arc score M > threshold; response = M - 1.  A corner whose response is 0 (M = 1 at
            // threshold 0) can never win the strict '>' NMS and equals a non-corner as a neighbour.
            score[(size_t)y * w + x] = fast9_arc_score(p, (ptrdiff_t)stride) - 1;
        }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const int s = score[(size_t)y * w + x];
            if (s <= 0) continue;
            const int* r0 = &score[(size_t)(y - 1) * w + x];
            const int* r1 = r0 + w;
            const int* r2 = r1 + w;
            if (s > r0[-1] && s > r0[0] && s > r0[1] && s > r1[-1] && s > r1[1] && s > r2[-1] && s > r2[0] && s > r2[1])
                out.push_back({(float)x, (float)y, (float)s});
        }
}

// ---------------------------------------------------------------------------------------------
// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 (ORBextractor.cc:1133)
// OpenCV fixed-point path (smooth.simd.hpp, ufixedpoint16 8.8 kernel).
// ---------------------------------------------------------------------------------------------
const int kGauss7[7] = {18, 34, 48, 56, 48, 34, 18};

void gaussian7_u8(const uint8_t* src, int w, int h, size_t ss, uint8_t* dst, size_t ds) {
    static thread_local std::vector<uint16_t> hbuf;
    hbuf.resize((size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* S = src + (size_t)y * ss;
        uint16_t* H = &hbuf[(size_t)y * w];
        for (int x = 0; x < w; x++) {
            if (x >= 3 && x < w - 3) {
                H[x] = (uint16_t)(18 * (S[x - 3] + S[x + 3]) + 34 * (S[x - 2] + S[x + 2]) + 48 * (S[x - 1] + S[x + 1]) + 56 * S[x]);
            } else {
                unsigned acc = 0;
                for (int k = -3; k <= 3; k++) acc += kGauss7[k + 3] * S[reflect101(x + k, w)];
                H[x] = (uint16_t)acc;
            }
        }
    }
    for (int y = 0; y < h; y++) {
        uint8_t* D = dst + (size_t)y * ds;
        const uint16_t* r[7];
        for (int k = -3; k <= 3; k++) r[k + 3] = &hbuf[(size_t)reflect101(y + k, h) * w];
        for (int x = 0; x < w; x++) {
            const unsigned acc = 18u * (r[0][x] + r[6][x]) + 34u * (r[1][x] + r[5][x]) + 48u * (r[2][x] + r[4][x]) + 56u * r[3][x];
            D[x] = (uint8_t)((acc + (1u << 15)) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// cv::fastAtan2(y, x) (core/mathfuncs_core.simd.hpp, scalar atan_f32), used at ORBextractor.cc:102
// ---------------------------------------------------------------------------------------------
float fast_atan2(float y, float x) {
    const float scale = (float)(180 / 3.141592653589793238462643383279502884197169399375105820974944);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------------
// DistributeOctTree (ORBextractor.cc:555-779) with ExtractorNode::DivideNode (:480-536) and
// compareNodes (:538-553).  Real std::list and real libstdc++ std::sort: the tie order of the
// unstable sort decides which nodes are split in the "last mile" phase.
// Keys are referred to by index into `in`; a node keeps its keys in arrival order.
// ---------------------------------------------------------------------------------------------
struct QNode {
    int x0, y0, x1, y1;                  // UL=(x0,y0) UR=(x1,y0) BL=(x0,y1) BR=(x1,y1)
    std::vector<int> keys;
    bool no_more = false;
    std::list<QNode>::iterator self;
};

void divide_node(const QNode& p, const std::vector<RawKey>& in, QNode c[4]) {
    const int halfX = (int)std::ceil(static_cast<float>(p.x1 - p.x0) / 2);   // :482
    const int halfY = (int)std::ceil(static_cast<float>(p.y1 - p.y0) / 2);   // :483
    const int mx = p.x0 + halfX, my = p.y0 + halfY;
    c[0] = {p.x0, p.y0, mx, my, {}, false, {}};     // n1 :486-489
    c[1] = {mx, p.y0, p.x1, my, {}, false, {}};     // n2 :492-495
    c[2] = {p.x0, my, mx, p.y1, {}, false, {}};     // n3 :498-501
    c[3] = {mx, my, p.x1, p.y1, {}, false, {}};     // n4 :504-507
    for (int k : p.keys) {                          // :511-525
        const RawKey& kp = in[k];
        if (kp.x < mx) { if (kp.y < my) c[0].keys.push_back(k); else c[2].keys.push_back(k); }
        else if (kp.y < my) c[1].keys.push_back(k);
        else c[3].keys.push_back(k);
    }
    for (int i = 0; i < 4; i++) if (c[i].keys.size() == 1) c[i].no_more = true;   // :527-534
}

typedef std::pair<int, QNode*> SizeNode;
bool compare_nodes(SizeNode& a, SizeNode& b) {      // :538-553
    if (a.first < b.first) return true;
    if (a.first > b.first) return false;
    return a.second->x0 < b.second->x0;
}

// returns 0 on success, -2 when the reference would hit undefined behaviour (nIni == 0)
int distribute_octtree(const std::vector<RawKey>& in, int minX, int maxX, int minY, int maxY, int N,
                       std::vector<int>& out) {
    out.clear();
    const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));   // :559
    if (nIni <= 0) return -2;
    const float hX = static_cast<float>(maxX - minX) / nIni;                              // :561
    std::list<QNode> nodes;
    std::vector<QNode*> ini(nIni);
    for (int i = 0; i < nIni; i++) {                                                      // :568-579
        QNode n;
        n.x0 = (int)(hX * static_cast<float>(i));
        n.x1 = (int)(hX * static_cast<float>(i + 1));
        n.y0 = 0;
        n.y1 = maxY - minY;
        nodes.push_back(n);
        ini[i] = &nodes.back();
    }
    for (size_t i = 0; i < in.size(); i++) {                                              // :582-586
        size_t slot = (size_t)(in[i].x / hX);
        if (slot >= (size_t)nIni) return -2;   // out-of-range write in the reference
        ini[slot]->keys.push_back((int)i);
    }
    for (auto it = nodes.begin(); it != nodes.end();) {                                   // :588-601
        if (it->keys.size() == 1) { it->no_more = true; ++it; }
        else if (it->keys.empty()) it = nodes.erase(it);
        else ++it;
    }
    bool finish = false;
    std::vector<SizeNode> pending;
    auto add_children = [&](QNode c[4], int* nToExpand) {                                 // :637-676 / :707-742
        for (int i = 0; i < 4; i++) {
            if (c[i].keys.empty()) continue;
            nodes.push_front(c[i]);
            if (c[i].keys.size() > 1) {
                if (nToExpand) ++*nToExpand;
                pending.push_back(std::make_pair((int)c[i].keys.size(), &nodes.front()));
                nodes.front().self = nodes.begin();
            }
        }
    };
    while (!finish) {                                                                     // :610
        const int prevSize = (int)nodes.size();
        int nToExpand = 0;
        pending.clear();
        for (auto it = nodes.begin(); it != nodes.end();) {                               // :622-681
            if (it->no_more) { ++it; continue; }
            QNode c[4];
            divide_node(*it, in, c);
            add_children(c, &nToExpand);
            it = nodes.erase(it);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {                    // :685
            finish = true;
        } else if ((int)nodes.size() + nToExpand * 3 > N) {                               // :689
            while (!finish) {
                const int prev2 = (int)nodes.size();
                std::vector<SizeNode> prevPending = pending;                              // :697
                pending.clear();
                std::sort(prevPending.begin(), prevPending.end(), compare_nodes);         // :700
                for (int j = (int)prevPending.size() - 1; j >= 0; j--) {                  // :701
                    QNode c[4];
                    divide_node(*prevPending[j].second, in, c);
                    add_children(c, nullptr);
                    nodes.erase(prevPending[j].second->self);                             // :744
                    if ((int)nodes.size() >= N) break;                                    // :746
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prev2) finish = true;  // :750
            }
        }
    }
    for (const QNode& n : nodes) {                                                        // :757-776
        int best = n.keys[0];
        float maxResp = in[best].response;
        for (size_t k = 1; k < n.keys.size(); k++)
            if (in[n.keys[k]].response > maxResp) { best = n.keys[k]; maxResp = in[best].response; }
        out.push_back(best);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// IC_Angle (ORBextractor.cc:76-103)
// ---------------------------------------------------------------------------------------------
float ic_angle(const uint8_t* center, ptrdiff_t step, const int* umax) {
    int m01 = 0, m10 = 0;
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * center[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int plus = center[u + v * step], minus = center[u - v * step];
            vsum += plus - minus;
            m10 += u * (plus + minus);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

// computeOrbDescriptor (ORBextractor.cc:107-146).  cos/sin are the float overloads (glibc cosf/sinf),
// float products/sums un-fused, cvRound = half-to-even.
void orb_descriptor(const uint8_t* center, ptrdiff_t step, float kp_angle, uint8_t* desc) {
    const float factorPI = (float)(3.141592653589793238462643383279502884197169399375105820974944 / 180.f);
    const float angle = kp_angle * factorPI;
    const float a = std::cos(angle), b = std::sin(angle);
    const int8_t* pat = &kPattern[0][0];
    for (int i = 0; i < 32; i++) {
        int val = 0;
        for (int j = 0; j < 8; j++, pat += 4) {
            const float x0 = pat[0], y0 = pat[1], x1 = pat[2], y1 = pat[3];
            const int t0 = center[cv_round(x0 * b + y0 * a) * step + cv_round(x0 * a - y0 * b)];
            const int t1 = center[cv_round(x1 * b + y1 * a) * step + cv_round(x1 * a - y1 * b)];
            val |= (t0 < t1) << j;
        }
        desc[i] = (uint8_t)val;
    }
}

// ORBmatcher::DescriptorDistance (ORBmatcher.cc:2058-2074): SWAR popcount over 8 x int32
inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t pa, pb;
        memcpy(&pa, a + 4 * i, 4);
        memcpy(&pb, b + 4 * i, 4);
        uint32_t v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

struct PortKP { float x, y, size, angle, response; int octave; };

// ---------------------------------------------------------------------------------------------
// The extractor (ORBextractor.cc:409-469 ctor, :781-896 keypoints, :1086-1195 operator()/pyramid)
// ---------------------------------------------------------------------------------------------
struct Extractor {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;                       // ORBextractor.h:96 -- a double initialised from a float
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> featPerLevel, umax;
    std::vector<Image> pyr, blurred;
    std::vector<std::vector<RawKey>> raw;     // vToDistributeKeys per level (debug / stage parity)
    std::vector<std::vector<PortKP>> sel;     // allKeypoints per level, after orientation

    Extractor(int nf, float sf, int nl, int ini, int mn) : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf) {
        scale.resize(nl); sigma2.resize(nl); invScale.resize(nl); invSigma2.resize(nl);
        scale[0] = 1.0f; sigma2[0] = 1.0f;
        for (int i = 1; i < nl; i++) { scale[i] = scale[i - 1] * scaleFactor; sigma2[i] = scale[i] * scale[i]; }   // :418-422
        for (int i = 0; i < nl; i++) { invScale[i] = 1.0f / scale[i]; invSigma2[i] = 1.0f / sigma2[i]; }           // :426-430
        featPerLevel.resize(nl);
        float factor = 1.0f / scaleFactor;                                                                         // :435
        float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));            // :436
        int sum = 0;
        for (int l = 0; l < nl - 1; l++) { featPerLevel[l] = cv_round(nDesired); sum += featPerLevel[l]; nDesired *= factor; }   // :439-444
        featPerLevel[nl - 1] = std::max(nfeatures - sum, 0);                                                       // :445
        umax.resize(kHalfPatch + 1);                                                                               // :453-468
        int v, v0, vmax = cv_floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
        int vmin = cv_ceil(kHalfPatch * std::sqrt(2.f) / 2);
        const double hp2 = kHalfPatch * kHalfPatch;
        for (v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
        for (v = kHalfPatch, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
        pyr.resize(nl); blurred.resize(nl); raw.resize(nl); sel.resize(nl);
    }

    void alloc(Image& im, int w, int h, int border) {
        im.w = w; im.h = h; im.stride = (size_t)w + 2 * border;
        im.buf.assign(im.stride * (size_t)(h + 2 * border), 0);
        im.roi = im.buf.data() + (size_t)border * im.stride + border;
    }

    void compute_pyramid(const uint8_t* img, int w, int h, size_t stride) {                                       // :1170-1195
        for (int l = 0; l < nlevels; l++) {
            const float s = invScale[l];
            const int lw = cv_round((float)w * s), lh = cv_round((float)h * s);
            alloc(pyr[l], lw, lh, kEdge);
            if (l != 0) resize_linear_u8(pyr[l - 1].roi, pyr[l - 1].w, pyr[l - 1].h, pyr[l - 1].stride, pyr[l].roi, lw, lh, pyr[l].stride);
            else for (int y = 0; y < h; y++) memcpy(pyr[0].roi + (size_t)y * pyr[0].stride, img + (size_t)y * stride, w);
            fill_border(pyr[l]);
        }
    }

    int compute_keypoints() {                                                                                      // :781-896
        const float W = 35;
        for (int level = 0; level < nlevels; ++level) {
            const Image& im = pyr[level];
            const int minBX = kEdge - 3, minBY = minBX, maxBX = im.w - kEdge + 3, maxBY = im.h - kEdge + 3;
            std::vector<RawKey>& toDist = raw[level];
            toDist.clear();
            const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
            const int nCols = (int)(width / W), nRows = (int)(height / W);
            if (nCols <= 0 || nRows <= 0) {
                // the reference's loops below do not execute (wCell / hCell = ceil(x / 0) are never used) and DistributeOctTree
                // of no keys returns nothing: the level has no keypoints.  With a non-positive span :559 yields a negative or
                // undefined root count (vector::resize throws); level 0 is kept out as well (an image without one cell).
                if (level == 0 || width <= 0 || height <= 0) return -2;
                sel[level].clear();
                continue;
            }
            const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
            std::vector<RawKey> cell;
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(minBY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = (float)maxBY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(minBX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = (float)maxBX;
                    const uint8_t* roi = im.row((int)iniY) + (int)iniX;
                    const int rw = (int)maxX - (int)iniX, rh = (int)maxY - (int)iniY;
                    fast9_nms(roi, rw, rh, im.stride, iniTh, cell);                                                // :826
                    if (cell.empty()) fast9_nms(roi, rw, rh, im.stride, minTh, cell);                             // :843-846
                    for (RawKey k : cell) { k.x += j * wCell; k.y += i * hCell; toDist.push_back(k); }            // :863-868
                }
            }
            std::vector<int> picked;
            int rc = distribute_octtree(toDist, minBX, maxBX, minBY, maxBY, featPerLevel[level], picked);          // :877
            if (rc) return rc;
            const int scaledPatch = (int)(kPatchSize * scale[level]);                                              // :880
            std::vector<PortKP>& kps = sel[level];
            kps.clear();
            for (int idx : picked) {                                                                               // :884-890
                PortKP kp;
                kp.x = toDist[idx].x + minBX; kp.y = toDist[idx].y + minBY;
                kp.size = (float)scaledPatch; kp.angle = -1; kp.response = toDist[idx].response; kp.octave = level;
                kps.push_back(kp);
            }
        }
        for (int level = 0; level < nlevels; ++level)                                                              // :894-895
            for (PortKP& kp : sel[level])
                kp.angle = ic_angle(pyr[level].row(cv_round(kp.y)) + cv_round(kp.x), (ptrdiff_t)pyr[level].stride, umax.data());
        return 0;
    }

    // operator() (:1086-1168)
    int extract(const uint8_t* img, int w, int h, size_t stride, int lap0, int lap1, PortKP* outK, uint8_t* outD, int cap,
                int* nOut, int* monoOut) {
        *nOut = 0; *monoOut = 0;
        if (!img || w <= 0 || h <= 0) return -1;                                                                   // :1090
        for (int l = 0; l < nlevels; l++) {          // geometry the reference cannot handle, rejected before any work: a level without
            const int lw = cv_round((float)w * invScale[l]), lh = cv_round((float)h * invScale[l]);      // pixels (cv::resize asserts
            if (lw < 1 || lh < 1) return -2;                                                              // dsize.area() > 0, :1183) ...
            if (lw - 2 * (kEdge - 3) <= 0 || lh - 2 * (kEdge - 3) <= 0) return -2;                        // ... or inside the FAST border (:559)
        }
        compute_pyramid(img, w, h, stride);
        int rc = compute_keypoints();
        if (rc) return rc;
        int n = 0;
        for (int l = 0; l < nlevels; l++) n += (int)sel[l].size();
        if (n > cap) return -3;
        int mono = 0, stereo = n - 1;
        for (int l = 0; l < nlevels; l++) {
            std::vector<PortKP>& kps = sel[l];
            if (kps.empty()) continue;
            alloc(blurred[l], pyr[l].w, pyr[l].h, 0);
            gaussian7_u8(pyr[l].roi, pyr[l].w, pyr[l].h, pyr[l].stride, blurred[l].roi, blurred[l].stride);      // :1132-1133
            const float s = scale[l];
            for (PortKP& kp : kps) {
                uint8_t d[32];
                orb_descriptor(blurred[l].row(cv_round(kp.y)) + cv_round(kp.x), (ptrdiff_t)blurred[l].stride, kp.angle, d);
                PortKP o = kp;
                if (l != 0) { o.x *= s; o.y *= s; }                                                                // :1149-1151
                int pos;
                if (o.x >= lap0 && o.x <= lap1) pos = stereo--; else pos = mono++;                                 // :1153-1162
                outK[pos] = o;
                memcpy(outD + (size_t)pos * 32, d, 32);
            }
        }
        *nOut = n; *monoOut = mono;
        return 0;
    }
};

}  // namespace

extern "C" {

void* port_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return new Extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void port_destroy(void* h) { delete (Extractor*)h; }

void port_tables(void* h, float* scale, float* inv, float* sig2, float* invsig2, int* nfeat, int* umax) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; i++) {
        scale[i] = e->scale[i]; inv[i] = e->invScale[i]; sig2[i] = e->sigma2[i]; invsig2[i] = e->invSigma2[i];
        nfeat[i] = e->featPerLevel[i];
    }
    for (int i = 0; i <= kHalfPatch; i++) umax[i] = e->umax[i];
}

int port_extract(void* h, const uint8_t* img, int w, int hh, size_t stride, int lap0, int lap1, PortKP* kps, uint8_t* desc,
                 int cap, int* nOut, int* monoOut) {
    return ((Extractor*)h)->extract(img, w, hh, stride, lap0, lap1, kps, desc, cap, nOut, monoOut);
}

// All-core throughput: frames are independent, one Extractor per worker thread (the reference uses one extractor
// instance per concurrently processed image, Frame.cc:122-125).  counts[f] = {n, mono}; kps/desc: cap per frame.
int port_extract_batch(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh, const uint8_t* imgs, int nframes, int w,
                       int hh, size_t rowStride, size_t frameStride, int lap0, int lap1, PortKP* kps, uint8_t* desc, int cap,
                       int32_t* counts, int nthreads) {
    nthreads = std::max(1, std::min(nthreads, nframes));
    std::vector<int> rcs(nthreads, 0);
    auto work = [&](int t) {
        Extractor e(nfeatures, scaleFactor, nlevels, iniTh, minTh);
        std::vector<PortKP> k(cap);
        std::vector<uint8_t> d((size_t)cap * 32);
        for (int f = t; f < nframes; f += nthreads) {
            int n = 0, mono = 0;
            int rc = e.extract(imgs + (size_t)f * frameStride, w, hh, rowStride, lap0, lap1, k.data(), d.data(), cap, &n, &mono);
            if (rc) { rcs[t] = rc; return; }
            counts[2 * f] = n; counts[2 * f + 1] = mono;
            if (kps) memcpy(kps + (size_t)f * cap, k.data(), sizeof(PortKP) * n);
            if (desc) memcpy(desc + (size_t)f * cap * 32, d.data(), (size_t)n * 32);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
    for (auto& t : th) t.join();
    for (int rc : rcs) if (rc) return rc;
    return 0;
}

// bordered=1: pointer to the (w+38)x(h+38) buffer origin, else to the ROI origin
int port_level(void* h, int level, int blurred, int bordered, const uint8_t** ptr, int* w, int* hh, size_t* stride) {
    Extractor* e = (Extractor*)h;
    if (level < 0 || level >= e->nlevels) return -1;
    const Image& im = blurred ? e->blurred[level] : e->pyr[level];
    if (im.buf.empty()) return -1;
    *w = im.w; *hh = im.h; *stride = im.stride;
    *ptr = (bordered && !blurred) ? im.buf.data() : im.roi;
    return 0;
}

int port_raw_count(void* h, int level) { return (int)((Extractor*)h)->raw[level].size(); }
void port_raw_keys(void* h, int level, float* xyr) {
    const auto& v = ((Extractor*)h)->raw[level];
    for (size_t i = 0; i < v.size(); i++) { xyr[3 * i] = v[i].x; xyr[3 * i + 1] = v[i].y; xyr[3 * i + 2] = v[i].response; }
}
int port_sel_count(void* h, int level) { return (int)((Extractor*)h)->sel[level].size(); }
void port_sel_keys(void* h, int level, PortKP* out) {
    const auto& v = ((Extractor*)h)->sel[level];
    if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(PortKP));
}

// ---- primitives exposed for the cv2 cross-checks -------------------------------------------------
void port_resize_linear(const uint8_t* src, int sw, int sh, size_t ss, uint8_t* dst, int dw, int dh, size_t ds) {
    resize_linear_u8(src, sw, sh, ss, dst, dw, dh, ds);
}
int port_fast9(const uint8_t* img, int w, int h, size_t stride, int th, float* xyr, int cap) {
    std::vector<RawKey> v;
    fast9_nms(img, w, h, stride, th, v);
    int n = (int)std::min<size_t>(v.size(), (size_t)cap);
    for (int i = 0; i < n; i++) { xyr[3 * i] = v[i].x; xyr[3 * i + 1] = v[i].y; xyr[3 * i + 2] = v[i].response; }
    return (int)v.size();
}
void port_gaussian7(const uint8_t* src, int w, int h, size_t ss, uint8_t* dst, size_t ds) { gaussian7_u8(src, w, h, ss, dst, ds); }
float port_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void port_fast_atan2_array(const float* y, const float* x, float* out, int n) { for (int i = 0; i < n; i++) out[i] = fast_atan2(y[i], x[i]); }
int port_cv_round(float v) { return cv_round(v); }

int port_distribute(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, int* outIdx, int cap) {
    std::vector<RawKey> in(n);
    for (int i = 0; i < n; i++) in[i] = {xyr[3 * i], xyr[3 * i + 1], xyr[3 * i + 2]};
    std::vector<int> out;
    int rc = distribute_octtree(in, minX, maxX, minY, maxY, N, out);
    if (rc) return rc;
    if ((int)out.size() > cap) return -3;
    for (size_t i = 0; i < out.size(); i++) outIdx[i] = out[i];
    return (int)out.size();
}

// std::sort with compareNodes on (size, x0) pairs; returns the permutation (for the introsort model test)
void port_sort_nodes(const int* sizes, const int* x0s, int n, int* perm) {
    std::vector<QNode> store(n);
    std::vector<SizeNode> v(n);
    for (int i = 0; i < n; i++) { store[i].x0 = x0s[i]; v[i] = std::make_pair(sizes[i], &store[i]); }
    std::sort(v.begin(), v.end(), compare_nodes);
    for (int i = 0; i < n; i++) perm[i] = (int)(v[i].second - store.data());
}

float port_ic_angle(const uint8_t* img, size_t stride, int x, int y, const int* umax) {
    return ic_angle(img + (size_t)y * stride + x, (ptrdiff_t)stride, umax);
}
// IC_Angle for n key points at once (xy = n x {x, y} level coordinates, cvRound'ed like :80); used by oracle/cv2_baseline.py
void port_ic_angles(const uint8_t* img, size_t stride, const float* xy, int n, const int* umax, float* out) {
    for (int i = 0; i < n; i++) out[i] = ic_angle(img + (size_t)cv_round(xy[2 * i + 1]) * stride + cv_round(xy[2 * i]), (ptrdiff_t)stride, umax);
}
void port_descriptors(const uint8_t* blurred, size_t stride, const float* xya, int n, uint8_t* desc) {
    for (int i = 0; i < n; i++)
        orb_descriptor(blurred + (size_t)cv_round(xya[3 * i + 1]) * stride + cv_round(xya[3 * i]), (ptrdiff_t)stride, xya[3 * i + 2],
                       desc + (size_t)i * 32);
}
int port_hamming(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }

// ---- brute-force 2-NN (cv::BFMatcher(NORM_HAMMING).knnMatch(q, db, 2), Frame.cc:1144) -----------------
// idx2/dist2: nq x 2; missing neighbours are (-1, INT_MAX).  Ties -> lowest train index first.
void port_knn2(const uint8_t* q, int nq, const uint8_t* db, long long nd, int32_t* idx2, int32_t* dist2, int nthreads) {
    auto work = [&](int q0, int q1) {
        for (int i = q0; i < q1; i++) {
            int d1 = INT_MAX, d2 = INT_MAX, i1 = -1, i2 = -1;
            const uint64_t* a = (const uint64_t*)(q + (size_t)i * 32);
            uint64_t a0, a1, a2, a3;
            memcpy(&a0, a, 8); memcpy(&a1, a + 1, 8); memcpy(&a2, a + 2, 8); memcpy(&a3, a + 3, 8);
            for (long long j = 0; j < nd; j++) {
                uint64_t b[4];
                memcpy(b, db + (size_t)j * 32, 32);
                int d = __builtin_popcountll(a0 ^ b[0]) + __builtin_popcountll(a1 ^ b[1]) + __builtin_popcountll(a2 ^ b[2]) +
                        __builtin_popcountll(a3 ^ b[3]);
                if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = (int)j; }
                else if (d < d2) { d2 = d; i2 = (int)j; }
            }
            idx2[2 * i] = i1; idx2[2 * i + 1] = i2; dist2[2 * i] = d1; dist2[2 * i + 1] = d2;
        }
    };
    if (nthreads <= 1) { work(0, nq); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(work, (int)((long long)nq * t / nthreads), (int)((long long)nq * (t + 1) / nthreads));
    for (auto& t : th) t.join();
}

// ---- best / second-best scan over candidate lists (ORBmatcher.cc:77-120 and the other M2 scans) ----------
// CSR: query i scans cand[rowptr[i] .. rowptr[i+1]) (indices into `train`), strict '<', start value init.
// out4[i] = {bestDist, bestIdx, secondDist, secondIdx}
void port_best2_csr(const uint8_t* q, int nq, const uint8_t* train, const int32_t* cand, const int32_t* rowptr, int init, int32_t* out4) {
    for (int i = 0; i < nq; i++) {
        int d1 = init, d2 = init, i1 = -1, i2 = -1;
        for (int c = rowptr[i]; c < rowptr[i + 1]; c++) {
            int d = descriptor_distance(q + (size_t)i * 32, train + (size_t)cand[c] * 32);
            if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = cand[c]; }
            else if (d < d2) { d2 = d; i2 = cand[c]; }
        }
        out4[4 * i] = d1; out4[4 * i + 1] = i1; out4[4 * i + 2] = d2; out4[4 * i + 3] = i2;
    }
}

// ---- Frame::AssignFeaturesToGrid (Frame.cc:385-416) + GetFeaturesInArea (:657-723) + the best/second scan of
// ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&) (ORBmatcher.cc:71-120), batched over queries --------------
// kps: n x {x, y} (undistorted), oct: n octaves, train: n x 32.  grid4 = {mnMinX, mnMinY, mfGridElementWidthInv,
// mfGridElementHeightInv}.  queries: nq x {x, y, r, projXR}; qlev: nq x {minLevel, maxLevel}; qdesc: nq x 32.
// skip (optional, n bytes): keypoint already holds an observed MapPoint (:88-90).  uRight (optional, n floats): stereo
// coordinate, checked against projXR when > 0 (:92-97).  out4[i] = {bestDist, bestIdx, secondDist, secondIdx}.
void port_search_area_best2(const float* kps, const int32_t* oct, const uint8_t* train, int n, const float* grid4, const float* queries,
                            const int32_t* qlev, const uint8_t* qdesc, int nq, const uint8_t* skip, const float* uRight, int init,
                            int32_t* out4) {
    const int COLS = 64, ROWS = 48;                                                      // Frame.h:44-45
    const float minX = grid4[0], minY = grid4[1], invW = grid4[2], invH = grid4[3];
    std::vector<std::vector<size_t>> grid((size_t)COLS * ROWS);
    for (int i = 0; i < n; i++) {                                                         // :403-415, PosInGrid :725-735
        const int px = (int)std::round((kps[2 * i] - minX) * invW), py = (int)std::round((kps[2 * i + 1] - minY) * invH);
        if (px < 0 || px >= COLS || py < 0 || py >= ROWS) continue;
        grid[(size_t)px * ROWS + py].push_back((size_t)i);
    }
    for (int q = 0; q < nq; q++) {
        const float x = queries[4 * q], y = queries[4 * q + 1], r = queries[4 * q + 2], xr = queries[4 * q + 3];
        const int minLevel = qlev[2 * q], maxLevel = qlev[2 * q + 1];
        std::vector<size_t> idxs;
        do {                                                                              // :657-723
            const float factorX = r, factorY = r;
            const int nMinCellX = std::max(0, (int)std::floor((x - minX - factorX) * invW));
            if (nMinCellX >= COLS) break;
            const int nMaxCellX = std::min(COLS - 1, (int)std::ceil((x - minX + factorX) * invW));
            if (nMaxCellX < 0) break;
            const int nMinCellY = std::max(0, (int)std::floor((y - minY - factorY) * invH));
            if (nMinCellY >= ROWS) break;
            const int nMaxCellY = std::min(ROWS - 1, (int)std::ceil((y - minY + factorY) * invH));
            if (nMaxCellY < 0) break;
            const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
            for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
                for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                    for (size_t j : grid[(size_t)ix * ROWS + iy]) {
                        if (bCheckLevels) {
                            if (oct[j] < minLevel) continue;
                            if (maxLevel >= 0 && oct[j] > maxLevel) continue;
                        }
                        const float distx = kps[2 * j] - x, disty = kps[2 * j + 1] - y;
                        if (std::fabs(distx) < factorX && std::fabs(disty) < factorY) idxs.push_back(j);
                    }
        } while (false);
        int bestDist = init, bestDist2 = init, bestIdx = -1, bestIdx2 = -1;               // ORBmatcher.cc:77-120
        for (size_t idx : idxs) {
            if (skip && skip[idx]) continue;
            if (uRight && uRight[idx] > 0) {
                const float er = std::fabs(xr - uRight[idx]);
                if (er > r) continue;
            }
            const int dist = descriptor_distance(qdesc + (size_t)q * 32, train + idx * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestIdx2 = bestIdx; bestDist = dist; bestIdx = (int)idx; }
            else if (dist < bestDist2) { bestDist2 = dist; bestIdx2 = (int)idx; }
        }
        out4[4 * q] = bestDist; out4[4 * q + 1] = bestIdx; out4[4 * q + 2] = bestDist2; out4[4 * q + 3] = bestIdx2;
    }
}

// ---- DBoW2 TemplatedVocabulary<FORB>::transform(features, BowVector, FeatureVector, levelsup)
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1213, :1230-1275; BowVector.cpp:30-80; FeatureVector.cpp:31-45), the call
// behind Frame::ComputeBoW (Frame.cc:738-745).  TF / TF_IDF weighting (addWeight).  Tree as flat arrays, see orbb200.h.
// Outputs packed per call: bow ids/values ascending; fv nodes ascending with their feature lists concatenated.
int port_bow_transform(int nnodes, const int32_t* childBegin, const int32_t* childCount, const int32_t* childList, const uint8_t* nodeDesc,
                       const double* nodeWeight, const int32_t* nodeWord, int depth, const uint8_t* desc, int n, int levelsup, int norm,
                       int32_t* bowId, double* bowVal, int32_t* fvNode, int32_t* fvStart, int32_t* fvFeat, int32_t* counts) {
    (void)nnodes;
    std::map<int, double> bow;                      // BowVector: std::map<WordId, WordValue>
    std::map<int, std::vector<unsigned>> fv;        // FeatureVector
    const int nidLevel = depth - levelsup;
    for (int f = 0; f < n; f++) {
        int nid = 0, finalId = 0, level = 0;         // :1240-1243
        do {                                          // :1245-1268
            ++level;
            const int beg = childBegin[finalId], cnt = childCount[finalId];
            int cur = childList[beg];
            double bestD = (double)descriptor_distance(desc + (size_t)f * 32, nodeDesc + (size_t)cur * 32);
            for (int c = 1; c < cnt; c++) {
                const int id = childList[beg + c];
                const double d = (double)descriptor_distance(desc + (size_t)f * 32, nodeDesc + (size_t)id * 32);
                if (d < bestD) { bestD = d; cur = id; }
            }
            finalId = cur;
            if (level == nidLevel) nid = finalId;
        } while (childCount[finalId] > 0);
        const double w = nodeWeight[finalId];
        if (w > 0) {                                  // :1169-1173
            auto it = bow.lower_bound(nodeWord[finalId]);
            if (it != bow.end() && !(nodeWord[finalId] < it->first)) it->second += w;
            else bow.insert(it, std::make_pair(nodeWord[finalId], w));
            fv[nid].push_back((unsigned)f);
        }
    }
    if (!norm && !bow.empty()) {                      // :1176-1182 (scoring without normalisation)
        const double nd = (double)bow.size();
        for (auto& kv : bow) kv.second /= nd;
    }
    if (norm) {                                       // BowVector::normalize
        double nrm = 0.0;
        if (norm == 1) for (auto& kv : bow) nrm += std::fabs(kv.second);
        else { for (auto& kv : bow) nrm += kv.second * kv.second; nrm = std::sqrt(nrm); }
        if (nrm > 0.0) for (auto& kv : bow) kv.second /= nrm;
    }
    int k = 0;
    for (auto& kv : bow) { bowId[k] = kv.first; bowVal[k] = kv.second; k++; }
    counts[0] = k;
    int m = 0, pos = 0;
    for (auto& kv : fv) {
        fvNode[m] = kv.first; fvStart[m] = pos; m++;
        for (unsigned f : kv.second) fvFeat[pos++] = (int)f;
    }
    counts[1] = m; counts[2] = pos;
    return 0;
}

// ---- MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:329-403), batched over groups ------------------------
// group g = descriptors desc[rowptr[g] .. rowptr[g+1]); best[g] = index (within the group) of the descriptor with the
// least median distance to the others (first minimum; median = sorted[(size_t)(0.5*(N-1))], self distance 0 included);
// -1 for an empty group (the reference returns early, :367-368)
void port_distinctive(const uint8_t* desc, const int32_t* rowptr, int ngroups, int32_t* best) {
    for (int g = 0; g < ngroups; g++) {
        const int N = rowptr[g + 1] - rowptr[g];
        const uint8_t* d = desc + (size_t)rowptr[g] * 32;
        if (N <= 0) { best[g] = -1; continue; }
        std::vector<float> dist((size_t)N * N);
        for (int i = 0; i < N; i++) {                                                     // :372-383
            dist[(size_t)i * N + i] = 0;
            for (int j = i + 1; j < N; j++) {
                const int dij = descriptor_distance(d + (size_t)i * 32, d + (size_t)j * 32);
                dist[(size_t)i * N + j] = (float)dij;
                dist[(size_t)j * N + i] = (float)dij;
            }
        }
        int bestMedian = INT_MAX, bestIdx = 0;                                            // :386-399
        for (int i = 0; i < N; i++) {
            std::vector<int> v(dist.begin() + (size_t)i * N, dist.begin() + (size_t)(i + 1) * N);
            std::sort(v.begin(), v.end());
            const int median = v[(size_t)(0.5 * (N - 1))];
            if (median < bestMedian) { bestMedian = median; bestIdx = i; }
        }
        best[g] = bestIdx;
    }
}

// ---- cv::cvtColor(src, dst, COLOR_{RGB,BGR,RGBA,BGRA}2GRAY) for 8-bit images (Tracking.cc:1498-1525) ----------------
// OpenCV 4.x fixed point: (R*9798 + G*19235 + B*3735 + 2^14) >> 15   (color_rgb.simd.hpp, RGB2Gray<uchar>)
void port_gray(const uint8_t* src, int w, int h, size_t stride, int channels, int rgb, uint8_t* dst, size_t dstride) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const uint8_t* p = src + (size_t)y * stride + (size_t)x * channels;
            const int r = rgb ? p[0] : p[2], g = p[1], b = rgb ? p[2] : p[0];
            dst[(size_t)y * dstride + x] = (uint8_t)((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15);
        }
}

// ---- Frame::ComputeStereoMatches (Frame.cc:811-981) ----------------------------------------------------
// hL/hR: extractors that have just processed the left/right image (their pyramids are read, :908,:923).
// Returns the number of surviving matches, or -1 on error.  sadOut (optional): best SAD per left kp (-1 none).
int port_stereo(void* hL, void* hR, const PortKP* kL, const uint8_t* dL, int nL, const PortKP* kR, const uint8_t* dR, int nR,
                float mbf, float mb, float* uRight, float* depth, int32_t* bestR, int32_t* sadOut) {
    Extractor* eL = (Extractor*)hL;
    Extractor* eR = (Extractor*)hR;
    for (int i = 0; i < nL; i++) { uRight[i] = -1.0f; depth[i] = -1.0f; if (bestR) bestR[i] = -1; if (sadOut) sadOut[i] = -1; }
    const int thOrbDist = (100 + 50) / 2;                                                 // :816
    const int nRows = eL->pyr[0].h;                                                       // :818
    std::vector<std::vector<size_t>> rowIdx(nRows);
    for (int iR = 0; iR < nR; iR++) {                                                     // :828-838
        const float kpY = kR[iR].y;
        const float r = 2.0f * eL->scale[kR[iR].octave];
        const int maxr = (int)std::ceil(kpY + r);
        const int minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) rowIdx[yi].push_back(iR);   // reference indexes unchecked
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;                                   // :841-843
    std::vector<std::pair<int, int>> distIdx;
    for (int iL = 0; iL < nL; iL++) {                                                     // :849
        const PortKP& kpL = kL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const size_t row = (size_t)vL;
        if (row >= (size_t)nRows) continue;
        const std::vector<size_t>& cands = rowIdx[row];
        if (cands.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = 100;                                                               // :867 TH_HIGH
        size_t bestIdxR = 0;
        for (size_t iC = 0; iC < cands.size(); iC++) {                                    // :873-894
            const size_t iR = cands[iC];
            const PortKP& kpR = kR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(dL + (size_t)iL * 32, dR + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {                                                       // :897
            if (bestR) bestR[iL] = (int)bestIdxR;
            const float uR0 = kR[bestIdxR].x;
            const float sf = eL->invScale[kpL.octave];
            const float scaleduL = std::round(kpL.x * sf);
            const float scaledvL = std::round(kpL.y * sf);
            const float scaleduR0 = std::round(uR0 * sf);
            const int w = 5, L = 5;
            const Image& pl = eL->pyr[kpL.octave];
            const Image& pr = eR->pyr[kpL.octave];
            int bestSad = INT_MAX, bestinc = 0;
            float dists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - w;
            const float endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= pr.w) continue;                                       // :918
            const int y0 = (int)(scaledvL - w), xl0 = (int)(scaleduL - w);
            for (int inc = -L; inc <= +L; inc++) {                                        // :921-933
                const int xr0 = (int)(scaleduR0 + inc - w);
                int sad = 0;
                for (int yy = 0; yy < 2 * w + 1; yy++) {
                    const uint8_t* a = pl.row(y0 + yy) + xl0;
                    const uint8_t* b = pr.row(y0 + yy) + xr0;
                    for (int xx = 0; xx < 2 * w + 1; xx++) sad += std::abs((int)a[xx] - (int)b[xx]);
                }
                const float dist = (float)sad;
                if (dist < bestSad) { bestSad = (int)dist; bestinc = inc; }
                dists[L + inc] = dist;
            }
            if (bestinc == -L || bestinc == L) continue;                                  // :935
            const float d1 = dists[L + bestinc - 1], d2 = dists[L + bestinc], d3 = dists[L + bestinc + 1];
            const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));              // :943
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = eL->scale[kpL.octave] * ((float)scaleduR0 + (float)bestinc + deltaR);   // :949
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {                                  // :953-963
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                distIdx.push_back(std::pair<int, int>(bestSad, iL));
                if (sadOut) sadOut[iL] = bestSad;
            }
        }
    }
    if (distIdx.empty()) return 0;     // the reference reads vDistIdx[0] of an empty vector here (UB)
    std::sort(distIdx.begin(), distIdx.end());                                            // :967
    const float median = distIdx[distIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;                                            // :969
    int kept = (int)distIdx.size();
    for (int i = (int)distIdx.size() - 1; i >= 0; i--) {                                  // :971-980
        if (distIdx[i].first < thDist) break;
        uRight[distIdx[i].second] = -1;
        depth[distIdx[i].second] = -1;
        kept--;
    }
    return kept;
}

}  // extern "C"

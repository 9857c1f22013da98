// liborbb200.so -- ORB extraction for B200 (sm_100a).  Hand-written CUDA; no CPU fallback.
//
// Replaces ORB_SLAM3::ORBextractor (reference orb_slam3/src/ORBextractor.cc) behind the C ABI of
// include/orbb200.h.  Stages (one launch each per BATCH of frames, all on the handle's stream):
//   k_pyr_level0_v / k_pyr_resize_t x (levels-1) / k_pyr_apron16   ComputePyramid   :1170-1195  (orbb_pyr.cuh: cv::resize fixed
//                                 point + reflect-101 apron)
//   k_fast_cell                   per-cell cv::FAST + ini/min threshold fallback   :787-872   (orbb_fast.cuh)
//   k_octree                      DistributeOctTree         :555-779    (array-rebuild formulation, see
//                                                                         tests/models/octree_array_model.cpp)
//   k_blur                        cv::GaussianBlur 7x7 s=2  :1133       (8.8 fixed point; on a side stream beside k_octree)
//   k_assemble                    output ordering / scaling / lapping split   :1105-1167 (batches; the prologue of the next kernel in
//                                 a call with a few frames)
//   k_orient_desc32               IC_Angle + fastAtan2 + computeOrbDescriptor :76-146
// plus the input-side rows: k_gray (cvtColor) and k_remap (cv::remap rectification, orbb_rectify.cuh).
//
// Compiled with -fmad=false: the un-fused float32 result is the specification.
#include <limits.h>
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "introsort.cuh"
#include "orbb_internal.cuh"

namespace orbb {

thread_local std::string g_lastError;

int set_err(orbb_extractor* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    g_lastError = buf;
    return code;
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len) {
    // cv::borderInterpolate(BORDER_REFLECT_101); |p| never exceeds len by more than the 19-px apron here
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// ------------------------------------------------------------------------------------------------
// K0 ("next" row): cv::cvtColor(COLOR_{RGB,BGR,RGBA,BGRA}2GRAY) for 8-bit input (Tracking.cc:1498-1525), OpenCV 4.x fixed
// point.  4 output pixels per thread into the tightly packed gray staging frame.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gray(const uint8_t* __restrict__ src, size_t rowStride, size_t frameStride, int channels,
                                              int rgbOrder, uint8_t* __restrict__ dst, int w, int h) {
    const int word = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int frame = blockIdx.z;
    if (word * 4 >= w || y >= h) return;
    const uint8_t* p = src + (size_t)frame * frameStride + (size_t)y * rowStride + (size_t)word * 4 * channels;
    uint8_t* d = dst + ((size_t)frame * h + y) * w + word * 4;
    const int n = min(4, w - word * 4);
    for (int k = 0; k < n; k++, p += channels) {
        const int r = rgbOrder ? p[0] : p[2], g = p[1], b = rgbOrder ? p[2] : p[0];
        d[k] = (uint8_t)((r * 9798 + g * 19235 + b * 3735 + (1 << 14)) >> 15);
    }
}

// ------------------------------------------------------------------------------------------------
// K1: pyramid (ComputePyramid :1170-1195).  Three kernels: level 0 = copy of the source into its slab, level l =
// cv::resize of level l-1 (image pixels only, 4 per thread), then ONE launch that fills the 19-px reflect-101 apron
// of every level from the level's own pixels (copyMakeBorder BORDER_REFLECT_101 [| BORDER_ISOLATED]).
// ------------------------------------------------------------------------------------------------
// cv::resize INTER_LINEAR 8UC1 (imgproc/resize.cpp, HResizeLinear/VResizeLinear fixed point): tables hold (source
// index, packed int16 coefficient pair) per destination column / row, computed on the host exactly as OpenCV does
// (build_plan).  horizontal: S[sx]*a0 + S[sx+1]*a1 (x2048); vertical: (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2.
__global__ void __launch_bounds__(256) k_pyr_resize(const Plan* __restrict__ P, Bufs B, int level) {
    const LevelPlan& L = P->lv[level];
    const LevelPlan& S = P->lv[level - 1];
    const int word = blockIdx.x * 32 + (threadIdx.x & 31);
    const int dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int frame = blockIdx.z;
    pdl_launch_dependents();
    pdl_wait();
    if (word * 4 >= L.w || dy >= L.h) return;
    const uint8_t* sroi = B.pyr + (size_t)frame * P->pyrStride + S.roiOff;
    const int2 ty = __ldg(B.tab + L.tabY + dy);
    const int sy0 = min(max(ty.x, 0), S.h - 1), sy1 = min(max(ty.x + 1, 0), S.h - 1);
    const int b0 = (short)(ty.y & 0xffff), b1 = ty.y >> 16;
    const uint8_t* r0 = sroi + (size_t)sy0 * S.pitch;
    const uint8_t* r1 = sroi + (size_t)sy1 * S.pitch;
    const int2* tx = B.tab + L.tabX + word * 4;     // the table is padded to a multiple of 4 entries
    unsigned packed = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int2 t = __ldg(tx + k);
        const int sx = t.x, sx1 = min(sx + 1, S.w - 1);
        const int a0 = (short)(t.y & 0xffff), a1 = t.y >> 16;
        const int h0 = r0[sx] * a0 + r0[sx1] * a1;
        const int h1 = r1[sx] * a0 + r1[sx1] * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        packed |= (unsigned)min(max(v, 0), 255) << (8 * k);
    }
    uint8_t* d = B.pyr + (size_t)frame * P->pyrStride + L.roiOff + (size_t)dy * L.pitch;
    reinterpret_cast<unsigned*>(d)[word] = packed;
}

// The same arithmetic for a free-standing image pair ("next" row, input side): the cv::resize(im, imToFeed, newImSize) that
// System::TrackStereo / TrackRGBD / TrackMonocular apply when the settings ask for another image size (reference
// orb_slam3/src/System.cc:241-244, :312-318, :383-388).  tab = [dw (padded to 4)] column entries, then [dh] row entries.
__global__ void __launch_bounds__(256) k_resize_image(const int2* __restrict__ tab, int tabY, const uint8_t* __restrict__ src, size_t srcStride,
                                                      size_t srcFrameStride, int sw, int sh, uint8_t* __restrict__ dst, int dw, int dh) {
    const int word = blockIdx.x * 32 + (threadIdx.x & 31);
    const int dy = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int frame = blockIdx.z;
    if (word * 4 >= dw || dy >= dh) return;
    const uint8_t* s = src + (size_t)frame * srcFrameStride;
    const int2 ty = __ldg(tab + tabY + dy);
    const int sy0 = min(max(ty.x, 0), sh - 1), sy1 = min(max(ty.x + 1, 0), sh - 1);
    const int b0 = (short)(ty.y & 0xffff), b1 = ty.y >> 16;
    const uint8_t* r0 = s + (size_t)sy0 * srcStride;
    const uint8_t* r1 = s + (size_t)sy1 * srcStride;
    const int2* tx = tab + word * 4;
    uint8_t* d = dst + ((size_t)frame * dh + dy) * dw + word * 4;     // tightly packed destination frames
    const int n = min(4, dw - word * 4);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (k < n) {
            const int2 t = __ldg(tx + k);
            const int sx = t.x, sx1 = min(sx + 1, sw - 1);
            const int a0 = (short)(t.y & 0xffff), a1 = t.y >> 16;
            const int h0 = __ldg(r0 + sx) * a0 + __ldg(r0 + sx1) * a1;
            const int h1 = __ldg(r1 + sx) * a0 + __ldg(r1 + sx1) * a1;
            const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
            d[k] = (uint8_t)min(max(v, 0), 255);
        }
    }
}

#include "orbb_pyr.cuh"
#include "orbb_rectify.cuh"

// ------------------------------------------------------------------------------------------------
// K5a: GaussianBlur 7x7 sigma 2, OpenCV's 8.8 fixed-point path: taps {18,34,48,56,48,34,18}/256, horizontal pass
// in 16 bit, vertical pass 32 bit, one rounding (acc + 2^15) >> 16; BORDER_REFLECT_101 at the IMAGE edge
// (the reference blurs a clone of the level, so the pyramid apron is not used) (:1132-1133).
// ------------------------------------------------------------------------------------------------
// No shared memory: 3 aligned word loads per row (served by L1; neighbouring lanes share two of them), the 7-tap
// horizontal sums as two IDP.4A each.  The kernel does its own reflect-101 at the image edge, so the pyramid needs no apron
// while a frame is processed (nothing else in the pipeline reads beyond the image; the reference's 19-px apron is built when
// the pyramid is handed out, ensure_full_apron): rows outside the image are read through reflected row pointers (first / last
// strip only), the word left of the image and the bytes right of it are byte permutes of the two nearest image words (first /
// last two word columns only).
#ifndef ORBB_BLUR_STRIP
#define ORBB_BLUR_STRIP 32
#endif
constexpr int BLUR_THREADS = 128, BLUR_STRIP = ORBB_BLUR_STRIP, BLUR_STRIP_EDGE = 8;

__device__ __forceinline__ unsigned hsum7(unsigned a, unsigned b) {
    // a = bytes x-3..x (taps 18,34,48,56), b = bytes x+1..x+4 (taps 48,34,18,0); result <= 255 * 256 fits 16 bits
    return __dp4a(a, 0x38302212u, __dp4a(b, 0x00122230u, 0u));
}

// horizontal sums of the 4 columns of one word (w1) of one input row; w0 / w2 = the words left / right of it
__device__ __forceinline__ void blur_hwords(unsigned w0, unsigned w1, unsigned w2, unsigned (&h)[4]) {
    h[0] = hsum7(__funnelshift_r(w0, w1, 8), __funnelshift_r(w1, w2, 8));
    h[1] = hsum7(__funnelshift_r(w0, w1, 16), __funnelshift_r(w1, w2, 16));
    h[2] = hsum7(__funnelshift_r(w0, w1, 24), __funnelshift_r(w1, w2, 24));
    h[3] = hsum7(w1, w2);
}

// One thread owns 4 adjacent columns (one word) of a strip of rows.  The horizontal sums of two consecutive input rows are
// kept PACKED in one register (u16 | u16 << 16), so the vertical 7-tap pass of one pixel is four IDP.2A (u16 x u8 dot
// products with the tap pairs) over the four row pairs that cover its window -- for even and for odd output rows with
// different tap constants -- instead of 4 multiplies + 3 adds on unpacked sums.
//
// Two launches, side by side on two streams:
//   EDGE = false  the interior: output rows 3 .. h-4 of the word columns 1 .. nwords-3 in 32-row strips.  Their 7 x 7 windows stay
//                 inside the image (the bytes of the last word beyond the image width are never read): no reflection code at all.
//   EDGE = true   the frame around it, where BORDER_REFLECT_101 applies, in short strips (a few hundred threads per frame):
//                 the word columns 0, nwords-2, nwords-1 over all rows (8-row strips; reflected bytes come from byte permutes of the
//                 two nearest image words), and output rows 0..2 / h-3..h-1 of the interior columns (reflected row pointers).
constexpr int BLUR_BAND = 3;                                 // output rows at the top / bottom whose windows leave the image


template <bool EDGE>
__global__ void __launch_bounds__(BLUR_THREADS, EDGE ? 8 : 12) k_blur(const Plan* __restrict__ P, Bufs B) {
    constexpr int STRIP = EDGE ? BLUR_STRIP_EDGE : BLUR_STRIP;
    const int frame = blockIdx.y;
    int level = 0;
    while (level + 1 < P->nlevels && (int)blockIdx.x >= (EDGE ? P->lv[level + 1].blurEdgeBase : P->lv[level + 1].blurTileBase)) level++;
    const LevelPlan& L = P->lv[level];
    // (the reference skips levels without keypoints, :1128; here the blur runs beside the detector, before that is known)
    const int nwords = L.blurTilesX, h = L.h;                // (nwords >= 17, h >= 67: a level holds at least one 35-px cell + borders)
    const int item = (blockIdx.x - (EDGE ? L.blurEdgeBase : L.blurTileBase)) * BLUR_THREADS + threadIdx.x;
    int wc, y0, rows, colFix = 3;                            // colFix: 0 / 1 / 2 = first / last-but-one / last word column, 3 = none
    if (!EDGE) {
        const int ncols = nwords - 3;
        if (item >= ncols * L.blurTilesY) return;
        const int strip = item / ncols;
        wc = item - strip * ncols + 1;
        y0 = BLUR_BAND + strip * STRIP;
        rows = min(STRIP, h - BLUR_BAND - y0);
    } else {
        const int nstrips = (h + STRIP - 1) / STRIP, nColItems = 3 * nstrips;
        if (item >= nColItems + 2 * (nwords - 3)) return;
        if (item < nColItems) {
            const int strip = item / 3;
            colFix = item - 3 * strip;
            wc = colFix == 0 ? 0 : nwords - 3 + colFix;
            y0 = strip * STRIP;
            rows = min(STRIP, h - y0);
        } else {
            const int j = item - nColItems, band = j >= nwords - 3;
            wc = 1 + j - band * (nwords - 3);
            y0 = band ? h - BLUR_BAND : 0;
            rows = BLUR_BAND;
        }
    }
    const int pitch = L.pitch, bpitch = L.bpitch;
    // input row i of the strip = level row y0 - 3 + i; output row i uses input rows i .. i + 6
    const uint8_t* col = B.pyr + (size_t)frame * P->pyrStride + L.roiOff + 4 * wc;
    const uint8_t* src = col + (ptrdiff_t)(y0 - 3) * pitch;  // (interior: walked two rows at a time)
    int yin = y0 - 3;                                        // (edge: next input row, reflected into the image)
    // column fix-ups (BORDER_REFLECT_101: column -k = column k, column w-1+k = column w-1-k).  r = valid bytes of the last word.
    //   colFix 0: the word left of the first word column = bytes {c4, c3, c2, c1} of the words {c0..c3}, {c4..c7}
    //   colFix 2: last word column, window {w0, w1}: byte j >= r of w1 = window byte 2r+2-j; byte j of the word right of it = 2r-2-j
    //   colFix 1: the one before it, window {w1, w2}: byte j >= r of w2 = window byte 2r+2-j
    unsigned selA = 0x7654u, selB = 0x7654u;
    if (EDGE && (colFix == 1 || colFix == 2)) {
        const int r = L.w - 4 * (nwords - 1);
        selA = selB = 0;
        for (int j = 0; j < 4; j++) {
            const unsigned keep = (unsigned)(j < r ? 4 + j : 2 * r + 2 - j);
            if (colFix == 2) {
                selA |= keep << (4 * j);
                selB |= (unsigned)max(2 * r - 2 - j, 0) << (4 * j);
            } else {
                selB |= keep << (4 * j);
            }
        }
    }
    unsigned pk[4][4];                                       // 4 row pairs x 4 columns
    unsigned raw[6];                                         // the next row pair's words, loaded one iteration ahead
    auto fetch = [&]() {
        const unsigned *r0, *r1;
        if (EDGE) {
            int ya = yin, yb = yin + 1;
            ya = ya < 0 ? -ya : (ya >= h ? 2 * h - 2 - ya : ya);
            yb = yb < 0 ? -yb : (yb >= h ? 2 * h - 2 - yb : yb);
            r0 = reinterpret_cast<const unsigned*>(col + (ptrdiff_t)ya * pitch);
            r1 = reinterpret_cast<const unsigned*>(col + (ptrdiff_t)yb * pitch);
            yin += 2;
        } else {
            r0 = reinterpret_cast<const unsigned*>(src);
            r1 = reinterpret_cast<const unsigned*>(src + pitch);
            src += 2 * pitch;
        }
        raw[0] = __ldg(r0 - 1); raw[1] = __ldg(r0); raw[2] = __ldg(r0 + 1);
        raw[3] = __ldg(r1 - 1); raw[4] = __ldg(r1); raw[5] = __ldg(r1 + 1);
        if (EDGE) {
            if (colFix == 0) {
                raw[0] = __byte_perm(raw[1], raw[2], 0x1234);
                raw[3] = __byte_perm(raw[4], raw[5], 0x1234);
            } else if (colFix == 2) {
                raw[2] = __byte_perm(raw[0], raw[1], selB); raw[1] = __byte_perm(raw[0], raw[1], selA);
                raw[5] = __byte_perm(raw[3], raw[4], selB); raw[4] = __byte_perm(raw[3], raw[4], selA);
            } else if (colFix == 1) {
                raw[2] = __byte_perm(raw[1], raw[2], selB);
                raw[5] = __byte_perm(raw[4], raw[5], selB);
            }
        }
    };
    auto pack = [&](unsigned (&dst)[4]) {
        unsigned h0[4], h1[4];
        blur_hwords(raw[0], raw[1], raw[2], h0);
        blur_hwords(raw[3], raw[4], raw[5], h1);
#pragma unroll
        for (int k = 0; k < 4; k++) dst[k] = h0[k] | (h1[k] << 16);
    };
    uint8_t* out = B.blur + (size_t)frame * P->blurStride + L.blurOff + 4 * wc + (size_t)y0 * bpitch;
#pragma unroll
    for (int j = 0; j < 3; j++) { fetch(); pack(pk[j]); }
    fetch();
#pragma unroll 1                                              // (fully unrolled: 2.3 x the code, no faster)
    for (int mo = 0; mo < STRIP / 2; mo += 4) {               // (4 = the period of the row-pair ring pk[])
        if (2 * mo >= rows) break;
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
            const int m = mo + mi;
            if (2 * m < rows) {                               // (the interior reads rows up to y0 + rows + 3 <= h: inside the image or its first apron row)
                pack(pk[(mi + 3) & 3]);
                if (2 * (m + 1) < rows) fetch();              // in flight while this iteration's vertical pass runs
                unsigned e[4], o[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned p0 = pk[mi & 3][k], p1 = pk[(mi + 1) & 3][k], p2 = pk[(mi + 2) & 3][k], p3 = pk[(mi + 3) & 3][k];
                    // even output row 2m: rows 2m..2m+6 = (18,34) (48,56) (48,34) (18,-)
                    e[k] = __dp2a_lo(p0, 0x2212u, __dp2a_lo(p1, 0x3830u, __dp2a_lo(p2, 0x2230u, __dp2a_lo(p3, 0x0012u, 32768u))));
                    // odd output row 2m+1: rows 2m+1..2m+7 = (-,18) (34,48) (56,48) (34,18)
                    o[k] = __dp2a_lo(p0, 0x1200u, __dp2a_lo(p1, 0x3022u, __dp2a_lo(p2, 0x3038u, __dp2a_lo(p3, 0x1222u, 32768u))));
                }
                // byte 2 of every accumulator = (acc + 2^15) >> 16
                *reinterpret_cast<unsigned*>(out) = __byte_perm(__byte_perm(e[0], e[1], 0x0062), __byte_perm(e[2], e[3], 0x0062), 0x5410);
                if (2 * m + 1 < rows)
                    *reinterpret_cast<unsigned*>(out + bpitch) = __byte_perm(__byte_perm(o[0], o[1], 0x0062), __byte_perm(o[2], o[3], 0x0062), 0x5410);
                out += 2 * bpitch;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K2: FAST-9/16 per 35-px cell with NMS and the iniTh -> minTh fallback (:805-872; cv::FAST == FAST_t<16>): orbb_fast.cuh
// ------------------------------------------------------------------------------------------------
#include "orbb_fast.cuh"

// ------------------------------------------------------------------------------------------------
// K3: DistributeOctTree, one CTA per (frame, level).
// ------------------------------------------------------------------------------------------------
constexpr int OT_THREADS = 128;          // CTA size of the quadtree kernel for ordinary levels ...
constexpr int OT_THREADS_BIG = 512;      // ... and for levels with many cells (4K-class images): more warps split nodes at once
constexpr int OT_BIG_CELLS = 1000;
constexpr int OT_SORT_SMEM = 1024;
constexpr int OT_SMEM_KEYS = 8192;        // keys per shared-memory buffer of the quadtree of a call with a few frames (2 x 64 KB)
constexpr int OT_SMEM_NODES = 512;        // ... and nodes per shared-memory node array
constexpr int OT_SMEM_BYTES = OT_SMEM_KEYS * 16 + OT_SMEM_NODES * (2 * 16 + 16 + 2 * 4 + 4 + 1);

__device__ __forceinline__ int node_count(const QNode& n) { return n.cntbuf & 0x7fffffff; }
__device__ __forceinline__ int node_buf(const QNode& n) { return (unsigned)n.cntbuf >> 31; }

__device__ __forceinline__ int key_quadrant(u64 k, int mx, int my) {
    const int x = (int)(k & 0xffff), y = (int)((k >> 16) & 0xffff);
    return x < mx ? (y < my ? 0 : 2) : (y < my ? 1 : 3);          // :514-524
}

// ExtractorNode::DivideNode key assignment (:511-525) by one warp: stable 4-way partition of the node's key
// segment into the other ping-pong buffer; child sizes -> *out.
__device__ void split_node_warp(const QNode nd, u64* k0, u64* k1, int4* out, int lane) {
    const int cnt = node_count(nd);
    const u64* src = node_buf(nd) ? k1 : k0;
    u64* dst = node_buf(nd) ? k0 : k1;
    const int mx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1);             // ceil(w/2) :482
    const int my = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);             // :483
    const unsigned lt = (1u << lane) - 1;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    if (cnt > 32) {
        for (int base = 0; base < cnt; base += 32) {
            const int i = base + lane;
            const int q = i < cnt ? key_quadrant(src[nd.start + i], mx, my) : 4;
            c0 += __popc(__ballot_sync(0xffffffffu, q == 0));
            c1 += __popc(__ballot_sync(0xffffffffu, q == 1));
            c2 += __popc(__ballot_sync(0xffffffffu, q == 2));
            c3 += __popc(__ballot_sync(0xffffffffu, q == 3));
        }
    }
    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int i = base + lane;
        const bool valid = i < cnt;
        const u64 k = valid ? src[nd.start + i] : 0;
        const int q = valid ? key_quadrant(k, mx, my) : 4;
        const unsigned b0 = __ballot_sync(0xffffffffu, q == 0), b1 = __ballot_sync(0xffffffffu, q == 1);
        const unsigned b2 = __ballot_sync(0xffffffffu, q == 2), b3 = __ballot_sync(0xffffffffu, q == 3);
        if (cnt <= 32) { c0 = __popc(b0); c1 = __popc(b1); c2 = __popc(b2); c3 = __popc(b3); }
        if (valid) {
            const unsigned bm = q == 0 ? b0 : q == 1 ? b1 : q == 2 ? b2 : b3;
            const int off = q == 0 ? r0 : q == 1 ? c0 + r1 : q == 2 ? c0 + c1 + r2 : c0 + c1 + c2 + r3;
            dst[nd.start + off + __popc(bm & lt)] = k;
        }
        r0 += __popc(b0); r1 += __popc(b1); r2 += __popc(b2); r3 += __popc(b3);
    }
    if (lane == 0) *out = make_int4(c0, c1, c2, c3);
}

// The same by the whole CTA for a node with many keys (the first passes over a large level have fewer nodes than warps):
// every warp takes a contiguous part of the keys -- four keys per lane are loaded before they are used, so four loads are
// in flight -- and the parts' counts give each warp its output offsets.  sPart: int[warps][4].
constexpr int OT_BIG_NODE = 4096, OT_BIG_NODE_LATENCY = 512, OT_BIG_LIST = 16;
__device__ void split_node_cta(const QNode nd, u64* k0, u64* k1, int4* out, int (*sPart)[4], int tid, int nw) {
    const int lane = tid & 31, warp = tid >> 5;
    const int cnt = node_count(nd);
    const u64* src = (node_buf(nd) ? k1 : k0) + nd.start;
    u64* dst = (node_buf(nd) ? k0 : k1) + nd.start;
    const int mx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1), my = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
    const int part = ((cnt + nw - 1) / nw + 127) & ~127;
    const int i0 = min(warp * part, cnt), i1 = min(i0 + part, cnt);
    const unsigned lt = (1u << lane) - 1;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int base = i0; base < i1; base += 128) {
        u64 k[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { const int i = base + 32 * j + lane; k[j] = i < i1 ? src[i] : 0; }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int q = base + 32 * j + lane < i1 ? key_quadrant(k[j], mx, my) : 4;
            c0 += __popc(__ballot_sync(0xffffffffu, q == 0));
            c1 += __popc(__ballot_sync(0xffffffffu, q == 1));
            c2 += __popc(__ballot_sync(0xffffffffu, q == 2));
            c3 += __popc(__ballot_sync(0xffffffffu, q == 3));
        }
    }
    if (lane == 0) { sPart[warp][0] = c0; sPart[warp][1] = c1; sPart[warp][2] = c2; sPart[warp][3] = c3; }
    __syncthreads();
    int t[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0};                   // totals, and what the warps before this one hold
    for (int w = 0; w < nw; w++)
#pragma unroll
        for (int q = 0; q < 4; q++) { const int v = sPart[w][q]; t[q] += v; if (w < warp) b[q] += v; }
    int o0 = b[0], o1 = t[0] + b[1], o2 = t[0] + t[1] + b[2], o3 = t[0] + t[1] + t[2] + b[3];
    for (int base = i0; base < i1; base += 128) {
        u64 k[4];
#pragma unroll
        for (int j = 0; j < 4; j++) { const int i = base + 32 * j + lane; k[j] = i < i1 ? src[i] : 0; }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const bool valid = base + 32 * j + lane < i1;
            const int q = valid ? key_quadrant(k[j], mx, my) : 4;
            const unsigned b0 = __ballot_sync(0xffffffffu, q == 0), b1 = __ballot_sync(0xffffffffu, q == 1);
            const unsigned b2 = __ballot_sync(0xffffffffu, q == 2), b3 = __ballot_sync(0xffffffffu, q == 3);
            if (valid) {
                const unsigned bm = q == 0 ? b0 : q == 1 ? b1 : q == 2 ? b2 : b3;
                const int off = q == 0 ? o0 : q == 1 ? o1 : q == 2 ? o2 : o3;
                dst[off + __popc(bm & lt)] = k[j];
            }
            o0 += __popc(b0); o1 += __popc(b1); o2 += __popc(b2); o3 += __popc(b3);
        }
    }
    if (tid == 0) *out = make_int4(t[0], t[1], t[2], t[3]);
    __syncthreads();
}

// child q (0..3 = n1..n4, :486-507) of parent p given the four child sizes
__device__ __forceinline__ QNode make_child(const QNode& p, const int4 c, int q) {
    const int mx = p.x0 + ((p.x1 - p.x0 + 1) >> 1), my = p.y0 + ((p.y1 - p.y0 + 1) >> 1);
    QNode n;
    n.x0 = (q & 1) ? (short)mx : p.x0;
    n.x1 = (q & 1) ? p.x1 : (short)mx;
    n.y0 = (q & 2) ? (short)my : p.y0;
    n.y1 = (q & 2) ? p.y1 : (short)my;
    const int cnt = q == 0 ? c.x : q == 1 ? c.y : q == 2 ? c.z : c.w;
    n.start = p.start + (q > 0 ? c.x : 0) + (q > 1 ? c.y : 0) + (q > 2 ? c.z : 0);
    n.cntbuf = cnt | ((node_buf(p) ^ 1) << 31);
    return n;
}

// Emit the children of one split node into the next node array.  `firstCreate` = creation index of its first
// non-empty child; creation index j lives at list position T-1-j (children are push_front'ed, :639-675);
// children with >1 keys are appended to the pending list (vSizeAndPointerToNode) in creation order.
__device__ __forceinline__ void emit_children(const QNode& p, const int4 c, int T, int firstCreate, int firstPend,
                                              QNode* next, int* pendNext) {
    int j = firstCreate, e = firstPend;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int cnt = q == 0 ? c.x : q == 1 ? c.y : q == 2 ? c.z : c.w;
        if (cnt > 0) {
            next[T - 1 - j] = make_child(p, c, q);
            if (cnt > 1) pendNext[e++] = T - 1 - j;
            j++;
        }
    }
}

__device__ __forceinline__ int nonempty4(const int4 c) { return (c.x > 0) + (c.y > 0) + (c.z > 0) + (c.w > 0); }
__device__ __forceinline__ int multi4(const int4 c) { return (c.x > 1) + (c.y > 1) + (c.z > 1) + (c.w > 1); }

// std::sort(a, a + n, compareNodes) by the whole CTA: the parallel formulation of introsort.cuh (std_sort_emul_pf is its
// host restatement, checked against the real std::sort).  Warp 0 runs the partition steps -- the positions l_k / r_k that
// the sequential two-pointer loop would stop at are listed with ballots, the K swaps happen at once -- and collects the
// leaf ranges; then one thread per leaf range does its insertion sort.  Lp, Rp: int[n]; leaf: unsigned[n]; stk: int[192].
__device__ void sort_emul_cta(u64* a, int n, int* Lp, int* Rp, unsigned* leaf, int* stk, int* sNLeaf, int tid, int nthreads) {
    if (n <= 1) return;                                     // (uniform)
    const int lane = tid & 31;
    if (tid < 32) {
        const int INF = 0x7fffffff;
        const unsigned lt = (1u << lane) - 1;
        int lg = 0;
        for (int t = n; t > 1; t >>= 1) lg++;
        int sp = 0, nLeaf = 0, first = 0, last = n, depth = 2 * lg;
        while (true) {
            bool heap = false;
            while (last - first > 16) {
                if (depth == 0) {                           // depth limit: heapsort of the range, serial (adversarial inputs only)
                    if (lane == 0) heap_sort_range(a, first, last);
                    __syncwarp();
                    heap = true;
                    break;
                }
                --depth;
                if (lane == 0) {                            // std::__move_median_to_first(first, first+1, mid, last-1)
                    const int x = first + 1, y = first + (last - first) / 2, z = last - 1;
                    if (rec_less(a[x], a[y])) {
                        if (rec_less(a[y], a[z])) rec_swap(a, first, y);
                        else if (rec_less(a[x], a[z])) rec_swap(a, first, z);
                        else rec_swap(a, first, x);
                    } else if (rec_less(a[x], a[z])) rec_swap(a, first, x);
                    else if (rec_less(a[y], a[z])) rec_swap(a, first, z);
                    else rec_swap(a, first, y);
                }
                __syncwarp();
                const u64 pk = a[first] >> ORBB_SORT_PAYLOAD_BITS;
                int nl = 0, nr = 0;
                for (int base = first + 1; base < last; base += 32) {
                    const int p = base + lane;
                    const bool f = p < last && !((a[min(p, last - 1)] >> ORBB_SORT_PAYLOAD_BITS) < pk);
                    const unsigned bal = __ballot_sync(0xffffffffu, f);
                    if (f) Lp[nl + __popc(bal & lt)] = p;
                    nl += __popc(bal);
                }
                for (int base = last - 1; base >= first; base -= 32) {
                    const int p = base - lane;
                    const bool f = p >= first && !(pk < (a[max(p, first)] >> ORBB_SORT_PAYLOAD_BITS));
                    const unsigned bal = __ballot_sync(0xffffffffu, f);
                    if (f) Rp[nr + __popc(bal & lt)] = p;
                    nr += __popc(bal);
                }
                __syncwarp();
                const int m = min(nl, nr);
                int K = 0;
                for (int base = 0; base < m; base += 32) {
                    const int k = base + lane;
                    K += __popc(__ballot_sync(0xffffffffu, k < m && Lp[k] < Rp[k]));
                }
                for (int k = lane; k < K; k += 32) rec_swap(a, Lp[k], Rp[k]);
                const int cut = min(K < nl ? Lp[K] : INF, K > 0 ? Rp[K - 1] : INF);
                __syncwarp();
                if (lane == 0) { stk[3 * sp] = cut; stk[3 * sp + 1] = last; stk[3 * sp + 2] = depth; }
                sp++;
                last = cut;
            }
            if (!heap) {
                if (lane == 0) leaf[nLeaf] = (unsigned)first | ((unsigned)last << 16);
                nLeaf++;
            }
            if (sp == 0) break;
            sp--;
            __syncwarp();
            first = stk[3 * sp]; last = stk[3 * sp + 1]; depth = stk[3 * sp + 2];
        }
        if (lane == 0) *sNLeaf = nLeaf;
    }
    __syncthreads();
    const int nLeaf = *sNLeaf;
    for (int s = tid; s < nLeaf; s += nthreads) insertion_sort(a, (int)(leaf[s] & 0xffffu), (int)(leaf[s] >> 16));
}

#ifdef ORBB_OT_TIMING
// debug build only (tools/ot_timing.py): clock64 at the phase boundaries of the level-0 quadtree of frame 0
__device__ long long g_otT[64];
__device__ int g_otN;
#define OT_TS(tag) do { if (tid == 0 && level == 0 && frame == 0 && g_otN < 31) { g_otT[2 * g_otN] = clock64(); g_otT[2 * g_otN + 1] = (tag); g_otN++; } } while (0)
#define OT_TS_RESET() do { if (tid == 0 && level == 0 && frame == 0) g_otN = 0; } while (0)
extern "C" int orbb_debug_ot_timing(long long* out) {
    int n = 0;
    cudaMemcpyFromSymbol(&n, g_otN, sizeof(int));
    cudaMemcpyFromSymbol(out, g_otT, sizeof(long long) * 64);
    return n;
}
#else
#define OT_TS(tag) do {} while (0)
#define OT_TS_RESET() do {} while (0)
#endif

// The same arrangement for n <= OT_SORT_SMEM records in shared memory, breadth first: the ranges left by a partition step are disjoint
// and may be partitioned in any order, so round r partitions ALL ranges of recursion depth r at once, one warp per range (the chain
// of dependent steps is the depth of the recursion, ~log2(n / 16), instead of the number of ranges, ~n / 10).  One ascending pass
// lists l_k and the r positions (ascending: r_k = Ra[nr - 1 - k]) in the range's own part of Lp / Ra; then one thread per RECORD
// ranks it inside its leaf range (a stable insertion sort of <= 16 records = position by (key, original index)).
// Lp, Ra: int[n]; leafOf: unsigned[n]; tmp: u64[n]; ranges: unsigned[2][OT_SORT_RANGES]; cnt: int[2].
constexpr int OT_SORT_RANGES = 64;          // ranges of > 16 records are disjoint: <= n / 17 of them
__device__ void sort_emul_bf(u64* a, int n, int* Lp, int* Ra, unsigned* leafOf, u64* tmp, unsigned (*ranges)[OT_SORT_RANGES], int* cnt, int tid,
                             int nthreads) {
    if (n <= 1) return;                                     // (uniform)
    const int lane = tid & 31, warp = tid >> 5, nw = nthreads >> 5;
    const unsigned lt = (1u << lane) - 1;
    const int INF = 0x7fffffff;
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) lg++;
    if (n <= 16) {
        if (tid < n) leafOf[tid] = (unsigned)n << 16;
        if (tid == 0) cnt[0] = 0;
    } else if (tid == 0) { ranges[0][0] = (unsigned)n << 16; cnt[0] = 1; }
    if (tid == 0) cnt[1] = 0;
    __syncthreads();
    int cur = 0;
    for (int depth = 2 * lg; ; depth--) {                   // (uniform: every range of a round has the same recursion depth)
        const int nr_ = cnt[cur];
        if (nr_ == 0) break;
        for (int ri = warp; ri < nr_; ri += nw) {
            const unsigned fl = ranges[cur][ri];
            const int first = (int)(fl & 0xffffu), last = (int)(fl >> 16);
            if (depth == 0) {                               // depth limit: heapsort of the range, serial (adversarial inputs only)
                if (lane == 0) heap_sort_range(a, first, last);
                for (int p = first + lane; p < last; p += 32) leafOf[p] = (unsigned)p | ((unsigned)(p + 1) << 16);
                continue;
            }
            {                                               // std::__move_median_to_first(first, first+1, mid, last-1): every lane decides, lane 0 swaps
                const int x = first + 1, y = first + (last - first) / 2, z = last - 1;
                const u64 ax = a[x], ay = a[y], az = a[z];
                int w;
                if (rec_less(ax, ay)) w = rec_less(ay, az) ? y : rec_less(ax, az) ? z : x;
                else w = rec_less(ax, az) ? x : rec_less(ay, az) ? z : y;
                __syncwarp();
                if (lane == 0) rec_swap(a, first, w);
                __syncwarp();
            }
            const u64 pk = a[first] >> ORBB_SORT_PAYLOAD_BITS;
            int nl = 0, nr = 0;
            for (int base = first; base < last; base += 32) {
                const int p = base + lane;
                const u64 key = p < last ? a[p] >> ORBB_SORT_PAYLOAD_BITS : 0;
                const bool fL = p > first && p < last && !(key < pk), fR = p < last && !(pk < key);
                const unsigned bL = __ballot_sync(0xffffffffu, fL), bR = __ballot_sync(0xffffffffu, fR);
                if (fL) Lp[first + nl + __popc(bL & lt)] = p;
                if (fR) Ra[first + nr + __popc(bR & lt)] = p;
                nl += __popc(bL);
                nr += __popc(bR);
            }
            __syncwarp();
            const int m = min(nl, nr);
            int K = 0;
            for (int base = 0; base < m; base += 32) {      // (l_k < r_k holds for a prefix of k)
                const int k = base + lane;
                const unsigned ok = __ballot_sync(0xffffffffu, k < m && Lp[first + min(k, m - 1)] < Ra[first + nr - 1 - min(k, m - 1)]);
                K += __popc(ok);
                if (ok != 0xffffffffu) break;
            }
            for (int k = lane; k < K; k += 32) rec_swap(a, Lp[first + k], Ra[first + nr - 1 - k]);
            const int cut = min(K < nl ? Lp[first + K] : INF, K > 0 ? Ra[first + nr - K] : INF);
            // children: [first, cut) and [cut, last); the ones that are still long go to the next round
            if (cut - first > 16) { if (lane == 0) ranges[cur ^ 1][atomicAdd(&cnt[cur ^ 1], 1)] = (unsigned)first | ((unsigned)cut << 16); }
            else if (lane < cut - first) leafOf[first + lane] = (unsigned)first | ((unsigned)cut << 16);
            if (last - cut > 16) { if (lane == 0) ranges[cur ^ 1][atomicAdd(&cnt[cur ^ 1], 1)] = (unsigned)cut | ((unsigned)last << 16); }
            else if (lane < last - cut) leafOf[cut + lane] = (unsigned)cut | ((unsigned)last << 16);
        }
        __syncthreads();
        if (tid == 0) cnt[cur] = 0;
        cur ^= 1;
        __syncthreads();
    }
    for (int i = tid; i < n; i += nthreads) {
        const unsigned fl = leafOf[i];
        const int f = (int)(fl & 0xffffu), l = (int)(fl >> 16);
        const u64 rec = a[i], key = rec >> ORBB_SORT_PAYLOAD_BITS;
        int pos = f;
        for (int j = f; j < l; j++) {
            const u64 kj = a[j] >> ORBB_SORT_PAYLOAD_BITS;
            pos += (kj < key) || (kj == key && j < i);
        }
        tmp[pos] = rec;
    }
    __syncthreads();
    for (int i = tid; i < n; i += nthreads) a[i] = tmp[i];
}

// LAT (a call with a few frames, one CTA per SM): keys and node arrays live in dynamic shared memory when the level fits, and the
// pending list is sorted breadth first -- every phase of the tree is a chain of dependent round trips, so their latency is the run time.
// Batches keep global arrays (plain LDG/STG instead of generic accesses, ~10 CTAs per SM) and the one-warp sort.
template <int OT_T, bool LAT>
__global__ void __launch_bounds__(OT_T) k_octree(const Plan* __restrict__ P, Bufs B, int level0, int bigNode) {
    constexpr int OT_W = OT_T / 32;
    // grid = (frames, levels): CTAs are dispatched x-fastest, so the long-running low levels of ALL frames start first and
    // the short high levels fill the tail
    const int level = level0 + blockIdx.y, frame = blockIdx.x;
    const LevelPlan& L = P->lv[level];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = L.nFeat;
    pdl_launch_dependents();
    pdl_wait();
    OT_TS_RESET();
    OT_TS(0);

    const int* cellCount = B.cellCount + (size_t)frame * P->cellsTotal + L.cellBase;
    int* cellOff = B.cellOff + (size_t)frame * P->cellsTotal + L.cellBase;
    const u64* cellKeys = B.cellKeys + (size_t)frame * P->cellKeyStride + L.cellKeyBase;
    u64* k0 = B.keys + ((size_t)frame * 2 + 0) * P->rawStride + L.rawBase;      // (not const: see sKeyBuf)
    u64* k1 = B.keys + ((size_t)frame * 2 + 1) * P->rawStride + L.rawBase;
    QNode* nodesAB[2] = {B.nodes + ((size_t)frame * 2 + 0) * P->nodeStride + L.nodeBase,
                         B.nodes + ((size_t)frame * 2 + 1) * P->nodeStride + L.nodeBase};
    int* pendAB[2] = {B.pend + ((size_t)frame * 2 + 0) * P->nodeStride + L.nodeBase,
                      B.pend + ((size_t)frame * 2 + 1) * P->nodeStride + L.nodeBase};
    u64* rec = B.rec + (size_t)frame * P->nodeStride + L.nodeBase;
    int4* cnt4 = B.cnt4 + (size_t)frame * P->nodeStride + L.nodeBase;
    int* elist = B.elist + (size_t)frame * P->nodeStride + L.nodeBase;
    uint8_t* erased = B.erased + (size_t)frame * P->nodeStride + L.nodeBase;

    __shared__ int sN, sNodes, sPend, sM, sFinish, sPhase2, sCur, sPcur;
    __shared__ int sSlotCnt[kMaxIni], sSlotStart[kMaxIni];
    __shared__ int sWarpSlot[OT_W][kMaxIni];
    __shared__ u64 sRec[OT_SORT_SMEM];
    __shared__ int sSplitPart[OT_W][4];
    __shared__ int sBigList[OT_BIG_LIST], sNBig;
    __shared__ int sSortL[OT_SORT_SMEM], sSortR[OT_SORT_SMEM], sSortStk[192], sSortLeaves, sSortCnt[2];
    __shared__ unsigned sSortLeaf[OT_SORT_SMEM], sSortRanges[2][OT_SORT_RANGES];
    __shared__ u64 sSortTmp[LAT ? OT_SORT_SMEM : 1];

    // ---- 1. gather vToDistributeKeys in the reference's order: cell-row-major, raster inside the cell ----
    const int nCells = L.nCols * L.nRows;
    {   // exclusive prefix of the cell counts by the whole CTA: every thread owns a contiguous run of cells (independent
        // loads), the runs' sums are scanned through shared memory
        const int chunk = (nCells + OT_T - 1) / OT_T, c0 = min(tid * chunk, nCells), c1 = min(c0 + chunk, nCells);
        int sum = 0;
        for (int c = c0; c < c1; c++) sum += __ldg(cellCount + c);
        const int inc = warp_incl_scan(sum, lane);
        if (lane == 31) sWarpSlot[warp][0] = inc;
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < OT_W; w++) { const int v = sWarpSlot[w][0]; total += v; if (w < warp) before += v; }
        int run = before + inc - sum;
        for (int c = c0; c < c1; c++) { cellOff[c] = run; run += __ldg(cellCount + c); }
        if (tid == 0) sN = total;
        __syncthreads();                                    // sWarpSlot is reused below
    }
    if (tid < kMaxIni) sSlotCnt[tid] = 0;
    __syncthreads();
    OT_TS(1);
    const int n = sN;
    if (LAT) {
        extern __shared__ __align__(16) unsigned char sDyn[];
        if (n <= OT_SMEM_KEYS) { k0 = reinterpret_cast<u64*>(sDyn); k1 = k0 + OT_SMEM_KEYS; }
        if (L.maxNodes <= OT_SMEM_NODES) {
            unsigned char* q = sDyn + (size_t)OT_SMEM_KEYS * 16;
            nodesAB[0] = reinterpret_cast<QNode*>(q); nodesAB[1] = nodesAB[0] + OT_SMEM_NODES; q += 2 * OT_SMEM_NODES * sizeof(QNode);
            cnt4 = reinterpret_cast<int4*>(q); q += OT_SMEM_NODES * sizeof(int4);
            pendAB[0] = reinterpret_cast<int*>(q); pendAB[1] = pendAB[0] + OT_SMEM_NODES; q += 2 * OT_SMEM_NODES * sizeof(int);
            elist = reinterpret_cast<int*>(q); q += OT_SMEM_NODES * sizeof(int);
            erased = q;
        }
    }
    // one thread per cell: the loads of a cell's keys are independent (read-only path), so their latencies overlap instead
    // of adding up cell after cell
    for (int c = tid; c < nCells; c += OT_T) {
        const int cnt = __ldg(cellCount + c), off = cellOff[c];
        const u64* src = cellKeys + (size_t)c * L.cellCap;
#pragma unroll 4
        for (int i = 0; i < cnt; i++) k0[off + i] = __ldg(src + i);
    }
    __syncthreads();
    OT_TS(2);

    // ---- 2. root nodes (:559-601): keys bucketed by (int)(x / hX), stable ----
    const int nIni = L.nIni;
    const float hX = L.hX;
    int rootBuf = 0;
    if (nIni > 1) {
        rootBuf = 1;
        // each warp owns a contiguous quarter of the keys: count per root, then write in order behind the warps before it
        const int seg = (n + OT_W - 1) / OT_W, s0 = min(warp * seg, n), s1 = min(s0 + seg, n);
        int cntS[kMaxIni];
#pragma unroll
        for (int s = 0; s < kMaxIni; s++) cntS[s] = 0;
        for (int base = s0; base < s1; base += 32) {
            const int i = base + lane;
            const int slot = i < s1 ? min(__float2int_rz(__fdiv_rn((float)(int)(k0[i] & 0xffff), hX)), nIni - 1) : -1;
#pragma unroll
            for (int s = 0; s < kMaxIni; s++)
                if (s < nIni) cntS[s] += __popc(__ballot_sync(0xffffffffu, slot == s));
        }
#pragma unroll
        for (int s = 0; s < kMaxIni; s++)
            if (s < nIni && lane == 0) sWarpSlot[warp][s] = cntS[s];
        __syncthreads();
        int dstS[kMaxIni];                                   // first output position of this warp's keys of root s
        {
            int acc = 0;
#pragma unroll
            for (int s = 0; s < kMaxIni; s++) {
                dstS[s] = 0;
                if (s < nIni) {
                    int tot = 0, before = 0;
                    for (int w = 0; w < OT_W; w++) { const int v = sWarpSlot[w][s]; tot += v; if (w < warp) before += v; }
                    dstS[s] = acc + before;
                    if (tid == 0) { sSlotStart[s] = acc; sSlotCnt[s] = tot; }
                    acc += tot;
                }
            }
        }
        for (int base = s0; base < s1; base += 32) {
            const int i = base + lane;
            const bool valid = i < s1;
            const u64 k = valid ? k0[i] : 0;
            const int slot = valid ? min(__float2int_rz(__fdiv_rn((float)(int)(k & 0xffff), hX)), nIni - 1) : -1;
#pragma unroll
            for (int s = 0; s < kMaxIni; s++) {
                if (s < nIni) {
                    const unsigned bal = __ballot_sync(0xffffffffu, slot == s);
                    if (slot == s) k1[dstS[s] + __popc(bal & ((1u << lane) - 1))] = k;
                    dstS[s] += __popc(bal);
                }
            }
        }
    } else if (tid == 0) {
        sSlotCnt[0] = n;
        sSlotStart[0] = 0;
    }
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int s = 0; s < nIni; s++) {
            if (sSlotCnt[s] == 0) continue;                               // :597-598 empty roots are erased
            QNode r;
            r.x0 = (short)__float2int_rz(__fmul_rn(hX, (float)s));        // :571
            r.x1 = (short)__float2int_rz(__fmul_rn(hX, (float)(s + 1)));  // :572
            r.y0 = 0;
            r.y1 = (short)(L.maxBY - kMinBorder);                         // :573
            r.start = sSlotStart[s];
            r.cntbuf = sSlotCnt[s] | (rootBuf << 31);
            nodesAB[0][m++] = r;
        }
        sNodes = m; sCur = 0; sPcur = 0; sPend = 0; sFinish = 0; sPhase2 = 0;
    }
    __syncthreads();
    OT_TS(3);

    // ---- 3. main loop (:610-755) ----
    while (true) {
        if (sFinish) break;
        const int cur = sCur, pcur = sPcur;
        QNode* nodes = nodesAB[cur];
        QNode* next = nodesAB[cur ^ 1];
        int* pendCur = pendAB[pcur];
        int* pendNext = pendAB[pcur ^ 1];
        const int nNodes = sNodes;
        const bool phase2 = sPhase2 != 0;
        __syncthreads();           // everyone has read the shared state before warp 0 rewrites it

        if (!phase2) {
            // ---------- phase 1: split every node that holds more than one key, in list order (:622-681) ----------
            if (warp == 0) {
                int run = 0;
                int nBig = 0;
                for (int base = 0; base < nNodes; base += 32) {
                    const int i = base + lane;
                    const int cnt = i < nNodes ? node_count(nodes[i]) : 0;
                    const bool ex = cnt > 1;
                    const unsigned bal = __ballot_sync(0xffffffffu, ex);
                    if (ex) elist[run + __popc(bal & ((1u << lane) - 1))] = i;
                    if (OT_T > 128) {                               // large levels: nodes with many keys are split by the whole CTA
                        const unsigned big = __ballot_sync(0xffffffffu, cnt >= bigNode);
                        const int slot = nBig + __popc(big & ((1u << lane) - 1));
                        if (cnt >= bigNode && slot < OT_BIG_LIST) sBigList[slot] = run + __popc(bal & ((1u << lane) - 1));
                        nBig += __popc(big);
                    }
                    run += __popc(bal);
                }
                if (lane == 0) { sM = run; sNBig = min(nBig, OT_BIG_LIST); }
            }
            __syncthreads();
            const int m = sM;
            OT_TS(10);
            if (m == 0) break;                                      // size == prevSize (:685)
            const int nBig = OT_T > 128 ? sNBig : 0;
            for (int b = 0; b < nBig; b++) {                        // (uniform)
                const int e = sBigList[b];
                split_node_cta(nodes[elist[e]], k0, k1, &cnt4[elist[e]], sSplitPart, tid, OT_W);
            }
            for (int e = warp; e < m; e += OT_W) {
                bool done = false;
                for (int b = 0; b < nBig; b++) done |= sBigList[b] == e;
                if (!done) split_node_warp(nodes[elist[e]], k0, k1, &cnt4[elist[e]], lane);
            }
            __syncthreads();
            OT_TS(11);
            if (warp == 0) {
                int T = 0, X = 0;
                for (int base = 0; base < m; base += 32) {
                    const int e = base + lane;
                    int4 c = make_int4(0, 0, 0, 0);
                    if (e < m) c = cnt4[elist[e]];
                    T += nonempty4(c);
                    X += multi4(c);
                }
                T = warp_sum(T);
                X = warp_sum(X);
                int runC = 0, runE = 0;
                for (int base = 0; base < m; base += 32) {
                    const int e = base + lane;
                    int4 c = make_int4(0, 0, 0, 0);
                    int idx = 0;
                    if (e < m) { idx = elist[e]; c = cnt4[idx]; }
                    const int nc = nonempty4(c), ne = multi4(c);
                    const int incC = warp_incl_scan(nc, lane), incE = warp_incl_scan(ne, lane);
                    if (e < m) emit_children(nodes[idx], c, T, runC + incC - nc, runE + incE - ne, next, pendNext);
                    runC += __shfl_sync(0xffffffffu, incC, 31);
                    runE += __shfl_sync(0xffffffffu, incE, 31);
                }
                int runK = 0;                                       // single-key nodes keep their relative order
                for (int base = 0; base < nNodes; base += 32) {
                    const int i = base + lane;
                    QNode nd;
                    bool keep = false;
                    if (i < nNodes) { nd = nodes[i]; keep = node_count(nd) == 1; }
                    const unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) next[T + runK + __popc(bal & ((1u << lane) - 1))] = nd;
                    runK += __popc(bal);
                }
                if (lane == 0) {
                    const int newSize = T + (nNodes - m);
                    sNodes = newSize; sPend = X; sCur = cur ^ 1; sPcur = pcur ^ 1;
                    if (newSize >= N || newSize == nNodes) sFinish = 1;             // :685
                    else if (newSize + X * 3 > N) sPhase2 = 1;                      // :689
                }
            }
            __syncthreads();
            OT_TS(12);
        } else {
            // ---------- phase 2: sort the pending nodes, split from the back until N nodes exist (:692-753) ----------
            const int len = sPend;
            u64* srt = len <= OT_SORT_SMEM ? sRec : rec;
            for (int i = tid; i < nNodes; i += OT_T) erased[i] = 0;
            for (int e = warp; e < len; e += OT_W) {
                const int idx = pendCur[e];
                const QNode nd = nodes[idx];
                split_node_warp(nd, k0, k1, &cnt4[idx], lane);
                if (lane == 0)
                    srt[e] = ((u64)(unsigned)node_count(nd) << 40) | ((u64)(unsigned short)nd.x0 << 24) | (u64)(unsigned)idx;
            }
            __syncthreads();
            OT_TS(20);
#ifdef ORBB_OT_TIMING
            if (tid == 0 && level == 0 && frame == 0) { g_otT[61] = len; g_otT[62] = nNodes; g_otT[63] = n; }
#endif
            // std::sort(..., compareNodes) :700
            if (LAT && len <= OT_SORT_SMEM) sort_emul_bf(srt, len, sSortL, sSortR, sSortLeaf, sSortTmp, sSortRanges, sSortCnt, tid, OT_T);
            else if (len <= OT_SORT_SMEM) sort_emul_cta(srt, len, sSortL, sSortR, sSortLeaf, sSortStk, &sSortLeaves, tid, OT_T);
            else {
                int* tmp = B.sortTmp + (size_t)frame * 3 * P->nodeStride + 3 * (size_t)L.nodeBase;
                sort_emul_cta(srt, len, tmp, tmp + L.maxNodes, reinterpret_cast<unsigned*>(tmp + 2 * L.maxNodes), sSortStk, &sSortLeaves, tid, OT_T);
            }
            __syncthreads();
            OT_TS(21);
            if (warp == 0) {
                // processing order i = 0..len-1 is the sorted vector walked from the back (:701)
                int kstar = len, run = 0;
                for (int base = 0; base < len && kstar == len; base += 32) {
                    const int i = base + lane;
                    int g = 0;
                    if (i < len) g = nonempty4(cnt4[(int)(srt[len - 1 - i] & 0xffffff)]) - 1;
                    const int inc = warp_incl_scan(g, lane);
                    const unsigned bal = __ballot_sync(0xffffffffu, i < len && nNodes + run + inc >= N);   // :746
                    if (bal) kstar = base + __ffs(bal);             // number of nodes split before the break
                    run += __shfl_sync(0xffffffffu, inc, 31);
                }
                int T = 0, X = 0;
                for (int base = 0; base < kstar; base += 32) {
                    const int i = base + lane;
                    int4 c = make_int4(0, 0, 0, 0);
                    if (i < kstar) c = cnt4[(int)(srt[len - 1 - i] & 0xffffff)];
                    T += nonempty4(c);
                    X += multi4(c);
                }
                T = warp_sum(T);
                X = warp_sum(X);
                int runC = 0, runE = 0;
                for (int base = 0; base < kstar; base += 32) {
                    const int i = base + lane;
                    int4 c = make_int4(0, 0, 0, 0);
                    int idx = 0;
                    if (i < kstar) { idx = (int)(srt[len - 1 - i] & 0xffffff); c = cnt4[idx]; erased[idx] = 1; }
                    const int nc = nonempty4(c), ne = multi4(c);
                    const int incC = warp_incl_scan(nc, lane), incE = warp_incl_scan(ne, lane);
                    if (i < kstar) emit_children(nodes[idx], c, T, runC + incC - nc, runE + incE - ne, next, pendNext);
                    runC += __shfl_sync(0xffffffffu, incC, 31);
                    runE += __shfl_sync(0xffffffffu, incE, 31);
                }
                __syncwarp();
                int runK = 0;
                for (int base = 0; base < nNodes; base += 32) {
                    const int i = base + lane;
                    QNode nd;
                    bool keep = false;
                    if (i < nNodes) { nd = nodes[i]; keep = !erased[i]; }
                    const unsigned bal = __ballot_sync(0xffffffffu, keep);
                    if (keep) next[T + runK + __popc(bal & ((1u << lane) - 1))] = nd;
                    runK += __popc(bal);
                }
                if (lane == 0) {
                    const int newSize = T + (nNodes - kstar);
                    sNodes = newSize; sPend = X; sCur = cur ^ 1; sPcur = pcur ^ 1;
                    if (newSize >= N || newSize == nNodes) sFinish = 1;             // :750
                }
            }
            __syncthreads();
            OT_TS(22);
        }
    }
    __syncthreads();

    // ---- 4. best key of every node, first maximum wins (:757-776); +16 border offset (:886-887) ----
    const QNode* fin = nodesAB[sCur];
    const int nOut = sNodes;
    u64* sel = B.sel + (size_t)frame * P->selStride + L.selBase;
    for (int i = tid; i < nOut && i < L.selCap; i += OT_T) {
        const QNode nd = fin[i];
        const u64* src = (node_buf(nd) ? k1 : k0) + nd.start;
        u64 best = src[0];
        const int cnt = node_count(nd);
        for (int k = 1; k < cnt; k++) {
            const u64 v = src[k];
            if ((unsigned)(v >> 32) > (unsigned)(best >> 32)) best = v;
        }
        sel[i] = best + (u64)kMinBorder + ((u64)kMinBorder << 16);
    }
#ifdef ORBB_OT_TIMING
    __syncthreads();
#endif
    OT_TS(4);
    if (tid == 0) {
        B.selCount[frame * ORBB_MAX_LEVELS + level] = min(nOut, L.selCap);
        if (nOut > L.selCap) atomicOr(&B.status[frame], 1);
    }
}

// ------------------------------------------------------------------------------------------------
// K6: output assembly (:1105-1167): level-major order, pt *= scale for level > 0, keypoints inside the lapping
// area are written from the back, the others from the front.  One CTA per frame.  (Batches; a call with a few frames does this in the
// prologue of k_orient_desc32<.., true> instead.)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_assemble(const Plan* __restrict__ P, Bufs B, int lap0, int lap1) {
    const int frame = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int sOff[ORBB_MAX_LEVELS + 1];
    __shared__ int sWarp[32];
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    if (warp == 0) {                                        // exclusive prefix of the per-level counts: one load per lane, one warp scan
        const int nl = P->nlevels;                          // (ORBB_MAX_LEVELS <= 32)
        const int c = lane < nl ? B.selCount[frame * ORBB_MAX_LEVELS + lane] : 0;
        const int inc = warp_incl_scan(c, lane);
        if (lane < nl) sOff[lane] = inc - c;
        if (lane == nl - 1) sOff[nl] = inc;
    }
    __syncthreads();
    const int n = min(sOff[P->nlevels], P->kpCap);
    orbb_keypoint* kps = B.kps + (size_t)frame * P->kpCap;
    WorkItem* work = B.work + (size_t)frame * P->kpCap;
    const float fl0 = (float)lap0, fl1 = (float)lap1;
    int stereoRun = 0;
    for (int base = 0; base < n; base += nthreads) {
        const int g = base + tid;
        bool valid = g < n, st = false;
        int level = 0, x = 0, y = 0;
        float xs = 0, ys = 0, resp = 0;
        if (valid) {
            while (g >= sOff[level + 1]) level++;
            const LevelPlan& L = P->lv[level];
            const u64 k = B.sel[(size_t)frame * P->selStride + L.selBase + (g - sOff[level])];
            x = (int)(k & 0xffff); y = (int)((k >> 16) & 0xffff); resp = (float)(int)(k >> 32);
            xs = (float)x; ys = (float)y;
            if (level != 0) { xs = __fmul_rn(xs, L.scale); ys = __fmul_rn(ys, L.scale); }     // :1149-1151
            st = xs >= fl0 && xs <= fl1;                                                       // :1153
        }
        const unsigned bal = __ballot_sync(0xffffffffu, st);
        if (lane == 0) sWarp[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < nwarps; w++) { woff += w < warp ? sWarp[w] : 0; tot += sWarp[w]; }
        if (valid) {
            const int stBefore = stereoRun + woff + __popc(bal & ((1u << lane) - 1));
            const int pos = st ? n - 1 - stBefore : g - stBefore;
            orbb_keypoint kp;
            kp.x = xs; kp.y = ys; kp.size = P->lv[level].kpSize; kp.angle = -1.f; kp.response = resp; kp.octave = level;
            kps[pos] = kp;
            work[g] = WorkItem{level, x, y, pos};
        }
        stereoRun += tot;
        __syncthreads();
    }
    if (tid == 0) {
        B.outCount[frame * 2] = n;
        B.outCount[frame * 2 + 1] = n - stereoRun;
        if (sOff[P->nlevels] > P->kpCap) atomicOr(&B.status[frame], 2);
    }
}

// ------------------------------------------------------------------------------------------------
// K4 + K5b: one warp per keypoint -- IC_Angle over the umax disc (:76-103), cv::fastAtan2 (un-fused float32),
// then the 256 steered rBRIEF tests on the blurred level (:107-146).
// ------------------------------------------------------------------------------------------------
__constant__ signed char cPattern[256][4] = {
#include "orb_pattern.inc"
};
// the same table as floats, transposed so that lane i (descriptor byte i) finds its test j at [j*32 + i]; filled once
// per process by k_init_pattern (identical for every handle)
__device__ float4 gPatF[256];
__global__ void k_init_pattern() {
    const int i = threadIdx.x >> 3, j = threadIdx.x & 7;
    const signed char* p = cPattern[8 * i + j];
    gPatF[j * 32 + i] = make_float4((float)p[0], (float)p[1], (float)p[2], (float)p[3]);
}

// cvRound(float) for |v| < 2^22 without the F2I (XU pipe) instruction: adding 1.5*2^23 makes the FADD itself round
// to nearest-even at integer granularity; the integer sits in the low mantissa bits.
__device__ __forceinline__ int round_rne(float v) { return __float_as_int(__fadd_rn(v, 12582912.f)) - 0x4B400000; }

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    // cv::fastAtan2 scalar path; p-coefficients are the float products OpenCV stores
    const float sc = (float)(180.0 / 3.141592653589793238462643383279502884197);
    const float p1 = 0.9997878412794807f * sc, p3 = -0.3258083974640975f * sc;
    const float p5 = 0.1555786518463281f * sc, p7 = -0.04432655554792128f * sc;
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// One warp takes OD_KPW keypoints of a frame through three phases, so that the per-keypoint scalar
// work (fastAtan2, the double-precision cos / sin) runs once per LANE instead of once per warp:
//   1. moments: the 31 x 31 window around the keypoint is read as 31 rows x 9 aligned words, consecutive lanes taking
//      consecutive words (coalesced); each word is dotted (IDP.4A) with two weight words from a table indexed by the
//      keypoint column's alignment -- u and v of the four pixels as int8, zero outside the umax circle -- which gives
//      m10 = sum(u*I) and m01 = sum(v*I) directly; warp-sum; lane k keeps keypoint k's moments.
//   2. lanes 0..OD_KPW-1: fastAtan2 + cos/sin of their own keypoint.
//   3. per keypoint, lane = descriptor byte: the 8 steered tests, pattern points held in registers.
constexpr int OD_THREADS = 256;
#ifndef ORBB_OD_KPW
#define ORBB_OD_KPW 8
#endif
constexpr int OD_KPW = ORBB_OD_KPW;
constexpr int OD_ITEMS = 31 * 9;                              // words of the moment window

__device__ int2 gMomW[4][9][32];                             // [column alignment][round][lane] -> (u weights, v weights)
__global__ void k_init_moment_weights() {
    constexpr int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    for (int i = threadIdx.x; i < 4 * 9 * 32; i += blockDim.x) {
        const int sh = i / (9 * 32), t = i % (9 * 32);       // t = round * 32 + lane = row * 9 + word
        unsigned wu = 0, wv = 0;
        if (t < OD_ITEMS) {
            const int row = t / 9, wq = t % 9, v = row - 15;
            for (int j = 0; j < 4; j++) {
                const int u = 4 * wq + j - sh - 15;          // word 0 starts at column (x - 15) & ~3
                if (u >= -15 && u <= 15 && (u < 0 ? -u : u) <= kUmax[v < 0 ? -v : v]) {
                    wu |= (unsigned)(u & 0xff) << (8 * j);
                    wv |= (unsigned)(v & 0xff) << (8 * j);
                }
            }
        }
        (&gMomW[0][0][0])[i] = make_int2((int)wu, (int)wv);
    }
}

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {         // unsigned bytes x signed bytes
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// STAGE: the 37 x 37 window of the blurred level that the rotated pattern can reach (|coordinate| <= 13 -> radius < 18.4) is first
// copied into shared memory -- 37 rows x four 16-byte chunks from the 16-byte aligned column left of the window, lane t of round q moves
// chunk 32q + t with one cp.async (5 rounds), the copy of keypoint k+1 running while keypoint k is sampled -- and the 512 samples
// are LDS.U8 with a few bank conflicts instead of global gathers that cost ~12 L1 data-pipe wavefronts each (the L1 data pipe, at
// 83 % of its peak, bounded the un-staged kernel: 0.251 -> 0.217 ms per 256 frames with 4-byte copies, 12 rounds).
constexpr int OD_PROWS = 37, OD_PPITCH = 64, OD_PCHUNKS = OD_PROWS * (OD_PPITCH / 16), OD_PROUNDS = (OD_PCHUNKS + 31) / 32;
constexpr int OD_PATCH_WORDS = OD_PROWS * OD_PPITCH / 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}

// The output assembly (K6, :1105-1167) is the kernel's prologue: a CTA derives the level, the level coordinates and the output position
// of its 64 key points itself -- level offsets from the quadtree's per-level counts, and for the lapping split the number of key points
// inside [lap0, lap1] BEFORE its first one, counted by the CTA over the selected keys (<= 1000 8-byte reads from L2).  That removes a
// kernel of one CTA per frame from the critical path (15 us per 256 frames, 11 of the 135 us of the single-frame call).
// In that form (calls with a few frames) a warp takes OD_KPW_LAT key points instead of OD_KPW: 1000 key points are 16 CTAs at 8 per
// warp -- a tenth of the machine, every warp walking through 8 moments and 8 descriptors one after the other -- and 126 CTAs at 1
// (measured per frame: 0.109 ms at 8, 0.106 at 2, 0.100 at 1).
#ifndef ORBB_OD_KPW_LAT
#define ORBB_OD_KPW_LAT 1
#endif
constexpr int OD_KPW_LAT = ORBB_OD_KPW_LAT;

template <bool STAGE, bool ASM>
__global__ void __launch_bounds__(OD_THREADS, 4) k_orient_desc32(const Plan* __restrict__ P, Bufs B, int lap0, int lap1) {
    __shared__ __align__(16) unsigned sPatch[STAGE ? OD_THREADS / 32 : 1][2][STAGE ? OD_PATCH_WORDS : 4];
    __shared__ int sOff[ASM ? ORBB_MAX_LEVELS + 1 : 1];
    __shared__ int sRed[ASM ? OD_THREADS / 32 + 2 : 1];
    constexpr int KPW = ASM ? OD_KPW_LAT : OD_KPW;           // key points per warp
    constexpr int CTA_KPS = (OD_THREADS / 32) * KPW;          // key points per CTA
    static_assert(!ASM || CTA_KPS <= 64, "the assembly prologue holds one key point per thread of warps 0 and 1");
    __shared__ WorkItem sWork[ASM ? CTA_KPS : 1];
    const int frame = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();
    int n, g0;
    const WorkItem* work;
    if constexpr (!ASM) {                                     // batches: k_assemble has run
        n = B.outCount[frame * 2];
        g0 = (blockIdx.x * (OD_THREADS / 32) + warp) * KPW;
        work = B.work + (size_t)frame * P->kpCap + g0;
    } else {
    // ---- 0. output assembly: level-major order, pt *= scale for level > 0, key points inside the lapping area from the back ----
    const int nl = P->nlevels;
    if (warp == 0) {                                          // exclusive prefix of the per-level counts (ORBB_MAX_LEVELS <= 32)
        const int c = lane < nl ? B.selCount[frame * ORBB_MAX_LEVELS + lane] : 0;
        const int inc = warp_incl_scan(c, lane);
        if (lane < nl) sOff[lane] = inc - c;
        if (lane == nl - 1) sOff[nl] = inc;
    }
    __syncthreads();
    n = min(sOff[nl], P->kpCap);
    const int G0 = blockIdx.x * CTA_KPS;
    if (blockIdx.x != 0 && G0 >= n) return;                   // (uniform; CTA 0 always runs: it writes the frame's counts)
    const float fl0 = (float)lap0, fl1 = (float)lap1;
    const u64* selF = B.sel + (size_t)frame * P->selStride;
    auto key_of = [&](int g, int& level) {                    // selected key g of the frame in level-major order
        level = 0;
        while (g >= sOff[level + 1]) level++;
        return selF[P->lv[level].selBase + (g - sOff[level])];
    };
    auto lapping = [&](u64 k, int level) {                    // :1149-1153
        float xs = (float)(int)(k & 0xffff);
        if (level != 0) xs = __fmul_rn(xs, P->lv[level].scale);
        return xs >= fl0 && xs <= fl1;
    };
    // key points inside the lapping area before this CTA's first one (CTA 0 also counts the whole frame: monoIndex)
    const int countTo = blockIdx.x == 0 ? n : min(G0, n);
    int before = 0, total = 0;
    for (int g = tid; g < countTo; g += OD_THREADS) {
        int level;
        const u64 k = key_of(g, level);
        const int st = lapping(k, level);
        total += st;
        before += g < G0 ? st : 0;
    }
    before = warp_sum(before);
    total = warp_sum(total);
    if (lane == 0) sRed[warp] = blockIdx.x == 0 ? total : before;
    __syncthreads();
    int stBefore = 0;
    for (int w = 0; w < OD_THREADS / 32; w++) stBefore += sRed[w];
    if (blockIdx.x == 0) {
        if (tid == 0) {
            B.outCount[frame * 2] = n;
            B.outCount[frame * 2 + 1] = n - stBefore;         // monoIndex: key points written at the front
            if (sOff[nl] > P->kpCap) atomicOr(&B.status[frame], 2);
        }
        stBefore = 0;                                         // (CTA 0 starts at key point 0)
    }
    if (G0 >= n) return;                                      // (uniform)
    __syncthreads();                                          // sRed is reused below
    if (tid < (CTA_KPS > 32 ? 64 : 32)) {                     // warp 0 (and warp 1 when a CTA holds more than 32): one key point per thread
        const int g = G0 + tid;
        const bool mine = tid < CTA_KPS && g < n;
        bool st = false;
        int level = 0, x = 0, y = 0;
        float xs = 0, ys = 0, resp = 0;
        if (mine) {
            const u64 k = key_of(g, level);
            x = (int)(k & 0xffff); y = (int)((k >> 16) & 0xffff); resp = (float)(int)(k >> 32);
            xs = (float)x; ys = (float)y;
            if (level != 0) { xs = __fmul_rn(xs, P->lv[level].scale); ys = __fmul_rn(ys, P->lv[level].scale); }      // :1149-1151
            st = xs >= fl0 && xs <= fl1;                                                                              // :1153
        }
        const unsigned bal = __ballot_sync(0xffffffffu, st);
        int stHere = stBefore + __popc(bal & ((1u << lane) - 1));
        if (CTA_KPS > 32) {
            if (lane == 0) sRed[OD_THREADS / 32 + warp] = __popc(bal);
            asm volatile("bar.sync 1, 64;" ::: "memory");     // the two warps that hold key points
            if (warp == 1) stHere += sRed[OD_THREADS / 32];
        }
        if (mine) {
            const int pos = st ? n - 1 - stHere : g - stHere;
            orbb_keypoint kp;
            kp.x = xs; kp.y = ys; kp.size = P->lv[level].kpSize; kp.angle = -1.f; kp.response = resp; kp.octave = level;
            B.kps[(size_t)frame * P->kpCap + pos] = kp;
            sWork[tid] = WorkItem{level, x, y, pos};
        }
    }
    __syncthreads();
    g0 = G0 + warp * KPW;
    work = sWork + warp * KPW;
    }
    if (g0 >= n) return;
    const int cnt = min(KPW, n - g0);
    const uint8_t* pyr = B.pyr + (size_t)frame * P->pyrStride;
    // ---- 1. IC_Angle moments (:76-103) ----
    int M10 = 0, M01 = 0;
    int rowOff[9], colOff[9];
#pragma unroll
    for (int q = 0; q < 9; q++) {
        const int t = min(q * 32 + lane, OD_ITEMS - 1);      // (lanes past the window re-read its last word with zero weights)
        rowOff[q] = t / 9 - 15;
        colOff[q] = 4 * (t % 9);
    }
    for (int k = 0; k < cnt; k++) {
        const WorkItem wi = work[k];
        const LevelPlan& L = P->lv[wi.level];
        const int xl = wi.x - 15;                            // level column of u = -15
        const uint8_t* base = pyr + L.roiOff + (ptrdiff_t)wi.y * L.pitch + (xl & ~3);
        const int2* wt = &gMomW[xl & 3][0][lane];
        int su = 0, sv = 0;
#pragma unroll
        for (int q = 0; q < 9; q++) {
            const unsigned a = __ldg(reinterpret_cast<const unsigned*>(base + rowOff[q] * L.pitch + colOff[q]));
            const int2 w = __ldg(wt + q * 32);
            su = dp4a_us(a, w.x, su);
            sv = dp4a_us(a, w.y, sv);
        }
        const int m10 = warp_sum(su), m01 = warp_sum(sv);
        if (lane == k) { M10 = m10; M01 = m01; }
    }
    // ---- 2. angle, cos, sin of lane's own keypoint ----
    float ca = 1.f, sa = 0.f;
    if (lane < cnt) {
        const float angle = fast_atan2_deg((float)M01, (float)M10);
        B.kps[(size_t)frame * P->kpCap + work[lane].pos].angle = angle;
        const float factorPI = (float)(3.141592653589793238462643383279502884197 / 180.f);     // :106
        const float rad = __fmul_rn(angle, factorPI);
        ca = (float)cos((double)rad); sa = (float)sin((double)rad);                            // :112 (correctly rounded)
    }
    // ---- 3. steered BRIEF on the blurred level (:107-146) ----
    float4 pat[8];
#pragma unroll
    for (int j = 0; j < 8; j++) pat[j] = __ldg(&gPatF[j * 32 + lane]);
    if constexpr (STAGE) {
        unsigned (*patch)[OD_PATCH_WORDS] = sPatch[warp];
        auto stage = [&](int k) {                              // window of keypoint k -> patch[k & 1]
            const WorkItem wi = work[k];
            const LevelPlan& L = P->lv[wi.level];
            // lane t of round q: row (32q + t) / 4, chunk (32q + t) % 4 = rows 8q + t / 4: the lane's source pointer advances by 8 rows per round
            const uint8_t* src = B.blur + (size_t)frame * P->blurStride + L.blurOff + (ptrdiff_t)(wi.y - 18 + (lane >> 2)) * L.bpitch +
                                 ((wi.x - 18) & ~15) + 16 * (lane & 3);
            uint8_t* dst = reinterpret_cast<uint8_t*>(patch[k & 1]) + 16 * lane;
            const ptrdiff_t step = (ptrdiff_t)8 * L.bpitch;
#pragma unroll
            for (int q = 0; q < OD_PROUNDS; q++) {
                if (q < OD_PROUNDS - 1 || q * 32 + lane < OD_PCHUNKS) cp_async16(dst + q * 512, src + q * step);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        stage(0);
        for (int k = 0; k < cnt; k++) {
            if (k + 1 < cnt) {
                stage(k + 1);
                asm volatile("cp.async.wait_group 1;" ::: "memory");
            } else {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
            }
            __syncwarp();
            const WorkItem wi = work[k];
            const float a = __shfl_sync(0xffffffffu, ca, k), b = __shfl_sync(0xffffffffu, sa, k);
            // byte (r, c) of the window lies at (r + 18) * 64 + c + 18 + (x - 18) % 16; r and c come out of the rounding trick biased by
            // 0x4B400000 each (see round_rne): everything constant goes into the base
            const uint8_t* pb = reinterpret_cast<const uint8_t*>(patch[k & 1]) +
                                (18 * OD_PPITCH + 18 + ((wi.x - 18) & 15) - (ptrdiff_t)0x4B400000 * (OD_PPITCH + 1));
            unsigned val = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 pt = pat[j];
                const int r0 = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(pt.x, b), __fmul_rn(pt.y, a)), 12582912.f));      // :118
                const int c0 = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(pt.x, a), __fmul_rn(pt.y, b)), 12582912.f));      // :119
                const int r1 = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(pt.z, b), __fmul_rn(pt.w, a)), 12582912.f));
                const int c1 = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(pt.z, a), __fmul_rn(pt.w, b)), 12582912.f));
                const int t0 = pb[(ptrdiff_t)r0 * OD_PPITCH + c0], t1 = pb[(ptrdiff_t)r1 * OD_PPITCH + c1];
                val |= (unsigned)(t0 < t1) << j;
            }
            B.desc[((size_t)frame * P->kpCap + wi.pos) * 32 + lane] = (uint8_t)val;
            __syncwarp();                                      // everyone is done with patch[k & 1] before keypoint k + 2 overwrites it
        }
    } else
    for (int k = 0; k < cnt; k++) {
        const WorkItem wi = work[k];
        const LevelPlan& L = P->lv[wi.level];
        const float a = __shfl_sync(0xffffffffu, ca, k), b = __shfl_sync(0xffffffffu, sa, k);
        const int step = L.bpitch;
        // the rounding trick (see round_rne) leaves the integer biased by 0x4B400000; the bias of row and column is folded
        // into the base pointer
        const uint8_t* bc = B.blur + (size_t)frame * P->blurStride + L.blurOff + (ptrdiff_t)wi.y * step + wi.x -
                            ((ptrdiff_t)0x4B400000 * step + 0x4B400000);
        unsigned val = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const float4 pt = pat[j];
            const int r0 = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(pt.x, b), __fmul_rn(pt.y, a)), 12582912.f));      // :118
            const int c0 = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(pt.x, a), __fmul_rn(pt.y, b)), 12582912.f));      // :119
            const int r1 = __float_as_int(__fadd_rn(__fadd_rn(__fmul_rn(pt.z, b), __fmul_rn(pt.w, a)), 12582912.f));
            const int c1 = __float_as_int(__fadd_rn(__fsub_rn(__fmul_rn(pt.z, a), __fmul_rn(pt.w, b)), 12582912.f));
            const int t0 = bc[(ptrdiff_t)r0 * step + c0], t1 = bc[(ptrdiff_t)r1 * step + c1];
            val |= (unsigned)(t0 < t1) << j;
        }
        B.desc[((size_t)frame * P->kpCap + wi.pos) * 32 + lane] = (uint8_t)val;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
static inline int cv_ceil_f(float v) { int i = (int)v; return i + (i < v); }
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ORBextractor::ORBextractor (:409-469): scale tables, features per level, umax
static void build_tables(orbb_extractor* h) {
    const int nl = h->prm.nlevels;
    const double scaleFactor = (double)h->prm.scale_factor;      // the member is a double initialised from a float
    h->scale.assign(nl, 1.f); h->sigma2.assign(nl, 1.f); h->invScale.assign(nl, 1.f); h->invSigma2.assign(nl, 1.f);
    for (int i = 1; i < nl; i++) {
        h->scale[i] = (float)(h->scale[i - 1] * scaleFactor);
        h->sigma2[i] = h->scale[i] * h->scale[i];
    }
    for (int i = 0; i < nl; i++) { h->invScale[i] = 1.0f / h->scale[i]; h->invSigma2[i] = 1.0f / h->sigma2[i]; }
    h->featPerLevel.assign(nl, 0);
    float factor = (float)(1.0f / scaleFactor);
    float nDesired = (float)(h->prm.nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nl)));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
        h->featPerLevel[l] = cv_round_f(nDesired);
        sum += h->featPerLevel[l];
        nDesired *= factor;
    }
    h->featPerLevel[nl - 1] = std::max(h->prm.nfeatures - sum, 0);
    const int HP = 15;
    int v, v0, vmax = cv_floor_f(HP * sqrtf(2.f) / 2 + 1), vmin = cv_ceil_f(HP * sqrtf(2.f) / 2);
    const double hp2 = HP * HP;
    for (v = 0; v <= vmax; ++v) h->umax[v] = (int)lrint(sqrt(hp2 - v * v));
    for (v = HP, v0 = 0; v >= vmin; --v) {
        while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
        h->umax[v] = v0;
        ++v0;
    }
}

static void free_bufs(orbb_extractor* h) {
    for (void* p : h->allocs) cudaFree(p);
    h->allocs.clear();
    memset(&h->b, 0, sizeof h->b);
    h->capacity = 0;
    h->planValid = false;
    if (h->g1Exec) cudaGraphExecDestroy(h->g1Exec);
    h->g1Exec = nullptr;
    h->g1Valid = false;
    h->tmapsValid = false;
}

template <typename T>
static int dev_alloc(orbb_extractor* h, T** p, size_t count) {
    void* q = nullptr;
    ORBB_CUDA(h, cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
    h->allocs.push_back(q);
    *p = (T*)q;
    return ORBB_OK;
}

// cv::resize INTER_LINEAR tables (imgproc/resize.cpp: resizeGeneric_ with HResizeLinear / VResizeLinear) for sw x sh -> dw x dh,
// appended to `tab`: dw column entries (source index, a0 | a1 << 16) padded to a multiple of 4, then dh row entries
// (source row, b0 | b1 << 16) padded likewise.  Returns the index of the first row entry.
static int append_resize_tables(std::vector<int2>& tab, int sw, int sh, int dw, int dh) {
    const double inv_sx = (double)dw / sw, inv_sy = (double)dh / sh;
    const double scale_x = 1. / inv_sx, scale_y = 1. / inv_sy;
    const size_t x0 = tab.size();
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor_f(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        const short a0 = (short)cv_round_f((1.f - fx) * 2048.f), a1 = (short)cv_round_f(fx * 2048.f);
        tab.push_back(make_int2(sx, (int)((unsigned short)a0 | ((unsigned)(unsigned short)a1 << 16))));
    }
    while ((tab.size() - x0) % 4) tab.push_back(tab.back());     // the kernels read 4 column entries per thread
    const int tabY = (int)tab.size();
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor_f(fy);
        fy -= sy;
        const short b0 = (short)cv_round_f((1.f - fy) * 2048.f), b1 = (short)cv_round_f(fy * 2048.f);
        tab.push_back(make_int2(sy, (int)((unsigned short)b0 | ((unsigned)(unsigned short)b1 << 16))));
    }
    while (tab.size() % 4) tab.push_back(tab.back());
    return tabY;
}

// One 3-D tensor map per level over the handle's pyramid slab: x = byte within the bordered row (pitch bytes), y = bordered row
// (h + 38), z = frame (stride = one frame's slab); box = one FAST cell tile (cellTp bytes x cellRows rows x 1 frame).  Rows or
// frames outside the tensor are zero-filled by the copy engine (only the don't-care rows below a short bottom cell can be).
// cuTensorMapEncodeTiled comes from the driver through the runtime (no link against libcuda).
static bool build_tensor_maps(orbb_extractor* h, int frames) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        cudaGetLastError();
        return (EncodeFn)fn;
    }();
    if (!encode) return false;
    const Plan& P = h->plan;
    if (P.pyrStride % 16) return false;
    for (int l = 0; l < P.nlevels; l++) {
        const LevelPlan& L = P.lv[l];
        const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)(L.h + 2 * kEdge), (cuuint64_t)frames};
        const cuuint64_t strides[2] = {(cuuint64_t)L.pitch, (cuuint64_t)P.pyrStride};
        const cuuint32_t box[3] = {(cuuint32_t)P.cellTp, (cuuint32_t)P.cellRows, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        if (encode(&h->tmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, h->b.pyr + L.pyrOff, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

// Geometry of every level for a WxH input + the cv::resize coefficient tables (imgproc/resize.cpp).
static int build_plan(orbb_extractor* h, int W, int H, int frames) {
    free_bufs(h);
    Plan& P = h->plan;
    memset(&P, 0, sizeof P);
    const int nl = h->prm.nlevels;
    P.nlevels = nl; P.W = W; P.H = H;
    P.iniTh = std::min(std::max(h->prm.ini_th_fast, 0), 255); P.minTh = std::min(std::max(h->prm.min_th_fast, 0), 255);
    P.k7Ini = (unsigned)(0x7f - (P.iniTh & 0x7f)) * 0x01010101u; P.k7Min = (unsigned)(0x7f - (P.minTh & 0x7f)) * 0x01010101u;
    for (int i = 0; i < 16; i++) P.umax[i] = h->umax[i];
    std::vector<int2> tab;
    std::vector<CellDesc> cellDesc;
    size_t pyrBytes = 0, blurBytes = 0;
    unsigned cellKeys = 0, raw = 0, nodes = 0, sel = 0;
    int cells = 0, tiles = 0, edgeTiles = 0, kpCap = 0;
    for (int l = 0; l < nl; l++) {
        LevelPlan& L = P.lv[l];
        L.w = cv_round_f((float)W * h->invScale[l]);                 // :1175
        L.h = cv_round_f((float)H * h->invScale[l]);
        if (L.w > 32767 || L.h > 32767) return set_err(h, ORBB_ERR_UNSUPPORTED, "image too large (%dx%d)", W, H);
        L.pitch = (int)align_up(kRoiX + L.w + kEdge, 128);
        L.pyrOff = (unsigned)pyrBytes;
        L.roiOff = L.pyrOff + kEdge * L.pitch + kRoiX;
        pyrBytes += align_up((size_t)L.pitch * (L.h + 2 * kEdge), 256);
        L.bpitch = (int)align_up(L.w, 128);
        L.blurOff = (unsigned)blurBytes;
        blurBytes += align_up((size_t)L.bpitch * L.h, 256);
        // FAST cell grid (:789-803)
        L.maxBX = L.w - kEdge + 3; L.maxBY = L.h - kEdge + 3;
        const float width = (float)(L.maxBX - kMinBorder), height = (float)(L.maxBY - kMinBorder);
        L.nCols = (int)(width / 35.f); L.nRows = (int)(height / 35.f);
        // A level smaller than one 35-px cell: the reference's cell loops (:805-822) do not execute -- its wCell / hCell =
        // ceil(x / 0) are never used -- and DistributeOctTree of no keys returns nothing, so the level simply contributes no
        // keypoints (its image is still built: the next level is resized from it, Frame.cc reads mvImagePyramid).  That holds
        // while both spans are positive; otherwise :559 gives a negative or undefined root count and vector::resize throws.
        const bool emptyLevel = L.nCols <= 0 || L.nRows <= 0;
        if (emptyLevel && (l == 0 || width <= 0.f || height <= 0.f))
            return set_err(h, ORBB_ERR_UNSUPPORTED, "level %d (%dx%d) is smaller than the FAST border / one 35-px cell: the reference has undefined behaviour here", l, L.w, L.h);
        if (emptyLevel) { L.nCols = L.nRows = 0; L.wCell = L.hCell = 1; }
        else { L.wCell = (int)ceilf(width / L.nCols); L.hCell = (int)ceilf(height / L.nRows); }
        if (L.wCell + 9 > kCellPix || L.hCell + 6 > kCellPix - 1) return set_err(h, ORBB_ERR_UNSUPPORTED, "cell %dx%d exceeds the kernel's tile", L.wCell, L.hCell);
        L.cellBase = cells;
        L.cellCap = ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);      // NMS survivors are never 8-adjacent
        L.cellKeyBase = cellKeys;
        for (int ci = 0; ci < L.nRows; ci++)
            for (int cj = 0; cj < L.nCols; cj++) {                   // :805-822
                const int iniX = kMinBorder + cj * L.wCell, iniY = kMinBorder + ci * L.hCell;
                const int maxX = std::min(iniX + L.wCell + 6, L.maxBX), maxY = std::min(iniY + L.hCell + 6, L.maxBY);
                CellDesc d;
                const bool skip = iniY >= L.maxBY - 3 || iniX >= L.maxBX - 6 || maxX - iniX < 7 || maxY - iniY < 7;
                d.gx0 = (short)(iniX + 3); d.gx1 = (short)(skip ? iniX + 3 : maxX - 3);
                d.gy0 = (short)(iniY + 3); d.gy1 = (short)(skip ? iniY + 3 : maxY - 3);
                d.level = l;
                d.outOff = L.cellKeyBase + (unsigned)((ci * L.nCols + cj) * L.cellCap);
                {   // phase-A constants of k_fast_cell (orbb_fast.cuh)
                    const int X0 = (d.gx0 - 3) & ~15, ih = d.gy1 - d.gy0;
                    d.cx0 = (short)(d.gx0 - X0); d.cx1 = (short)(d.gx1 - X0);
                    d.wa = (short)(d.cx0 >> 2);
                    d.wLast = (short)(skip ? 0 : ((d.cx1 - 1) >> 2) - d.wa);
                    d.nwc = std::max(d.wLast + 1, 2);                                // (>= 2 keeps the reciprocal in 32 bits; extra words are masked)
                    d.items = std::max((ih + 1) >> 1, 0) * d.nwc;
                    d.mInv = 0xffffffffu / (unsigned)d.nwc + 1u;
                    d.mF7 = 0x80808080u << (8 * (d.cx0 & 3));                        // first word: bytes of columns >= cx0
                    d.mL7 = skip ? 0u : 0x80808080u >> (8 * (3 - ((d.cx1 - 1) & 3)));   // last word: bytes of columns < cx1
                    d.pad = 0;
                }
                cellDesc.push_back(d);
            }
        cells += L.nCols * L.nRows;
        cellKeys += (unsigned)(L.nCols * L.nRows * L.cellCap);
        // quadtree (:559-561)
        L.nFeat = h->featPerLevel[l];
        L.nIni = emptyLevel ? 1 : (int)roundf((float)(L.maxBX - kMinBorder) / (L.maxBY - kMinBorder));      // (no keys: the root count is immaterial)
        if (L.nIni <= 0)
            return set_err(h, ORBB_ERR_UNSUPPORTED, "level %d aspect ratio gives %d root nodes: undefined behaviour in the reference (out-of-range write at ORBextractor.cc:585)", l, L.nIni);
        if (L.nIni > kMaxIni)
            return set_err(h, ORBB_ERR_UNSUPPORTED, "level %d is more than %d.5 times wider than high (%d root nodes): beyond this library's limit", l, kMaxIni, L.nIni);
        L.hX = (float)(L.maxBX - kMinBorder) / L.nIni;
        L.rawBase = raw; L.rawCap = L.nCols * L.nRows * L.cellCap; raw += (unsigned)L.rawCap;
        L.maxNodes = std::max(L.nFeat + 2, 4 * L.nIni) + 6;
        L.nodeBase = nodes; nodes += (unsigned)L.maxNodes;
        L.selBase = sel; L.selCap = L.maxNodes; sel += (unsigned)L.selCap;
        kpCap += L.selCap;
        L.scale = h->scale[l];
        L.kpSize = (float)(int)(31 * h->scale[l]);                   // :880 (PATCH_SIZE*mvScaleFactor -> int)
        L.blurTilesX = (L.w + 3) / 4; L.blurTilesY = (L.h - 2 * BLUR_BAND + BLUR_STRIP - 1) / BLUR_STRIP;     // word columns; 32-row strips of the rows 3 .. h-4
        if (L.blurTilesX < 4 || L.h < 2 * BLUR_BAND + 1) return set_err(h, ORBB_ERR_UNSUPPORTED, "level %d is too small for the blur kernel", l);
        L.blurTileBase = tiles; tiles += ((L.blurTilesX - 3) * L.blurTilesY + BLUR_THREADS - 1) / BLUR_THREADS;      // interior: word columns 1 .. nwords-3
        L.blurEdgeBase = edgeTiles;                                                                                  // the frame around it (k_blur<true>)
        edgeTiles += (3 * ((L.h + BLUR_STRIP_EDGE - 1) / BLUR_STRIP_EDGE) + 2 * (L.blurTilesX - 3) + BLUR_THREADS - 1) / BLUR_THREADS;
        // cv::resize tables for level l from level l-1
        if (l > 0) {
            const LevelPlan& S = P.lv[l - 1];
            L.tabX = (int)tab.size();
            L.tabY = append_resize_tables(tab, S.w, S.h, L.w, L.h);
            L.fastResize = 1;
            L.prmtTaps = 1;
            for (size_t g = (size_t)L.tabX; g < (size_t)L.tabY; g += 4) {
                if (tab[g + 3].x - 4 * (tab[g].x >> 2) > 7 || tab[g + 3].x < tab[g].x) L.fastResize = 0;      // taps must lie in bytes 0..8
                for (int k = 0; k < 3; k++) {
                    const int o = tab[g + k].x - 4 * (tab[g].x >> 2);
                    if (o < 0 || o + 1 > 7) L.prmtTaps = 0;                                                   // columns 0..2: both taps in bytes 0..7
                }
            }
            if (L.fastResize) {   // k_pyr_resize_t: the source window of every 128-column x 32-row CTA must fit its shared-memory tile
                bool fits = true;
                for (int g = 0; g * 128 < L.w; g++) {
                    const int first = tab[(size_t)L.tabX + g * 128].x & ~15;
                    const int last = tab[(size_t)L.tabX + std::min(g * 128 + 127, L.w - 1)].x + 12;      // three aligned words from the word of the last column
                    fits &= last - first <= PR_SPITCH;
                }
                for (int y = 0; y < L.h; y += 32) {
                    const int r0 = std::max(tab[(size_t)L.tabY + y].x, 0), r1 = std::min(tab[(size_t)L.tabY + std::min(y + 31, L.h - 1)].x + 1, S.h - 1);
                    fits &= r1 - r0 + 1 <= PR_SROWS;
                }
                if (fits) L.fastResize = 2;
            }

        }
    }
    {   // k_pyr_apron16 work items of the 19-px apron (built on demand, ensure_full_apron)
        int items = 0;
        for (int l = 0; l < nl; l++) {
            const LevelPlan& L = P.lv[l];
            ApronLevel& A = P.apron[l];
            A.itemBase = items;
            A.rows = kEdge;
            A.leftChunk0 = (kRoiX - kEdge) / 16;                                  // chunks that hold the columns -19 .. -1
            A.nLeft = 2 - A.leftChunk0;                                           // (kRoiX = 32: chunks 0,1)
            A.rightChunk0 = (kRoiX + L.w) / 16;                                   // first chunk with a column >= w
            A.nRight = (kRoiX + L.w + kEdge - 1) / 16 - A.rightChunk0 + 1;
            A.interiorChunks = std::max(A.rightChunk0 - 2, 1);                    // chunks 2 .. rightChunk0-1 lie inside the image
            A.invIC = A.interiorChunks > 1 ? 0xffffffffu / (unsigned)A.interiorChunks + 1u : 0u;
            A.invNR = A.nRight > 1 ? 0xffffffffu / (unsigned)A.nRight + 1u : 0u;
            items += 2 * kEdge * A.interiorChunks + (L.h + 2 * kEdge) * (A.nLeft + A.nRight);
        }
        P.apronItems = items;
    }
    {   // k_fast_cell: tile pitch 64 when every cell (+6 margin, +15 alignment) fits, else 96; smem for the tallest cell
        int maxW = 0, maxH = 0;
        for (int l = 0; l < nl; l++) { maxW = std::max(maxW, P.lv[l].wCell); maxH = std::max(maxH, P.lv[l].hCell); }
        P.cellTp = maxW + 21 <= 64 ? 64 : 96;
        P.cellRows = maxH + 6;
        P.cellSmem = (int)align_up(P.cellRows * P.cellTp + (maxH + 2) * (P.cellTp == 64 ? 48 : 80) + FC_STACK_BYTES, 128);
    }
    P.cellsTotal = cells; P.blurTilesTotal = tiles; P.blurEdgeTotal = edgeTiles; P.kpCap = kpCap;
    P.pyrStride = pyrBytes; P.blurStride = blurBytes;
    P.cellKeyStride = cellKeys; P.rawStride = raw; P.nodeStride = nodes; P.selStride = sel;

    Bufs& b = h->b;
    const size_t F = (size_t)frames;
    int rc;
#define A(ptr, count) if ((rc = dev_alloc(h, &ptr, (count))) != ORBB_OK) return rc
    A(b.pyr, F * pyrBytes + 512);          // slack: tile rows of the last level may be read a few bytes past their pitch
    A(b.blur, F * blurBytes + 256);        // slack: descriptor patch rows are read as whole words
    A(b.tab, tab.size());
    CellDesc* dCellDesc = nullptr;
    A(dCellDesc, cellDesc.size());
    b.cellDesc = dCellDesc;
    A(b.cellCount, F * cells);
    A(b.cellOff, F * cells);
    A(b.cellKeys, F * cellKeys);
    A(b.keys, F * 2 * raw);
    A(b.nodes, F * 2 * nodes);
    A(b.rec, F * nodes);
    A(b.cnt4, F * nodes);
    A(b.pend, F * 2 * nodes);
    A(b.elist, F * nodes);
    A(b.erased, F * nodes);
    A(b.sortTmp, F * 3 * nodes);
    A(b.sel, F * sel);
    A(b.selCount, F * ORBB_MAX_LEVELS);
    A(b.work, F * kpCap);
    A(b.kps, F * kpCap);
    A(b.desc, F * kpCap * 32);
    A(b.outCount, F * 3);                  // counts (2 per frame), then the status words: one copy brings both to the host
    b.status = b.outCount + F * 2;
    A(b.uRight, F * kpCap);
    A(b.depth, F * kpCap);
    A(b.bestR, F * kpCap);
    A(b.sad, F * kpCap);
    A(b.stRec, F * kpCap);
    A(b.stRowStart, F * (H + 2));
    A(h->dPlan, 1);
#undef A
    if (!tab.empty()) ORBB_CUDA(h, cudaMemcpyAsync(b.tab, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    ORBB_CUDA(h, cudaMemcpyAsync(dCellDesc, cellDesc.data(), cellDesc.size() * sizeof(CellDesc), cudaMemcpyHostToDevice, h->stream));
    ORBB_CUDA(h, (cudaFuncSetAttribute(k_octree<OT_THREADS_BIG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, OT_SMEM_BYTES)));
    ORBB_CUDA(h, (cudaFuncSetAttribute(k_fast_cell<64, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)));
    ORBB_CUDA(h, (cudaFuncSetAttribute(k_fast_cell<96, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)));
    ORBB_CUDA(h, (cudaFuncSetAttribute(k_fast_cell<64, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)));
    ORBB_CUDA(h, (cudaFuncSetAttribute(k_fast_cell<96, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100)));
    h->tmapsValid = build_tensor_maps(h, frames);
    ORBB_CUDA(h, cudaMemcpyAsync(h->dPlan, &P, sizeof P, cudaMemcpyHostToDevice, h->stream));
    ORBB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->capacity = frames;
    h->planValid = true;
    return ORBB_OK;
}

static int ensure_plan(orbb_extractor* h, int W, int H, int frames) {
    if (W <= 0 || H <= 0) return set_err(h, ORBB_ERR_EMPTY, "empty image");
    frames = std::max(frames, 1);
    if (h->planValid && h->plan.W == W && h->plan.H == H && h->capacity >= frames) return ORBB_OK;
    return build_plan(h, W, H, std::max(frames, h->prm.max_batch));
}

static void mark(orbb_extractor* h, int stage) {
    if (h->profiling) cudaEventRecord(h->ev[stage], h->stream);
}

// view of the per-frame buffers starting at frame f0 (kernels index frames from 0)
static Bufs shift_bufs(const Bufs& b, const Plan& P, int f0) {
    Bufs s = b;
    const size_t f = (size_t)f0;
    s.pyr += f * P.pyrStride; s.blur += f * P.blurStride;
    s.cellCount += f * P.cellsTotal; s.cellOff += f * P.cellsTotal;
    s.cellKeys += f * P.cellKeyStride; s.keys += f * 2 * P.rawStride; s.nodes += f * 2 * P.nodeStride;
    s.rec += f * P.nodeStride; s.cnt4 += f * P.nodeStride; s.pend += f * 2 * P.nodeStride; s.elist += f * P.nodeStride;
    s.erased += f * P.nodeStride; s.sortTmp += f * 3 * P.nodeStride; s.sel += f * P.selStride; s.selCount += f * ORBB_MAX_LEVELS;
    s.work += f * P.kpCap; s.kps += f * P.kpCap; s.desc += f * P.kpCap * 32; s.outCount += f * 2; s.status += f;
    s.uRight += f * P.kpCap; s.depth += f * P.kpCap; s.bestR += f * P.kpCap; s.sad += f * P.kpCap;
    s.stRec += f * P.kpCap; s.stRowStart += f * (P.H + 2);
    return s;
}

// kernel launch, optionally as a programmatic dependent of the previous kernel of the stream (see pdl_wait in orbb_internal.cuh)
template <typename... KArgs, typename... Args>
static void launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// the launch sequence for `nframes` device-resident frames whose buffers start at frame f0
static int run_lane(orbb_extractor* h, int lane, const uint8_t* dImgs, int nframes, size_t rowStride, size_t frameStride, int lap0, int lap1,
                    int f0) {
    const Plan& P = h->plan;
    const Bufs B = f0 ? shift_bufs(h->b, P, f0) : h->b;
    const orbb_extractor::Lane& ln = h->lanes[lane];
    cudaStream_t st = ln.st;
    // ORBB_NO_PDL=1: every kernel waits for the full completion of its predecessor before it is scheduled (A/B switch).  Stage
    // profiling records events between the kernels and keeps plain launches.
    static const bool noPdl = getenv("ORBB_NO_PDL") != nullptr;
    // (the single-frame CUDA graph keeps the programmatic edges: 0.143 -> 0.135 ms per call; ORBB_GRAPH_NO_PDL=1 captures plain edges)
    static const bool graphPdl = getenv("ORBB_GRAPH_NO_PDL") == nullptr;
    const bool pdl = !noPdl && !h->profiling && (!h->capturing || graphPdl);
    // ORBB_BLUR_EARLY=1: the blur forks right after the pyramid (beside the detector) instead of after the detector (beside the quadtree)
    static const bool blurEarly = getenv("ORBB_BLUR_EARLY") != nullptr;
    const bool fork = !h->profiling;
    // a call with a few frames: every level gets its own branch (detector + quadtree of level l start as soon as level l of the pyramid
    // exists, beside the resize chain of the smaller levels; the blur forks after the last resize).  The critical path is then the longest
    // level branch instead of pyramid + slowest cell of any level + slowest quadtree of any level.  ORBB_BRANCH_FRAMES=0 keeps one chain.
    static const int branchFrames = getenv("ORBB_BRANCH_FRAMES") ? atoi(getenv("ORBB_BRANCH_FRAMES")) : 4;
    const bool perLevel = fork && !blurEarly && lane == 0 && nframes <= branchFrames;
    // ORBB_FAST_NO_TMAP=1: stage the cell tiles with one bulk copy per row instead of one tensor-map copy (A/B switch; also the
    // path taken when the driver cannot encode the tensor maps)
    static const bool noTmap = getenv("ORBB_FAST_NO_TMAP") != nullptr;
    // a call with a few frames: four warps per cell (the kernel lasts as long as its slowest cell)
    static const int fastLatencyFrames = getenv("ORBB_FAST_LATENCY_FRAMES") ? atoi(getenv("ORBB_FAST_LATENCY_FRAMES")) : 4;
    auto launchFast = [&](cudaStream_t s, bool pdlFast, int cell0, int ncells) {
        const dim3 grid(ncells, nframes);
        const bool tm = h->tmapsValid && !noTmap;
        if (nframes <= fastLatencyFrames) {
            const int smem = P.cellSmem + (FC_MW - 1) * 2 * FC_CANDS;
            if (P.cellTp == 64) { if (tm) launch_k(pdlFast, k_fast_cell_mw<64, true>, grid, 32 * FC_MW, smem, s, h->dPlan, B, h->tmaps, f0, cell0); else launch_k(pdlFast, k_fast_cell_mw<64, false>, grid, 32 * FC_MW, smem, s, h->dPlan, B, h->tmaps, f0, cell0); }
            else { if (tm) launch_k(pdlFast, k_fast_cell_mw<96, true>, grid, 32 * FC_MW, smem, s, h->dPlan, B, h->tmaps, f0, cell0); else launch_k(pdlFast, k_fast_cell_mw<96, false>, grid, 32 * FC_MW, smem, s, h->dPlan, B, h->tmaps, f0, cell0); }
        } else if (tm) {
            if (P.cellTp == 64) launch_k(pdlFast, k_fast_cell<64, true>, grid, 32, P.cellSmem, s, h->dPlan, B, h->tmaps, f0, cell0);
            else launch_k(pdlFast, k_fast_cell<96, true>, grid, 32, P.cellSmem, s, h->dPlan, B, h->tmaps, f0, cell0);
        } else {
            if (P.cellTp == 64) launch_k(pdlFast, k_fast_cell<64, false>, grid, 32, P.cellSmem, s, h->dPlan, B, h->tmaps, f0, cell0);
            else launch_k(pdlFast, k_fast_cell<96, false>, grid, 32, P.cellSmem, s, h->dPlan, B, h->tmaps, f0, cell0);
        }
        h->launches++;
    };
    // images whose first level has many cells (4K class) get the large CTA on every level: more warps split nodes at once
    // 512 threads per (frame, level) for levels with many cells -- and for a call with a few frames, where the quadtree of
    // level 0 is one CTA on the critical path and a warp issues one dependent instruction every few cycles: more warps split
    // more nodes at once, and nodes with many keys are split by the whole CTA
    static const int otLatencyFrames = getenv("ORBB_OCTREE_LATENCY_FRAMES") ? atoi(getenv("ORBB_OCTREE_LATENCY_FRAMES")) : 4;
    // ORBB_OCTREE_NO_SMEM=1: the quadtree of a call with a few frames keeps keys and nodes in global memory and sorts with one warp (A/B switch)
    static const bool otNoSmem = getenv("ORBB_OCTREE_NO_SMEM") != nullptr;
    auto launchTree = [&](cudaStream_t s, bool pdlTree, int level0, int nlev) {
        if (P.lv[0].nCols * P.lv[0].nRows >= OT_BIG_CELLS) launch_k(pdlTree, k_octree<OT_THREADS_BIG, false>, dim3(nframes, nlev), OT_THREADS_BIG, 0, s, h->dPlan, B, level0, (int)OT_BIG_NODE);
        else if (nframes <= otLatencyFrames && !otNoSmem) launch_k(pdlTree, k_octree<OT_THREADS_BIG, true>, dim3(nframes, nlev), OT_THREADS_BIG, OT_SMEM_BYTES, s, h->dPlan, B, level0, (int)OT_BIG_NODE_LATENCY);
        else if (nframes <= otLatencyFrames) launch_k(pdlTree, k_octree<OT_THREADS_BIG, false>, dim3(nframes, nlev), OT_THREADS_BIG, 0, s, h->dPlan, B, level0, (int)OT_BIG_NODE_LATENCY);
        else launch_k(pdlTree, k_octree<OT_THREADS, false>, dim3(nframes, nlev), OT_THREADS, 0, s, h->dPlan, B, level0, (int)OT_BIG_NODE);
        h->launches++;
    };
    mark(h, ST_PYRAMID);
    ORBB_CUDA(h, cudaMemsetAsync(B.status, 0, sizeof(int) * nframes, st));
    for (int l = 0; l < P.nlevels; l++) {
        const LevelPlan& L = P.lv[l];
        dim3 grid(((L.w + 3) / 4 + 31) / 32, (L.h + 7) / 8, nframes);
        if (l == 0) {
            const int aligned16 = ((uintptr_t)dImgs % 16 == 0) && rowStride % 16 == 0 && frameStride % 16 == 0;
            k_pyr_level0_v<<<dim3(((L.w + 15) / 16 + 31) / 32, (L.h + 7) / 8, nframes), 256, 0, st>>>(h->dPlan, B, dImgs, rowStride, frameStride,
                                                                                                  aligned16);
        } else if (L.fastResize) {
            constexpr int rowsPerCta = PR_ROWS * (PR_THREADS / 32);
            static const bool tmaResize = getenv("ORBB_RESIZE_NO_TMA") == nullptr;      // (A/B switch; TMA staging measured 3 % faster)
            constexpr int rowsPerCtaLat = PR_ROWS_LATENCY * (PR_THREADS / 32);
            static const int pyrLatencyFrames = getenv("ORBB_PYR_LATENCY_FRAMES") ? atoi(getenv("ORBB_PYR_LATENCY_FRAMES")) : 4;
            static const bool noPrmt = getenv("ORBB_RESIZE_NO_PRMT") != nullptr;      // (A/B switch)
            const bool prmt = L.prmtTaps && !noPrmt;
            const dim3 gLat(grid.x, (L.h + rowsPerCtaLat - 1) / rowsPerCtaLat, nframes), gBat(grid.x, (L.h + rowsPerCta - 1) / rowsPerCta, nframes);
            if (tmaResize && L.fastResize == 2 && nframes <= pyrLatencyFrames) {
                if (prmt) launch_k(pdl, k_pyr_resize_t<PR_ROWS_LATENCY, true>, gLat, PR_THREADS, 0, st, h->dPlan, B, l);
                else launch_k(pdl, k_pyr_resize_t<PR_ROWS_LATENCY, false>, gLat, PR_THREADS, 0, st, h->dPlan, B, l);
            } else if (tmaResize && L.fastResize == 2) {
                if (prmt) launch_k(pdl, k_pyr_resize_t<PR_ROWS, true>, gBat, PR_THREADS, 0, st, h->dPlan, B, l);
                else launch_k(pdl, k_pyr_resize_t<PR_ROWS, false>, gBat, PR_THREADS, 0, st, h->dPlan, B, l);
            }
            else launch_k(pdl, k_pyr_resize_s, dim3(grid.x, (L.h + rowsPerCta - 1) / rowsPerCta, nframes), PR_THREADS, 0, st, h->dPlan, B, l);
        } else launch_k(pdl, k_pyr_resize, grid, 256, 0, st, h->dPlan, B, l);
        h->launches++;
        if (perLevel) {
            cudaStream_t ls = h->lvlSt[l];
            ORBB_CUDA(h, cudaEventRecord(h->evLvl[l], st));
            ORBB_CUDA(h, cudaStreamWaitEvent(ls, h->evLvl[l], 0));
            if (L.nCols * L.nRows > 0) launchFast(ls, false, L.cellBase, L.nCols * L.nRows);
            launchTree(ls, pdl && L.nCols * L.nRows > 0, l, 1);
            ORBB_CUDA(h, cudaEventRecord(h->evTree[l], ls));
        }
    }
    if (fork && (blurEarly || perLevel)) ORBB_CUDA(h, cudaEventRecord(ln.evFork, st));
    mark(h, ST_FAST);
    if (!perLevel) launchFast(st, pdl && !(fork && blurEarly), 0, P.cellsTotal);      // (an event record between two kernels makes the edge a full dependency)
    mark(h, ST_OCTREE);
    // fork: the blur only needs the pyramid; on its own stream it fills the SMs that the latency-bound quadtree leaves idle.
    // With stage profiling on, everything stays on one stream so that the stage events mean what they say.
    if (fork && !blurEarly && !perLevel) ORBB_CUDA(h, cudaEventRecord(ln.evFork, st));
    if (!perLevel) launchTree(st, pdl && (!fork || blurEarly), 0, P.nlevels);
    else for (int l = 0; l < P.nlevels; l++) ORBB_CUDA(h, cudaStreamWaitEvent(st, h->evTree[l], 0));
    if (fork) {
        ORBB_CUDA(h, cudaStreamWaitEvent(ln.blurSt, ln.evFork, 0));
        ORBB_CUDA(h, cudaStreamWaitEvent(ln.blurEdgeSt, ln.evFork, 0));
        k_blur<true><<<dim3(P.blurEdgeTotal, nframes), BLUR_THREADS, 0, ln.blurEdgeSt>>>(h->dPlan, B);
        k_blur<false><<<dim3(P.blurTilesTotal, nframes), BLUR_THREADS, 0, ln.blurSt>>>(h->dPlan, B);
        ORBB_CUDA(h, cudaEventRecord(ln.evJoin, ln.blurSt));
        ORBB_CUDA(h, cudaEventRecord(ln.evJoinEdge, ln.blurEdgeSt));
    }
    mark(h, ST_BLUR);
    if (fork) {
        ORBB_CUDA(h, cudaStreamWaitEvent(st, ln.evJoin, 0));
        ORBB_CUDA(h, cudaStreamWaitEvent(st, ln.evJoinEdge, 0));
    } else {
        k_blur<false><<<dim3(P.blurTilesTotal, nframes), BLUR_THREADS, 0, st>>>(h->dPlan, B);
        k_blur<true><<<dim3(P.blurEdgeTotal, nframes), BLUR_THREADS, 0, st>>>(h->dPlan, B);
    }
    // output assembly: its own kernel in batches (one CTA per frame, free beside the other lane); folded into the prologue of the
    // next kernel in a call with a few frames, where it is 11 us on the critical path (measured: fused 1.289 vs 1.249 ms per 256
    // frames, but 0.128 vs 0.135 ms for one frame)
    static const int asmLatencyFrames = getenv("ORBB_ASM_LATENCY_FRAMES") ? atoi(getenv("ORBB_ASM_LATENCY_FRAMES")) : 4;
    const bool fusedAsm = nframes <= asmLatencyFrames;
    mark(h, ST_ASSEMBLE);
    if (!fusedAsm) { k_assemble<<<nframes, 256, 0, st>>>(h->dPlan, B, lap0, lap1); h->launches++; }
    mark(h, ST_ORIENT_DESC);
    {   // ORBB_DESC_NO_STAGE=1: sample the blurred level through L1 instead of a shared-memory copy of the window (A/B switch)
        static const bool noStage = getenv("ORBB_DESC_NO_STAGE") != nullptr;
        const int ctaKps = OD_THREADS / 32 * (fusedAsm ? OD_KPW_LAT : OD_KPW);
        const dim3 grid((P.kpCap + ctaKps - 1) / ctaKps, nframes);
        if (fusedAsm) {                                        // (behind the joins of the blur streams the edge is a full dependency anyway)
            if (noStage) launch_k(pdl && !fork, k_orient_desc32<false, true>, grid, OD_THREADS, 0, st, h->dPlan, B, lap0, lap1);
            else launch_k(pdl && !fork, k_orient_desc32<true, true>, grid, OD_THREADS, 0, st, h->dPlan, B, lap0, lap1);
        } else {
            if (noStage) launch_k(pdl, k_orient_desc32<false, false>, grid, OD_THREADS, 0, st, h->dPlan, B, lap0, lap1);
            else launch_k(pdl, k_orient_desc32<true, false>, grid, OD_THREADS, 0, st, h->dPlan, B, lap0, lap1);
        }
    }
    mark(h, ST_D2H);
    h->launches += 3;
    ORBB_CUDA(h, cudaGetLastError());
    return ORBB_OK;
}

// the launch sequence for `nframes` device-resident frames whose buffers start at frame f0: contiguous parts of the batch
// go to the handle's lanes (all ordered after what is already queued on h->stream, and joined back into it)
static int run_batch(orbb_extractor* h, const uint8_t* dImgs, int nframes, size_t rowStride, size_t frameStride, int lap0, int lap1,
                     int f0 = 0, bool hostChunk = false) {
    // lanes: a resident batch of >= 64 frames per lane is split over the handle's lanes (2 by default: 1.329 -> 1.277 ms per 256
    // frames, the latency-bound quadtree / descriptor kernels of one half run under the issue-bound detector of the other); the
    // chunks of the host pipeline are already interleaved with their copies and stay on one lane unless ORBB_LANES_HOST=1
    static const bool lanesHost = getenv("ORBB_LANES_HOST") != nullptr;
    static const int lanesMin = getenv("ORBB_LANES_MIN") ? std::max(1, atoi(getenv("ORBB_LANES_MIN"))) : 64;      // frames per lane below which a batch is not split
    const int lanes = (h->profiling || h->capturing || (hostChunk && !lanesHost) || nframes < lanesMin * h->nLanes) ? 1 : h->nLanes;
    if (lanes > 1) ORBB_CUDA(h, cudaEventRecord(h->lanes[0].evStart, h->stream));
    int done = 0;
    for (int l = 0; l < lanes; l++) {
        const int n = (nframes - done) / (lanes - l);
        if (l > 0) ORBB_CUDA(h, cudaStreamWaitEvent(h->lanes[l].st, h->lanes[0].evStart, 0));
        const int rc = run_lane(h, l, dImgs + (size_t)done * frameStride, n, rowStride, frameStride, lap0, lap1, f0 + done);
        if (rc) return rc;
        if (l > 0) {
            ORBB_CUDA(h, cudaEventRecord(h->lanes[l].evDone, h->lanes[l].st));
            ORBB_CUDA(h, cudaStreamWaitEvent(h->stream, h->lanes[l].evDone, 0));
        }
        done += n;
    }
    h->lastFrames = f0 + nframes;
    h->hPyrFresh = false;
    h->apronFull = false;
    return ORBB_OK;
}

}  // namespace orbb

using namespace orbb;

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* orbb_version(void) { return "orbb200 0.1 (sm_100a)"; }

const char* orbb_last_error(const orbb_extractor* h) { return h ? h->err.c_str() : g_lastError.c_str(); }

int orbb_create(const orbb_params* prm, orbb_extractor** out) {
    if (!prm || !out) return set_err(nullptr, ORBB_ERR_ARG, "null argument");
    *out = nullptr;
    if (prm->nlevels < 1 || prm->nlevels > ORBB_MAX_LEVELS || prm->nfeatures < 1 || !(prm->scale_factor > 1.0f))
        return set_err(nullptr, ORBB_ERR_ARG, "bad parameters (nlevels 1..%d, nfeatures >= 1, scale_factor > 1)", ORBB_MAX_LEVELS);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return set_err(nullptr, ORBB_ERR_CUDA, "no CUDA device (%s): liborbb200 has no CPU fallback", cudaGetErrorString(e));
    if (prm->device < 0 || prm->device >= ndev) return set_err(nullptr, ORBB_ERR_ARG, "device %d out of range", prm->device);
    orbb_extractor* h = new orbb_extractor();
    h->prm = *prm;
    h->prm.max_batch = std::max(prm->max_batch, 1);
    h->device = prm->device;
    build_tables(h);
    if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_err(nullptr, ORBB_ERR_CUDA, "cannot create stream on device %d", h->device);
        delete h;
        return ORBB_ERR_CUDA;
    }
    for (int i = 0; i <= ST_COUNT; i++) cudaEventCreate(&h->ev[i]);
    cudaStreamCreateWithFlags(&h->h2dStream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&h->d2hStream, cudaStreamNonBlocking);
    if (const char* e = getenv("ORBB_LANES")) h->nLanes = std::min(4, std::max(1, atoi(e)));
    for (int l = 0; l < 4; l++) {
        orbb_extractor::Lane& ln = h->lanes[l];
        if (l == 0) ln.st = h->stream;
        else cudaStreamCreateWithFlags(&ln.st, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&ln.blurSt, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&ln.blurEdgeSt, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&ln.evFork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ln.evJoin, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ln.evJoinEdge, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ln.evStart, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ln.evDone, cudaEventDisableTiming);
    }
    for (int l = 0; l < ORBB_MAX_LEVELS; l++) {
        cudaStreamCreateWithFlags(&h->lvlSt[l], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&h->evLvl[l], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->evTree[l], cudaEventDisableTiming);
    }
    for (int i = 0; i < 8; i++) {
        cudaEventCreateWithFlags(&h->evH2D[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&h->evDone[i], cudaEventDisableTiming);
    }
    k_init_pattern<<<1, 256, 0, h->stream>>>();
    k_init_moment_weights<<<1, 256, 0, h->stream>>>();
    if (cudaStreamSynchronize(h->stream) != cudaSuccess) {
        set_err(nullptr, ORBB_ERR_CUDA, "pattern init failed: %s", cudaGetErrorString(cudaGetLastError()));
        orbb_destroy(h);
        return ORBB_ERR_CUDA;
    }
    *out = h;
    return ORBB_OK;
}

void orbb_destroy(orbb_extractor* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_bufs(h);
    if (h->hImg) cudaFree(h->hImg);
    if (h->dColor) cudaFree(h->dColor);
    if (h->rsTab) cudaFree(h->rsTab);
    if (h->hPyr) cudaFreeHost(h->hPyr);
    if (h->hCounts) cudaFreeHost(h->hCounts);
    for (int i = 0; i <= ST_COUNT; i++) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 8; i++) { if (h->evH2D[i]) cudaEventDestroy(h->evH2D[i]); if (h->evDone[i]) cudaEventDestroy(h->evDone[i]); }
    if (h->h2dStream) cudaStreamDestroy(h->h2dStream);
    if (h->d2hStream) cudaStreamDestroy(h->d2hStream);
    for (int l = 0; l < 4; l++) {
        orbb_extractor::Lane& ln = h->lanes[l];
        if (l > 0 && ln.st) { cudaStreamSynchronize(ln.st); cudaStreamDestroy(ln.st); }
        if (ln.blurSt) { cudaStreamSynchronize(ln.blurSt); cudaStreamDestroy(ln.blurSt); }
        if (ln.blurEdgeSt) { cudaStreamSynchronize(ln.blurEdgeSt); cudaStreamDestroy(ln.blurEdgeSt); }
        if (ln.evFork) cudaEventDestroy(ln.evFork);
        if (ln.evJoin) cudaEventDestroy(ln.evJoin);
        if (ln.evJoinEdge) cudaEventDestroy(ln.evJoinEdge);
        if (ln.evStart) cudaEventDestroy(ln.evStart);
        if (ln.evDone) cudaEventDestroy(ln.evDone);
    }
    for (int l = 0; l < ORBB_MAX_LEVELS; l++) {
        if (h->lvlSt[l]) { cudaStreamSynchronize(h->lvlSt[l]); cudaStreamDestroy(h->lvlSt[l]); }
        if (h->evLvl[l]) cudaEventDestroy(h->evLvl[l]);
        if (h->evTree[l]) cudaEventDestroy(h->evTree[l]);
    }
    cudaStreamDestroy(h->stream);
    delete h;
}

int orbb_get_tables(const orbb_extractor* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* feats) {
    if (!h) return ORBB_ERR_ARG;
    for (int i = 0; i < h->prm.nlevels; i++) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->invScale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->invSigma2[i];
        if (feats) feats[i] = h->featPerLevel[i];
    }
    return ORBB_OK;
}

int orbb_max_keypoints(const orbb_extractor* h) {
    if (!h) return ORBB_ERR_ARG;
    if (h->planValid) return h->plan.kpCap;
    int cap = 0;
    for (int l = 0; l < h->prm.nlevels; l++) cap += std::max(h->featPerLevel[l] + 2, 4 * kMaxIni) + 6;
    return cap;
}

long long orbb_launch_count(const orbb_extractor* h) { return h ? h->launches : 0; }
void* orbb_stream(orbb_extractor* h) { return h ? (void*)h->stream : nullptr; }

int orbb_set_profiling(orbb_extractor* h, int enabled) {
    if (!h) return ORBB_ERR_ARG;
    h->profiling = enabled != 0;
    return ORBB_OK;
}

const char* orbb_stage_name(int i) {
    static const char* names[ST_COUNT] = {"h2d", "pyramid", "fast", "octree", "blur", "assemble", "orient_desc", "d2h"};
    return (i >= 0 && i < ST_COUNT) ? names[i] : "";
}

int orbb_stage_times(orbb_extractor* h, float* ms, int cap) {
    if (!h || !ms) return ORBB_ERR_ARG;
    ORBB_CUDA(h, cudaSetDevice(h->device));
    ORBB_CUDA(h, cudaStreamSynchronize(h->stream));
    int n = std::min(cap, (int)ST_COUNT);
    for (int i = 0; i < n; i++) {
        ms[i] = 0.f;
        if (i >= ST_PYRAMID && i < ST_D2H) cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]);
    }
    cudaGetLastError();
    return n;
}

int orbb_extract_batch(orbb_extractor* h, const uint8_t* dev_imgs, int nframes, int width, int height, size_t row_stride,
                       size_t frame_stride, int lap0, int lap1) {
    if (!h) return ORBB_ERR_ARG;
    if (!dev_imgs || nframes <= 0 || width <= 0 || height <= 0) return set_err(h, ORBB_ERR_EMPTY, "empty image");
    ORBB_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_plan(h, width, height, nframes);
    if (rc) return rc;
    return run_batch(h, dev_imgs, nframes, row_stride, frame_stride, lap0, lap1);
}

// device staging area for frames that arrive from the host / from an input-side kernel.  The single-frame CUDA graph captured
// the old pointer (its level-0 kernel reads hImg): a reallocation invalidates it.
static int ensure_staging(orbb_extractor* h, size_t need) {
    if (h->hImgBytes >= need) return ORBB_OK;
    if (h->g1Exec) { cudaGraphExecDestroy(h->g1Exec); h->g1Exec = nullptr; }
    h->g1Valid = false;
    if (h->hImg) { cudaStreamSynchronize(h->stream); cudaFree(h->hImg); }
    h->hImg = nullptr; h->hImgBytes = 0;
    ORBB_CUDA(h, cudaMalloc((void**)&h->hImg, need));
    h->hImgBytes = need;
    return ORBB_OK;
}

int orbb_extract_batch_color(orbb_extractor* h, const uint8_t* dev_imgs, int nframes, int width, int height, size_t row_stride,
                             size_t frame_stride, int channels, int rgb_order, int lap0, int lap1) {
    if (!h) return ORBB_ERR_ARG;
    if (!dev_imgs || nframes <= 0 || width <= 0 || height <= 0) return set_err(h, ORBB_ERR_EMPTY, "empty image");
    if (channels != 3 && channels != 4) return set_err(h, ORBB_ERR_ARG, "channels must be 3 or 4 (got %d)", channels);
    ORBB_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_plan(h, width, height, nframes);
    if (rc) return rc;
    if ((rc = ensure_staging(h, (size_t)nframes * width * height))) return rc;
    dim3 grid(((width + 3) / 4 + 31) / 32, (height + 7) / 8, nframes);
    k_gray<<<grid, 256, 0, h->stream>>>(dev_imgs, row_stride, frame_stride, channels, rgb_order, h->hImg, width, height);
    h->launches++;
    return run_batch(h, h->hImg, nframes, (size_t)width, (size_t)width * height, lap0, lap1);
}

int orbb_extract_color(orbb_extractor* h, const uint8_t* img, int width, int height, size_t stride, int channels, int rgb_order,
                       int lap0, int lap1, orbb_keypoint* kps, uint8_t* desc, int capacity, int* n_out, int* mono_index) {
    if (!h) return ORBB_ERR_ARG;
    if (n_out) *n_out = 0;
    if (mono_index) *mono_index = 0;
    if (!img || width <= 0 || height <= 0) return set_err(h, ORBB_ERR_EMPTY, "empty image");
    if (channels != 3 && channels != 4) return set_err(h, ORBB_ERR_ARG, "channels must be 3 or 4 (got %d)", channels);
    ORBB_CUDA(h, cudaSetDevice(h->device));
    const size_t rowBytes = (size_t)width * channels;
    if (h->colorBytes < rowBytes * height) {
        if (h->dColor) cudaFree(h->dColor);
        h->dColor = nullptr; h->colorBytes = 0;
        ORBB_CUDA(h, cudaMalloc((void**)&h->dColor, rowBytes * height));
        h->colorBytes = rowBytes * height;
    }
    ORBB_CUDA(h, cudaMemcpy2DAsync(h->dColor, rowBytes, img, stride, rowBytes, height, cudaMemcpyHostToDevice, h->stream));
    int rc = orbb_extract_batch_color(h, h->dColor, 1, width, height, rowBytes, rowBytes * height, channels, rgb_order, lap0, lap1);
    if (rc) return rc;
    int32_t counts[2] = {0, 0};
    rc = orbb_batch_fetch(h, 1, kps, desc, capacity, counts);
    if (n_out) *n_out = counts[0];
    if (mono_index) *mono_index = counts[1];
    return rc;
}

int orbb_sync(orbb_extractor* h) {
    if (!h) return ORBB_ERR_ARG;
    ORBB_CUDA(h, cudaSetDevice(h->device));
    ORBB_CUDA(h, cudaStreamSynchronize(h->stream));
    return ORBB_OK;
}

static cudaError_t copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, cudaMemcpyKind kind,
                             cudaStream_t st);

static int ensure_counts(orbb_extractor* h, int nframes) {
    if (h->hCountsCap >= nframes) return ORBB_OK;
    if (h->hCounts) cudaFreeHost(h->hCounts);
    h->hCounts = nullptr;
    ORBB_CUDA(h, cudaMallocHost((void**)&h->hCounts, sizeof(int32_t) * 3 * nframes));
    h->hCountsCap = nframes;
    return ORBB_OK;
}

int orbb_batch_fetch(orbb_extractor* h, int nframes, orbb_keypoint* kps, uint8_t* desc, int capacity, int32_t* counts) {
    if (!h || !counts) return ORBB_ERR_ARG;
    if (nframes > h->lastFrames) return set_err(h, ORBB_ERR_ARG, "fetch of %d frames, last batch had %d", nframes, h->lastFrames);
    ORBB_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_counts(h, nframes);
    if (rc) return rc;
    const Plan& P = h->plan;
    ORBB_CUDA(h, cudaMemcpyAsync(h->hCounts, h->b.outCount, sizeof(int) * 2 * nframes, cudaMemcpyDeviceToHost, h->stream));
    ORBB_CUDA(h, cudaMemcpyAsync(h->hCounts + 2 * nframes, h->b.status, sizeof(int) * nframes, cudaMemcpyDeviceToHost, h->stream));
    const int ncopy = std::min(capacity, P.kpCap);
    if (kps && ncopy > 0)
        ORBB_CUDA(h, copy_rows(kps, sizeof(orbb_keypoint) * capacity, h->b.kps, sizeof(orbb_keypoint) * P.kpCap,
                               sizeof(orbb_keypoint) * ncopy, nframes, cudaMemcpyDeviceToHost, h->stream));
    if (desc && ncopy > 0)
        ORBB_CUDA(h, copy_rows(desc, (size_t)32 * capacity, h->b.desc, (size_t)32 * P.kpCap, (size_t)32 * ncopy, nframes,
                               cudaMemcpyDeviceToHost, h->stream));
    ORBB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int f = 0; f < nframes; f++) {
        counts[2 * f] = h->hCounts[2 * f];
        counts[2 * f + 1] = h->hCounts[2 * f + 1];
        if (h->hCounts[2 * nframes + f]) return set_err(h, ORBB_ERR_INTERNAL, "frame %d: device status 0x%x (capacity overflow)", f, h->hCounts[2 * nframes + f]);
        if (counts[2 * f] > capacity && (kps || desc)) return set_err(h, ORBB_ERR_CAPACITY, "frame %d has %d keypoints, capacity %d", f, counts[2 * f], capacity);
    }
    return ORBB_OK;
}

int orbb_batch_device_ptrs(orbb_extractor* h, const orbb_keypoint** kps, const uint8_t** desc, const int32_t** counts) {
    if (!h || !h->planValid) return ORBB_ERR_ARG;
    if (kps) *kps = h->b.kps;
    if (desc) *desc = h->b.desc;
    if (counts) *counts = h->b.outCount;
    return ORBB_OK;
}

// strided-or-contiguous async copy of `rows` rows of `width` bytes
static cudaError_t copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, cudaMemcpyKind kind,
                             cudaStream_t st) {
    if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * rows, kind, st);      // one linear DMA
    return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st);
}

int orbb_extract_batch_host_submit(orbb_extractor* h, const uint8_t* host_imgs, int nframes, int width, int height, size_t row_stride,
                                   size_t frame_stride, int lap0, int lap1, orbb_keypoint* kps, uint8_t* desc, int capacity) {
    if (!h) return ORBB_ERR_ARG;
    h->pendingFrames = 0;
    if (!host_imgs || nframes <= 0 || width <= 0 || height <= 0) return set_err(h, ORBB_ERR_EMPTY, "empty image");
    ORBB_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_plan(h, width, height, nframes);
    if (rc) return rc;
    if ((rc = ensure_counts(h, nframes))) return rc;
    const Plan& P = h->plan;
    // device staging area for the raw frames: tightly packed WxH
    const size_t fbytes = (size_t)width * height, need = (size_t)nframes * fbytes;
    if ((rc = ensure_staging(h, need))) return rc;
    static const bool noGraph = getenv("ORBB_NO_GRAPH") != nullptr;
    if (nframes == 1 && !h->profiling && !noGraph) {
        // ---- latency path: one stream, the kernels of the frame replayed as a CUDA graph ----
        cudaStream_t st = h->stream;
        // (uploading straight into the pitched level 0 of the pyramid instead -- no level-0 kernel -- was measured: the 2-D copy takes
        // 44 us longer than this linear one)
        ORBB_CUDA(h, copy_rows(h->hImg, width, host_imgs, row_stride, width, height, cudaMemcpyHostToDevice, st));
        if (!h->g1Valid || h->g1Lap0 != lap0 || h->g1Lap1 != lap1) {
            if (h->g1Exec) { cudaGraphExecDestroy(h->g1Exec); h->g1Exec = nullptr; }
            h->g1Valid = false;
            cudaGraph_t graph = nullptr;
            const long long before = h->launches;
            ORBB_CUDA(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            h->capturing = true;
            rc = run_batch(h, h->hImg, 1, (size_t)width, fbytes, lap0, lap1, 0);
            h->capturing = false;
            const cudaError_t ce = cudaStreamEndCapture(st, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            if (ce != cudaSuccess) return set_err(h, ORBB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
            const cudaError_t ie = cudaGraphInstantiate(&h->g1Exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) return set_err(h, ORBB_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(ie));
            h->g1Launches = h->launches - before;
            h->launches = before;
            h->g1Lap0 = lap0; h->g1Lap1 = lap1; h->g1Valid = true;
        }
        ORBB_CUDA(h, cudaGraphLaunch(h->g1Exec, st));
        h->launches += h->g1Launches;
        h->lastFrames = 1;
        h->hPyrFresh = false;
        h->apronFull = false;
        const int ncopy1 = std::min(capacity, P.kpCap);
        if (h->b.status == h->b.outCount + 2) {             // (a handle sized for one frame: counts, then status, in one copy)
            ORBB_CUDA(h, cudaMemcpyAsync(h->hCounts, h->b.outCount, sizeof(int) * 3, cudaMemcpyDeviceToHost, st));
        } else {
            ORBB_CUDA(h, cudaMemcpyAsync(h->hCounts, h->b.outCount, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
            ORBB_CUDA(h, cudaMemcpyAsync(h->hCounts + 2, h->b.status, sizeof(int), cudaMemcpyDeviceToHost, st));
        }
        if (kps && ncopy1 > 0) ORBB_CUDA(h, cudaMemcpyAsync(kps, h->b.kps, sizeof(orbb_keypoint) * ncopy1, cudaMemcpyDeviceToHost, st));
        if (desc && ncopy1 > 0) ORBB_CUDA(h, cudaMemcpyAsync(desc, h->b.desc, (size_t)32 * ncopy1, cudaMemcpyDeviceToHost, st));
        h->pendingFrames = 1;
        h->pendingCapacity = (kps || desc) ? capacity : INT_MAX;
        return ORBB_OK;
    }
    // Pipeline in chunks of frames (about 64 frames each, at most 4; up to 8 with ORBB_CHUNKS): H2D (copy engine 1) -> kernels (handle stream) ->
    // D2H (copy engine 2), so the upload of chunk c+1 and the download of chunk c-1 overlap the kernels of chunk c; the first upload and the
    // last chunk's kernels are what cannot be hidden, so smaller chunks help until the per-launch tails of small batches cost more.
    static const int chunkOverride = getenv("ORBB_CHUNKS") ? atoi(getenv("ORBB_CHUNKS")) : 0;
    // measured end to end, 256 frames per call, two alternating handles: 2 chunks 141.5 k frames/s, 3: 144.3 k, 4: 145.1 k (0.969 of the copy
    // ceiling), 5: 141.6 k, 6: 140.3 k, 8: 136.0 k
    const int nchunks = chunkOverride > 0 ? std::min(std::min(chunkOverride, 8), nframes) : (nframes >= 192 ? 4 : nframes >= 96 ? 3 : nframes >= 8 ? 2 : 1);
    const int per = (nframes + nchunks - 1) / nchunks;
    const int ncopy = std::min(capacity, P.kpCap);
    ORBB_CUDA(h, cudaEventRecord(h->evDone[0], h->stream));                  // earlier work on the handle's stream ...
    ORBB_CUDA(h, cudaStreamWaitEvent(h->h2dStream, h->evDone[0], 0));       // ... must finish before the staging area is reused
    mark(h, ST_H2D);
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * per, n = std::min(per, nframes - f0);
        if (n <= 0) break;
        const uint8_t* src = host_imgs + (size_t)f0 * frame_stride;
        if (frame_stride == (size_t)height * row_stride) {
            ORBB_CUDA(h, copy_rows(h->hImg + f0 * fbytes, width, src, row_stride, width, (size_t)height * n, cudaMemcpyHostToDevice, h->h2dStream));
        } else {
            for (int f = 0; f < n; f++)
                ORBB_CUDA(h, copy_rows(h->hImg + (f0 + f) * fbytes, width, src + (size_t)f * frame_stride, row_stride, width, height,
                                       cudaMemcpyHostToDevice, h->h2dStream));
        }
        ORBB_CUDA(h, cudaEventRecord(h->evH2D[c], h->h2dStream));
    }
    for (int c = 0; c < nchunks; c++) {
        const int f0 = c * per, n = std::min(per, nframes - f0);
        if (n <= 0) break;
        ORBB_CUDA(h, cudaStreamWaitEvent(h->stream, h->evH2D[c], 0));
        if ((rc = run_batch(h, h->hImg + f0 * fbytes, n, (size_t)width, fbytes, lap0, lap1, f0, true))) return rc;
        ORBB_CUDA(h, cudaEventRecord(h->evDone[c], h->stream));
        ORBB_CUDA(h, cudaStreamWaitEvent(h->d2hStream, h->evDone[c], 0));
        ORBB_CUDA(h, cudaMemcpyAsync(h->hCounts + 2 * f0, h->b.outCount + 2 * f0, sizeof(int) * 2 * n, cudaMemcpyDeviceToHost, h->d2hStream));
        ORBB_CUDA(h, cudaMemcpyAsync(h->hCounts + 2 * nframes + f0, h->b.status + f0, sizeof(int) * n, cudaMemcpyDeviceToHost, h->d2hStream));
        if (kps && ncopy > 0)
            ORBB_CUDA(h, copy_rows(kps + (size_t)f0 * capacity, sizeof(orbb_keypoint) * capacity, h->b.kps + (size_t)f0 * P.kpCap,
                                   sizeof(orbb_keypoint) * P.kpCap, sizeof(orbb_keypoint) * ncopy, n, cudaMemcpyDeviceToHost, h->d2hStream));
        if (desc && ncopy > 0)
            ORBB_CUDA(h, copy_rows(desc + (size_t)f0 * capacity * 32, (size_t)32 * capacity, h->b.desc + (size_t)f0 * P.kpCap * 32,
                                   (size_t)32 * P.kpCap, (size_t)32 * ncopy, n, cudaMemcpyDeviceToHost, h->d2hStream));
    }
    h->pendingFrames = nframes;
    h->pendingCapacity = (kps || desc) ? capacity : INT_MAX;
    return ORBB_OK;
}

int orbb_extract_batch_host_wait(orbb_extractor* h, int32_t* counts) {
    if (!h || !counts) return ORBB_ERR_ARG;
    if (h->pendingFrames <= 0) return set_err(h, ORBB_ERR_ARG, "no submitted batch to wait for");
    ORBB_CUDA(h, cudaSetDevice(h->device));
    ORBB_CUDA(h, cudaStreamSynchronize(h->d2hStream));
    ORBB_CUDA(h, cudaStreamSynchronize(h->stream));
    const int nframes = h->pendingFrames;
    h->pendingFrames = 0;
    for (int f = 0; f < nframes; f++) {
        counts[2 * f] = h->hCounts[2 * f];
        counts[2 * f + 1] = h->hCounts[2 * f + 1];
        if (h->hCounts[2 * nframes + f]) return set_err(h, ORBB_ERR_INTERNAL, "frame %d: device status 0x%x (capacity overflow)", f, h->hCounts[2 * nframes + f]);
        if (counts[2 * f] > h->pendingCapacity) return set_err(h, ORBB_ERR_CAPACITY, "frame %d has %d keypoints, capacity %d", f, counts[2 * f], h->pendingCapacity);
    }
    return ORBB_OK;
}

int orbb_extract_batch_host(orbb_extractor* h, const uint8_t* host_imgs, int nframes, int width, int height, size_t row_stride,
                            size_t frame_stride, int lap0, int lap1, orbb_keypoint* kps, uint8_t* desc, int capacity, int32_t* counts) {
    if (!h || !counts) return ORBB_ERR_ARG;
    const int rc = orbb_extract_batch_host_submit(h, host_imgs, nframes, width, height, row_stride, frame_stride, lap0, lap1, kps, desc, capacity);
    if (rc) return rc;
    return orbb_extract_batch_host_wait(h, counts);
}

int orbb_extract(orbb_extractor* h, const uint8_t* img, int width, int height, size_t stride, int lap0, int lap1,
                 orbb_keypoint* kps, uint8_t* desc, int capacity, int* n_out, int* mono_index) {
    if (!h) return ORBB_ERR_ARG;
    if (n_out) *n_out = 0;
    if (mono_index) *mono_index = 0;
    if (!img || width <= 0 || height <= 0) return set_err(h, ORBB_ERR_EMPTY, "empty image");     // :1090-1091
    int32_t counts[2] = {0, 0};
    int rc = orbb_extract_batch_host(h, img, 1, width, height, stride, stride * (size_t)height, lap0, lap1, kps, desc, capacity, counts);
    if (n_out) *n_out = counts[0];
    if (mono_index) *mono_index = counts[1];
    return rc;
}

// The 19-px reflect-101 apron around every level (ComputePyramid's copyMakeBorder, :1185-1191) of the frames of the last
// extraction.  Nothing in the pipeline reads it (the blur reflects at the image edge itself), so it is built here, when the pyramid is
// handed out (mvImagePyramid views, stage taps), instead of with every frame.
static int ensure_full_apron(orbb_extractor* h) {
    if (h->apronFull || h->lastFrames <= 0) return ORBB_OK;
    const Plan& P = h->plan;
    ApronTable T;
    for (int l = 0; l <= ORBB_MAX_LEVELS; l++) T.base[l] = l < P.nlevels ? P.apron[l].itemBase : P.apronItems;
    k_pyr_apron16<<<dim3((P.apronItems + 255) / 256, h->lastFrames), 256, 0, h->stream>>>(h->dPlan, h->b, T, P.nlevels);
    ORBB_CUDA(h, cudaGetLastError());
    h->launches++;
    h->apronFull = true;
    return ORBB_OK;
}

int orbb_pyramid_level(orbb_extractor* h, int level, const uint8_t** ptr, int* width, int* height, size_t* stride) {
    if (!h || !ptr) return ORBB_ERR_ARG;
    if (!h->planValid || h->lastFrames == 0) return set_err(h, ORBB_ERR_ARG, "no frame has been extracted yet");
    if (level < 0 || level >= h->plan.nlevels) return set_err(h, ORBB_ERR_ARG, "level %d out of range", level);
    ORBB_CUDA(h, cudaSetDevice(h->device));
    const Plan& P = h->plan;
    if (!h->hPyrFresh) {      // lazy D2H of frame 0's whole pyramid slab
        const int arc = ensure_full_apron(h);
        if (arc) return arc;
        if (h->hPyrBytes < P.pyrStride) {
            if (h->hPyr) cudaFreeHost(h->hPyr);
            h->hPyr = nullptr; h->hPyrBytes = 0;
            ORBB_CUDA(h, cudaMallocHost((void**)&h->hPyr, P.pyrStride));
            h->hPyrBytes = P.pyrStride;
        }
        ORBB_CUDA(h, cudaMemcpyAsync(h->hPyr, h->b.pyr, P.pyrStride, cudaMemcpyDeviceToHost, h->stream));
        ORBB_CUDA(h, cudaStreamSynchronize(h->stream));
        h->hPyrFresh = true;
    }
    const LevelPlan& L = P.lv[level];
    *ptr = h->hPyr + L.roiOff;
    if (width) *width = L.w;
    if (height) *height = L.h;
    if (stride) *stride = (size_t)L.pitch;
    return ORBB_OK;
}

// ---- stage taps ----------------------------------------------------------------------------------
int orbb_debug_level(orbb_extractor* h, int frame, int level, int blurred, int bordered, uint8_t* dst, size_t dst_cap, int* width, int* height) {
    if (!h || !h->planValid || frame < 0 || frame >= h->lastFrames || level < 0 || level >= h->plan.nlevels) return ORBB_ERR_ARG;
    ORBB_CUDA(h, cudaSetDevice(h->device));
    const Plan& P = h->plan;
    const LevelPlan& L = P.lv[level];
    int w = L.w, hh = L.h;
    const uint8_t* src;
    size_t pitch;
    if (blurred) { src = h->b.blur + (size_t)frame * P.blurStride + L.blurOff; pitch = L.bpitch; }
    else if (bordered) { const int arc = ensure_full_apron(h); if (arc) return arc; src = h->b.pyr + (size_t)frame * P.pyrStride + L.pyrOff + (kRoiX - kEdge); pitch = L.pitch; w += 2 * kEdge; hh += 2 * kEdge; }
    else { src = h->b.pyr + (size_t)frame * P.pyrStride + L.roiOff; pitch = L.pitch; }
    if (width) *width = w;
    if (height) *height = hh;
    if (!dst) return ORBB_OK;
    if (dst_cap < (size_t)w * hh) return set_err(h, ORBB_ERR_CAPACITY, "level buffer too small");
    ORBB_CUDA(h, cudaMemcpy2DAsync(dst, w, src, pitch, w, hh, cudaMemcpyDeviceToHost, h->stream));
    ORBB_CUDA(h, cudaStreamSynchronize(h->stream));
    return ORBB_OK;
}

int orbb_debug_raw_keys(orbb_extractor* h, int frame, int level, float* xyr, int cap, int* n) {
    if (!h || !h->planValid || frame < 0 || frame >= h->lastFrames || level < 0 || level >= h->plan.nlevels || !n) return ORBB_ERR_ARG;
    ORBB_CUDA(h, cudaSetDevice(h->device));
    const Plan& P = h->plan;
    const LevelPlan& L = P.lv[level];
    const int nCells = L.nCols * L.nRows;
    std::vector<int> cnt(nCells);
    std::vector<u64> keys((size_t)nCells * L.cellCap);
    ORBB_CUDA(h, cudaStreamSynchronize(h->stream));
    ORBB_CUDA(h, cudaMemcpy(cnt.data(), h->b.cellCount + (size_t)frame * P.cellsTotal + L.cellBase, sizeof(int) * nCells, cudaMemcpyDeviceToHost));
    ORBB_CUDA(h, cudaMemcpy(keys.data(), h->b.cellKeys + (size_t)frame * P.cellKeyStride + L.cellKeyBase, sizeof(u64) * keys.size(), cudaMemcpyDeviceToHost));
    int m = 0;
    for (int c = 0; c < nCells; c++)
        for (int i = 0; i < cnt[c]; i++, m++) {
            if (m < cap && xyr) {
                const u64 k = keys[(size_t)c * L.cellCap + i];
                xyr[3 * m] = (float)(k & 0xffff); xyr[3 * m + 1] = (float)((k >> 16) & 0xffff); xyr[3 * m + 2] = (float)(k >> 32);
            }
        }
    *n = m;
    return (xyr && m > cap) ? ORBB_ERR_CAPACITY : ORBB_OK;
}

int orbb_debug_selected(orbb_extractor* h, int frame, int level, float* xyr, int cap, int* n) {
    if (!h || !h->planValid || frame < 0 || frame >= h->lastFrames || level < 0 || level >= h->plan.nlevels || !n) return ORBB_ERR_ARG;
    ORBB_CUDA(h, cudaSetDevice(h->device));
    const Plan& P = h->plan;
    const LevelPlan& L = P.lv[level];
    int cnt = 0;
    ORBB_CUDA(h, cudaStreamSynchronize(h->stream));
    ORBB_CUDA(h, cudaMemcpy(&cnt, h->b.selCount + frame * ORBB_MAX_LEVELS + level, sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<u64> keys(std::max(cnt, 1));
    ORBB_CUDA(h, cudaMemcpy(keys.data(), h->b.sel + (size_t)frame * P.selStride + L.selBase, sizeof(u64) * cnt, cudaMemcpyDeviceToHost));
    *n = cnt;
    for (int i = 0; i < cnt && i < cap && xyr; i++) {
        xyr[3 * i] = (float)(keys[i] & 0xffff); xyr[3 * i + 1] = (float)((keys[i] >> 16) & 0xffff); xyr[3 * i + 2] = (float)(keys[i] >> 32);
    }
    return (xyr && cnt > cap) ? ORBB_ERR_CAPACITY : ORBB_OK;
}

void* orbb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void orbb_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- rectification ("next" row): cv::remap with the maps of cv::initUndistortRectifyMap, then extraction -------------
struct orbb_rectifier {
    int device = 0;
    int dw = 0, dh = 0, sw = 0, sh = 0;
    uint2* map = nullptr;                 // [dh][dw] {sx | sy << 16, a}
    uint8_t* dSrc = nullptr; size_t srcBytes = 0;      // staging of one host source frame
    uint8_t* dDst = nullptr; size_t dstBytes = 0;
    cudaStream_t stream = nullptr;
};

int orbb_rectifier_create(int device, const float* map_x, const float* map_y, size_t map_stride, int dst_width, int dst_height,
                          int src_width, int src_height, orbb_rectifier** out) {
    if (!out || !map_x || !map_y || dst_width <= 0 || dst_height <= 0 || src_width <= 0 || src_height <= 0 || map_stride < (size_t)dst_width)
        return set_err(nullptr, ORBB_ERR_ARG, "bad rectifier arguments");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return set_err(nullptr, ORBB_ERR_CUDA, "no CUDA device (%s): liborbb200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return set_err(nullptr, ORBB_ERR_ARG, "device %d out of range", device);
    // RemapInvoker's float -> fixed-point map conversion (imgwarp.cpp), done once
    std::vector<uint2> fixed((size_t)dst_width * dst_height);
    for (int y = 0; y < dst_height; y++)
        for (int x = 0; x < dst_width; x++) {
            const int sx = cv_round_f(map_x[(size_t)y * map_stride + x] * 32.f), sy = cv_round_f(map_y[(size_t)y * map_stride + x] * 32.f);
            const int ix = std::min(std::max(sx >> 5, -32768), 32767), iy = std::min(std::max(sy >> 5, -32768), 32767);      // saturate_cast<short>
            fixed[(size_t)y * dst_width + x] = make_uint2((unsigned)(ix & 0xffff) | ((unsigned)(iy & 0xffff) << 16), (unsigned)((sy & 31) * 32 + (sx & 31)));
        }
    orbb_rectifier* r = new orbb_rectifier();
    r->device = device; r->dw = dst_width; r->dh = dst_height; r->sw = src_width; r->sh = src_height;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void**)&r->map, fixed.size() * sizeof(uint2)) != cudaSuccess ||
        cudaMemcpy(r->map, fixed.data(), fixed.size() * sizeof(uint2), cudaMemcpyHostToDevice) != cudaSuccess) {
        const cudaError_t le = cudaGetLastError();
        orbb_rectifier_destroy(r);
        return set_err(nullptr, ORBB_ERR_CUDA, "rectifier setup failed: %s", cudaGetErrorString(le));
    }
    *out = r;
    return ORBB_OK;
}

void orbb_rectifier_destroy(orbb_rectifier* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    if (r->stream) { cudaStreamSynchronize(r->stream); cudaStreamDestroy(r->stream); }
    cudaFree(r->map); cudaFree(r->dSrc); cudaFree(r->dDst);
    delete r;
}

static void launch_remap(const orbb_rectifier* r, const uint8_t* dSrc, size_t srcStride, size_t srcFrameStride, uint8_t* dDst, size_t dstStride,
                         size_t dstFrameStride, int nframes, cudaStream_t st) {
    dim3 grid(((r->dw + 3) / 4 + 31) / 32, (r->dh + 7) / 8, nframes);
    k_remap<<<grid, 256, 0, st>>>(r->map, dSrc, srcStride, srcFrameStride, r->sw, r->sh, dDst, dstStride, dstFrameStride, r->dw, r->dh);
}

int orbb_remap(orbb_rectifier* r, const uint8_t* src, size_t src_stride, uint8_t* dst, size_t dst_stride) {
    if (!r || !src || !dst || src_stride < (size_t)r->sw || dst_stride < (size_t)r->dw) return set_err(nullptr, ORBB_ERR_ARG, "bad remap arguments");
    if (cudaSetDevice(r->device) != cudaSuccess) return set_err(nullptr, ORBB_ERR_CUDA, "cudaSetDevice failed");
    const size_t sb = (size_t)r->sw * r->sh, db = (size_t)r->dw * r->dh;
    if (r->srcBytes < sb) { cudaFree(r->dSrc); r->dSrc = nullptr; r->srcBytes = 0; if (cudaMalloc((void**)&r->dSrc, sb) != cudaSuccess) return set_err(nullptr, ORBB_ERR_CUDA, "out of device memory"); r->srcBytes = sb; }
    if (r->dstBytes < db) { cudaFree(r->dDst); r->dDst = nullptr; r->dstBytes = 0; if (cudaMalloc((void**)&r->dDst, db) != cudaSuccess) return set_err(nullptr, ORBB_ERR_CUDA, "out of device memory"); r->dstBytes = db; }
    cudaError_t e = cudaMemcpy2DAsync(r->dSrc, r->sw, src, src_stride, r->sw, r->sh, cudaMemcpyHostToDevice, r->stream);
    if (e == cudaSuccess) {
        launch_remap(r, r->dSrc, r->sw, sb, r->dDst, r->dw, db, 1, r->stream);
        e = cudaMemcpy2DAsync(dst, dst_stride, r->dDst, r->dw, r->dw, r->dh, cudaMemcpyDeviceToHost, r->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(r->stream);
    if (e != cudaSuccess) return set_err(nullptr, ORBB_ERR_CUDA, "remap failed: %s", cudaGetErrorString(e));
    return ORBB_OK;
}

int orbb_extract_batch_rectified(orbb_extractor* h, orbb_rectifier* r, const uint8_t* dev_imgs, int nframes, size_t row_stride,
                                 size_t frame_stride, int lap0, int lap1) {
    if (!h || !r) return ORBB_ERR_ARG;
    if (!dev_imgs || nframes <= 0) return set_err(h, ORBB_ERR_EMPTY, "empty image");
    if (r->device != h->device) return set_err(h, ORBB_ERR_ARG, "rectifier and extractor live on different devices");
    ORBB_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_plan(h, r->dw, r->dh, nframes);
    if (rc) return rc;
    if ((rc = ensure_staging(h, (size_t)nframes * r->dw * r->dh))) return rc;
    launch_remap(r, dev_imgs, row_stride, frame_stride, h->hImg, (size_t)r->dw, (size_t)r->dw * r->dh, nframes, h->stream);
    h->launches++;
    return run_batch(h, h->hImg, nframes, (size_t)r->dw, (size_t)r->dw * r->dh, lap0, lap1);
}

int orbb_extract_rectified(orbb_extractor* h, orbb_rectifier* r, const uint8_t* img, size_t stride, int lap0, int lap1, orbb_keypoint* kps,
                           uint8_t* desc, int capacity, int* n_out, int* mono_index) {
    if (!h || !r) return ORBB_ERR_ARG;
    if (n_out) *n_out = 0;
    if (mono_index) *mono_index = 0;
    if (!img) return set_err(h, ORBB_ERR_EMPTY, "empty image");
    ORBB_CUDA(h, cudaSetDevice(h->device));
    const size_t sb = (size_t)r->sw * r->sh;
    if (h->colorBytes < sb) {                              // (the raw-input staging buffer is shared with the colour path)
        if (h->dColor) cudaFree(h->dColor);
        h->dColor = nullptr; h->colorBytes = 0;
        ORBB_CUDA(h, cudaMalloc((void**)&h->dColor, sb));
        h->colorBytes = sb;
    }
    ORBB_CUDA(h, cudaMemcpy2DAsync(h->dColor, r->sw, img, stride, r->sw, r->sh, cudaMemcpyHostToDevice, h->stream));
    int rc = orbb_extract_batch_rectified(h, r, h->dColor, 1, (size_t)r->sw, sb, lap0, lap1);
    if (rc) return rc;
    int32_t counts[2] = {0, 0};
    rc = orbb_batch_fetch(h, 1, kps, desc, capacity, counts);
    if (n_out) *n_out = counts[0];
    if (mono_index) *mono_index = counts[1];
    return rc;
}

// ---- input-side resize ("next" row): cv::resize(im, imToFeed, newImSize) of System::Track* (System.cc:241-244), then extraction --
static int ensure_resize_tables(orbb_extractor* h, int sw, int sh, int dw, int dh) {
    if (h->rsTab && h->rsSw == sw && h->rsSh == sh && h->rsDw == dw && h->rsDh == dh) return ORBB_OK;
    std::vector<int2> tab;
    const int tabY = append_resize_tables(tab, sw, sh, dw, dh);
    if (h->rsTab) cudaFree(h->rsTab);
    h->rsTab = nullptr;
    ORBB_CUDA(h, cudaMalloc((void**)&h->rsTab, tab.size() * sizeof(int2)));
    ORBB_CUDA(h, cudaMemcpyAsync(h->rsTab, tab.data(), tab.size() * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    ORBB_CUDA(h, cudaStreamSynchronize(h->stream));          // `tab` is a local
    h->rsTabY = tabY; h->rsSw = sw; h->rsSh = sh; h->rsDw = dw; h->rsDh = dh;
    return ORBB_OK;
}

int orbb_extract_batch_resized(orbb_extractor* h, const uint8_t* dev_imgs, int nframes, int width, int height, size_t row_stride,
                               size_t frame_stride, int new_width, int new_height, int lap0, int lap1) {
    if (!h) return ORBB_ERR_ARG;
    if (!dev_imgs || nframes <= 0 || width <= 0 || height <= 0 || new_width <= 0 || new_height <= 0) return set_err(h, ORBB_ERR_EMPTY, "empty image");
    ORBB_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_plan(h, new_width, new_height, nframes);
    if (rc) return rc;
    if ((rc = ensure_staging(h, (size_t)nframes * new_width * new_height))) return rc;
    if ((rc = ensure_resize_tables(h, width, height, new_width, new_height))) return rc;
    dim3 grid(((new_width + 3) / 4 + 31) / 32, (new_height + 7) / 8, nframes);
    k_resize_image<<<grid, 256, 0, h->stream>>>(h->rsTab, h->rsTabY, dev_imgs, row_stride, frame_stride, width, height, h->hImg, new_width, new_height);
    h->launches++;
    return run_batch(h, h->hImg, nframes, (size_t)new_width, (size_t)new_width * new_height, lap0, lap1);
}

int orbb_extract_resized(orbb_extractor* h, const uint8_t* img, int width, int height, size_t stride, int new_width, int new_height, int lap0,
                         int lap1, orbb_keypoint* kps, uint8_t* desc, int capacity, int* n_out, int* mono_index) {
    if (!h) return ORBB_ERR_ARG;
    if (n_out) *n_out = 0;
    if (mono_index) *mono_index = 0;
    if (!img || width <= 0 || height <= 0) return set_err(h, ORBB_ERR_EMPTY, "empty image");
    ORBB_CUDA(h, cudaSetDevice(h->device));
    const size_t sb = (size_t)width * height;
    if (h->colorBytes < sb) {                              // (the raw-input staging buffer is shared with the colour / rectify paths)
        if (h->dColor) cudaFree(h->dColor);
        h->dColor = nullptr; h->colorBytes = 0;
        ORBB_CUDA(h, cudaMalloc((void**)&h->dColor, sb));
        h->colorBytes = sb;
    }
    ORBB_CUDA(h, cudaMemcpy2DAsync(h->dColor, width, img, stride, width, height, cudaMemcpyHostToDevice, h->stream));
    int rc = orbb_extract_batch_resized(h, h->dColor, 1, width, height, (size_t)width, sb, new_width, new_height, lap0, lap1);
    if (rc) return rc;
    int32_t counts[2] = {0, 0};
    rc = orbb_batch_fetch(h, 1, kps, desc, capacity, counts);
    if (n_out) *n_out = counts[0];
    if (mono_index) *mono_index = counts[1];
    return rc;
}

}  // extern "C"

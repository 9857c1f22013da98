// "Next" row of the scope table (input side): the stereo rectification that System::TrackStereo applies to both images
// before Tracking sees them -- cv::remap(src, dst, M1, M2, cv::INTER_LINEAR) with the CV_32FC1 maps of
// cv::initUndistortRectifyMap (reference orb_slam3/src/System.cc:233-240, orb_slam3/src/Settings.cc:506-509).
// (included INSIDE namespace orbb)
//
// OpenCV's arithmetic for 8UC1 / INTER_LINEAR / BORDER_CONSTANT(0) (imgproc/imgwarp.cpp: RemapInvoker + remapBilinear):
//   sx = cvRound(mapx * 32), sy = cvRound(mapy * 32);  integer source pixel (sx >> 5, sy >> 5) saturated to short,
//   fraction a = (sy & 31) * 32 + (sx & 31);  weights = BilinearTab_i[a] = {(32-fy)(32-fx), (32-fy)fx, fy(32-fx), fy*fx} * 32
//   as shorts -- except a == 0, where 32768 saturates to 32767 and the table's sum fix-up puts the missing 1 on the LAST
//   tap: {32767, 0, 0, 1};  dst = (sum(w_i * p_i) + 2^14) >> 15;  taps outside the source count as 0.
// The map is converted to (sx, sy, a) once, on the host, when the rectifier is created.
#pragma once

__device__ __forceinline__ unsigned remap_px(const uint8_t* __restrict__ src, size_t stride, int sw, int sh, uint2 m) {
    const int sx = (short)(m.x & 0xffffu), sy = (short)(m.x >> 16);
    const int a = (int)m.y;
    const int fx = a & 31, fy = a >> 5;
    int w0 = (32 - fy) * (32 - fx) * 32, w1 = (32 - fy) * fx * 32, w2 = fy * (32 - fx) * 32, w3 = fy * fx * 32;
    if (a == 0) { w0 = 32767; w3 = 1; }
    int p0, p1, p2, p3;
    if ((unsigned)sx < (unsigned)(sw - 1) && (unsigned)sy < (unsigned)(sh - 1)) {
        const uint8_t* s = src + (size_t)sy * stride + sx;
        p0 = __ldg(s); p1 = __ldg(s + 1); p2 = __ldg(s + stride); p3 = __ldg(s + stride + 1);
    } else {
        if (sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0) return 0;
        const bool x0 = sx >= 0, x1 = sx + 1 < sw, y0 = sy >= 0, y1 = sy + 1 < sh;      // (sx < sw, sy < sh, sx+1 >= 0, sy+1 >= 0 hold here)
        const uint8_t* s = src + (ptrdiff_t)sy * (ptrdiff_t)stride + sx;
        p0 = (x0 && y0) ? __ldg(s) : 0;
        p1 = (x1 && y0) ? __ldg(s + 1) : 0;
        p2 = (x0 && y1) ? __ldg(s + stride) : 0;
        p3 = (x1 && y1) ? __ldg(s + stride + 1) : 0;
    }
    return (unsigned)((p0 * w0 + p1 * w1 + p2 * w2 + p3 * w3 + (1 << 14)) >> 15);
}

// 4 destination pixels per thread; map entries are 8 bytes per pixel {sx | sy << 16, a}
__global__ void __launch_bounds__(256) k_remap(const uint2* __restrict__ map, const uint8_t* __restrict__ src, size_t srcStride,
                                               size_t srcFrameStride, int sw, int sh, uint8_t* __restrict__ dst, size_t dstStride,
                                               size_t dstFrameStride, int dw, int dh) {
    const int word = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int frame = blockIdx.z;
    if (word * 4 >= dw || y >= dh) return;
    const uint8_t* s = src + (size_t)frame * srcFrameStride;
    uint8_t* d = dst + (size_t)frame * dstFrameStride + (size_t)y * dstStride + word * 4;
    const uint2* m = map + (size_t)y * dw + word * 4;
    const int n = min(4, dw - word * 4);
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k < n) out |= remap_px(s, srcStride, sw, sh, __ldg(m + k)) << (8 * k);
    if (n == 4 && (dstStride & 3) == 0 && (dstFrameStride & 3) == 0) *reinterpret_cast<unsigned*>(d) = out;
    else
        for (int k = 0; k < n; k++) d[k] = (uint8_t)(out >> (8 * k));
}

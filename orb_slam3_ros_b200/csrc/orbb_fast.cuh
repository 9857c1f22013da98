// K2 (production version): per-cell FAST-9/16 + NMS + iniTh->minTh fallback, restructured so that the work is
// proportional to what survives each test instead of paying the full ring test in every warp:
//
//   A. quick reject for ALL interior pixels, 4 pixels per thread with byte-SIMD: a 9-long arc always contains ring
//      pixel 0 or 8 and ring pixel 4 or 12 (cv::FAST's high-speed test), so a pixel can only be a corner if
//      min(max(|v-p0|,|v-p8|), max(|v-p4|,|v-p12|)) > t.  VABSDIFF4 + VIMNMX.U16x2, survivors -> candidate list.
//   B. candidates (typically 10-20 % of the pixels): 16-pixel bright/dark masks + 9-arc test -> corner list.
//   C. corners (2-4 %): exact score  max over arcs of the arc minimum, minus 1  (3-input min network) -> score map.
//   D. NMS on the corners only (strict '>' against the 8 neighbours, non-corners count as 0).
//   E. survivors are few; each finds its raster rank by counting (the reference's output order, fast.cpp row scan).
//
// Included by orbb_extract.cu (needs Plan/Bufs and warp_incl_scan).  Reference: ORBextractor.cc:805-872, cv::FAST.
#pragma once
// (included INSIDE namespace orbb)

constexpr int FAST_THREADS = 256;
constexpr int FAST_LIST = 74 * 74;      // max interior pixels of one cell

__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}
__device__ __forceinline__ unsigned umax16x2(unsigned a, unsigned b) { return __vmaxu2(a, b); }
__device__ __forceinline__ unsigned umin16x2(unsigned a, unsigned b) { return __vminu2(a, b); }

__device__ __forceinline__ bool arc9(unsigned m16) {
    const unsigned m = m16 | (m16 << 16);
    unsigned t = m & (m >> 1);
    t &= t >> 2;
    t &= t >> 4;
    t &= m >> 8;
    return (t & 0xffffu) != 0;
}

__device__ __forceinline__ int min3i(int a, int b, int c) { return min(min(a, b), c); }
__device__ __forceinline__ int max3i(int a, int b, int c) { return max(max(a, b), c); }

// mode 0: both thresholds (iniTh, then minTh if the cell stays empty).  mode 1: only cells that k_fast_cells marked -1
// (no corner at iniTh), minTh only.
__global__ void __launch_bounds__(FAST_THREADS) k_fast(const Plan* __restrict__ P, Bufs B, int mode) {
    constexpr int PS = kCellPix;
    __shared__ __align__(16) uint8_t sPix[PS * PS];
    __shared__ __align__(16) uint8_t sScore[PS * PS];
    __shared__ unsigned short sCand[FAST_LIST];      // phase A -> B; reused for the survivors (as 32-bit) in D/E
    __shared__ unsigned short sCorner[FAST_LIST];    // phase B -> C/D: position | polarity << 15
    __shared__ int sCnt[3];

    const int frame = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    // mode 1: persistent CTAs walk the frame's list of cells that stayed empty at iniThFAST
    const int nWork = mode == 1 ? B.fbCount[frame] : P->cellsTotal;
    for (int work = blockIdx.x; work < nWork; work += gridDim.x) {
    const int gcell = mode == 1 ? B.fbList[(size_t)frame * P->cellsTotal + work] : work;
    const CellDesc cd = B.cellDesc[gcell];
    const LevelPlan& L = P->lv[cd.level];
    int* cellCount = B.cellCount + (size_t)frame * P->cellsTotal + gcell;
    const int iniX = cd.gx0 - 3, iniY = cd.gy0 - 3;
    const int rw = cd.gx1 - cd.gx0 + 6, rh = cd.gy1 - cd.gy0 + 6;
    if (cd.gx1 <= cd.gx0) {                        // cell skipped by the reference (:810,:819) or smaller than 7 px
        if (tid == 0) *cellCount = 0;
        continue;
    }
    __syncthreads();                               // previous work item is done with the shared tiles
    // ---- stage the ROI with aligned 32-bit loads: tile column 0 = level column (iniX & ~3) ----
    const int sh = iniX & 3;                       // ROI column x lives at tile column x + sh
    const int nw = (rw + sh + 3) >> 2;             // words per tile row (<= 20)
    const unsigned mw = (65536u + nw - 1) / nw;    // i / nw == (i * mw) >> 16 for i < 3276 (rh * nw <= 1600)
    {
        const uint8_t* g = B.pyr + (size_t)frame * P->pyrStride + L.roiOff + (size_t)iniY * L.pitch + (iniX - sh);
        for (int i = tid; i < rh * nw; i += FAST_THREADS) {
            const int r = (int)(((unsigned)i * mw) >> 16), w = i - r * nw;
            reinterpret_cast<unsigned*>(sPix)[r * (PS / 4) + w] = __ldg(reinterpret_cast<const unsigned*>(g + (size_t)r * L.pitch) + w);
        }
    }
    const int ih = rh - 6;
    const int x0 = 3 + sh, x1 = rw - 3 + sh;       // interior tile columns [x0, x1)
    const int items = ih * nw;
    u64* out = B.cellKeys + (size_t)frame * P->cellKeyStride + cd.outOff;
    const int kx = iniX - kMinBorder - sh, ky = iniY - kMinBorder;          // :865-866 (tile column -> ROI column)
    unsigned* sSurv = reinterpret_cast<unsigned*>(sCand);
    int total = 0;

    for (int pass = mode; pass < 2 && total == 0; pass++) {
        const int th = min(max(pass == 0 ? P->iniTh : P->minTh, 0), 255);
        __syncthreads();
        if (tid < 3) sCnt[tid] = 0;
        for (int i = tid; i < rh * (PS / 16); i += FAST_THREADS) reinterpret_cast<uint4*>(sScore)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();

        // ---- A: quick reject, 4 pixels per thread ----
        const unsigned K = (unsigned)(0x7fff - th) * 0x00010001u;
        for (int base = 0; base < items; base += FAST_THREADS) {
            const int i = base + tid;
            unsigned cand = 0;
            int pos0 = 0;
            if (i < items) {
                const int r = (int)(((unsigned)i * mw) >> 16), w = i - r * nw;
                const int y = r + 3;
                const unsigned* row = reinterpret_cast<const unsigned*>(sPix) + y * (PS / 4) + w;
                const unsigned wc = row[0];
                const unsigned wl = w > 0 ? row[-1] : 0u, wr = row[1];
                const unsigned wt = row[3 * (PS / 4)], wb = row[-3 * (PS / 4)];
                const unsigned pl = __funnelshift_r(wl, wc, 8);     // bytes x-3 .. x
                const unsigned pr = __funnelshift_r(wc, wr, 24);    // bytes x+3 .. x+6
                const unsigned a0 = __vabsdiffu4(wc, wt), a8 = __vabsdiffu4(wc, wb);
                const unsigned a4 = __vabsdiffu4(wc, pr), a12 = __vabsdiffu4(wc, pl);
                const unsigned me = umin16x2(umax16x2(prmt(a0, 0, 0x4240), prmt(a8, 0, 0x4240)),
                                             umax16x2(prmt(a4, 0, 0x4240), prmt(a12, 0, 0x4240)));
                const unsigned mo = umin16x2(umax16x2(prmt(a0, 0, 0x4341), prmt(a8, 0, 0x4341)),
                                             umax16x2(prmt(a4, 0, 0x4341), prmt(a12, 0, 0x4341)));
                const unsigned te = me + K, to = mo + K;            // bit 15 / 31 set  <=>  lane value > th
                cand = ((te >> 15) & 1u) | ((to >> 14) & 2u) | ((te >> 29) & 4u) | ((to >> 28) & 8u);
                const int xb = 4 * w;
                const int lo = max(x0 - xb, 0), hi = min(x1 - xb, 4);
                cand &= hi > lo ? (((1u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
                pos0 = y * PS + xb;
            }
            const int cnt = __popc(cand);
            const int inc = warp_incl_scan(cnt, lane);
            const int wtot = __shfl_sync(0xffffffffu, inc, 31);
            int wbase = 0;
            if (lane == 31 && wtot) wbase = atomicAdd(&sCnt[0], wtot);
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
            int o = wbase + inc - cnt;
            while (cand) {
                const int k = __ffs(cand) - 1;
                cand &= cand - 1;
                sCand[o++] = (unsigned short)(pos0 + k);
            }
        }
        __syncthreads();

        // ---- B: full 16-pixel ring test on the candidates ----
        const int nCand = sCnt[0];
        for (int base = 0; base < nCand; base += FAST_THREADS) {
            const int i = base + tid;
            bool corner = false;
            unsigned rec = 0;
            if (i < nCand) {
                const int pos = sCand[i];
                const uint8_t* q = &sPix[pos];
                const int v = q[0], hi = v + th, lo = v - th;
                unsigned mb = 0, md = 0;        // ring pixel darker than v-th ("bright centre") / brighter than v+th
#define ORBB_RING(off)                                                     \
    {                                                                      \
        const int p = q[off];                                              \
        mb = __funnelshift_l((unsigned)(p - lo), mb, 1);  /* p < lo */     \
        md = __funnelshift_l((unsigned)(hi - p), md, 1);  /* p > hi */     \
    }
                ORBB_RING(3 * PS) ORBB_RING(3 * PS + 1) ORBB_RING(2 * PS + 2) ORBB_RING(PS + 3)
                ORBB_RING(3) ORBB_RING(-PS + 3) ORBB_RING(-2 * PS + 2) ORBB_RING(-3 * PS + 1)
                ORBB_RING(-3 * PS) ORBB_RING(-3 * PS - 1) ORBB_RING(-2 * PS - 2) ORBB_RING(-PS - 3)
                ORBB_RING(-3) ORBB_RING(PS - 3) ORBB_RING(2 * PS - 2) ORBB_RING(3 * PS - 1)
#undef ORBB_RING
                const bool cb = arc9(mb & 0xffffu), cd = arc9(md & 0xffffu);
                corner = cb | cd;
                rec = (unsigned)pos | (cd ? 0x8000u : 0u);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, corner);
            int wbase = 0;
            if (lane == 0 && bal) wbase = atomicAdd(&sCnt[1], __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (corner) sCorner[wbase + __popc(bal & ((1u << lane) - 1))] = (unsigned short)rec;
        }
        __syncthreads();

        // ---- C: exact score of the corners ----
        const int nCorner = sCnt[1];
        for (int i = tid; i < nCorner; i += FAST_THREADS) {
            const unsigned rec = sCorner[i];
            const int pos = rec & 0x7fff;
            const uint8_t* q = &sPix[pos];
            const int v = q[0];
            const int sgn = (rec & 0x8000u) ? -1 : 1;
            int d[16];
            d[0] = sgn * (v - q[3 * PS]);       d[1] = sgn * (v - q[3 * PS + 1]);   d[2] = sgn * (v - q[2 * PS + 2]);
            d[3] = sgn * (v - q[PS + 3]);       d[4] = sgn * (v - q[3]);            d[5] = sgn * (v - q[-PS + 3]);
            d[6] = sgn * (v - q[-2 * PS + 2]);  d[7] = sgn * (v - q[-3 * PS + 1]);  d[8] = sgn * (v - q[-3 * PS]);
            d[9] = sgn * (v - q[-3 * PS - 1]);  d[10] = sgn * (v - q[-2 * PS - 2]); d[11] = sgn * (v - q[-PS - 3]);
            d[12] = sgn * (v - q[-3]);          d[13] = sgn * (v - q[PS - 3]);      d[14] = sgn * (v - q[2 * PS - 2]);
            d[15] = sgn * (v - q[3 * PS - 1]);
            int m3[16];
#pragma unroll
            for (int k = 0; k < 16; k++) m3[k] = min3i(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
            int M = -256;
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
                const int e0 = min3i(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
                const int e1 = min3i(m3[k + 1], m3[(k + 4) & 15], m3[(k + 7) & 15]);
                M = max3i(M, e0, e1);
            }
            sScore[pos] = (uint8_t)(M - 1);          // response = M - 1 (M > th >= 0)
        }
        __syncthreads();

        // ---- D: non-maximum suppression on the corners ----
        for (int base = 0; base < nCorner; base += FAST_THREADS) {
            const int i = base + tid;
            bool keep = false;
            unsigned rec = 0;
            if (i < nCorner) {
                const int pos = sCorner[i] & 0x7fff;
                const uint8_t* q = &sScore[pos];
                const int s = q[0];
                keep = s > 0 && s > q[-1] && s > q[1] && s > q[-PS - 1] && s > q[-PS] && s > q[-PS + 1] && s > q[PS - 1] &&
                       s > q[PS] && s > q[PS + 1];
                rec = ((unsigned)pos << 8) | (unsigned)s;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            int wbase = 0;
            if (lane == 0 && bal) wbase = atomicAdd(&sCnt[2], __popc(bal));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (keep) sSurv[wbase + __popc(bal & ((1u << lane) - 1))] = rec;     // sCand is dead after phase B
        }
        __syncthreads();

        // ---- E: raster order by rank counting (pos = y*PS + x is the raster key) ----
        const int nSurv = sCnt[2];
        for (int i = tid; i < nSurv; i += FAST_THREADS) {
            const unsigned rec = sSurv[i];
            int rank = 0;
            for (int j = 0; j < nSurv; j++) rank += sSurv[j] < rec;
            const int pos = (int)(rec >> 8);
            const int y = pos / PS, x = pos - y * PS;
            out[rank] = (u64)(unsigned)(x + kx) | ((u64)(unsigned)(y + ky) << 16) | ((u64)(rec & 0xffu) << 32);
        }
        total = nSurv;
    }
    if (tid == 0) *cellCount = total;
    }
}


// K2, fused formulations.  (included INSIDE namespace orbb, after orbb_fast.cuh / orbb_fast2.cuh)
// Production: k_fast_cell (one warp = one CTA = one cell, at the end of this file).  k_fast_band (ORBB_FAST_MODE=band) is
// the same per-cell state machine with the tile shared by up to 4 cells of a cell row:
//
// One CTA stages a BAND SEGMENT -- a run of up to 4 adjacent 35-px cells of one cell row of one level -- and then each
// WARP finishes one cell on its own (reference ORBextractor.cc:805-872: cv::FAST on each cell at iniThFAST, again at
// minThFAST when the cell stays empty).  The cell interiors tile the level and the non-maximum suppression of cv::FAST never
// looks across a cell border, so after the tile is in shared memory there is no CTA-wide barrier any more:
//
//   stage   the segment's pixels (+3 px ring margin) by one TMA bulk copy per row (cp.async.bulk + mbarrier)
//   per warp / cell, as a small state machine over two stacks in shared memory, so that the expensive steps always run
//   with 32 busy lanes:
//     A     quick reject, 4 px x 2 rows per lane, byte-SIMD (see orbb_fast.cuh)  -> candidate stack
//     B     pops 32 candidates: 16-pixel ring masks + 9-arc test                 -> corner stack
//     C     pops 32 corners: exact score (max over arcs of the arc minimum, -1)  -> score byte map + the cell's corner list
//     D     NMS over the cell's corner list (strict '>' against the 8 neighbours inside the cell) -> survivors
//     retry a cell without survivors runs A-D again at minThFAST on the pixels that are still in shared memory
//     E     raster order by rank counting -> cellKeys / cellCount  (= vToDistributeKeys, in order)
//
// Compared with the split formulation (k_fast_score + k_fast_cells + k_fast mode 1) the score map never goes to HBM, the
// retry does not reload the cell, and ring test / score never run with a partly filled warp except once per cell.
#pragma once

constexpr int FB_WARPS = 4;                // = max cells per segment
constexpr int FB_THREADS = 32 * FB_WARPS;
constexpr int FB_TP = 192;                 // tile pitch (bytes): 16-byte aligned window of <= 156 interior px + margins
constexpr int FB_MAXW = FB_TP - 36;        // max interior width of a segment (+6 margin +30 alignment slop)
constexpr int FB_MAXCELLS = FB_WARPS;
constexpr int FB_CANDS = 32 + 256;         // candidate stack: < 32 left over + one round of phase A (32 lanes x 8 px)
constexpr int FB_CORNS = 64;               // corner stack: < 32 left over + one round of phase B
constexpr int FB_NC = 128;                 // corners of one cell kept for the list-driven NMS (more: NMS scans the score map)
constexpr int FB_WARP_SMEM = 2 * (FB_CANDS + FB_CORNS + FB_NC);      // bytes of stacks per warp

// Quick reject of 4 pixels (one word) as a byte mask: bit 7 of byte k is set when pixel k can still be a corner, i.e.
// (|v-p0| > th or |v-p8| > th) and (|v-p4| > th or |v-p12| > th).  Per-byte unsigned "x > th" without widening:
// s = (x & 0x7f) + (0x7f - (th & 0x7f)) carries into bit 7 iff low7(x) > low7(th); for th < 128 the answer is s | x, for
// th >= 128 it is s & x (bit 7).  K7 = (0x7f - (th & 0x7f)) * 0x01010101, M = th >= 128 ? ~0u : 0.
__device__ __forceinline__ unsigned gt_bytes(unsigned x, unsigned K7, unsigned M) {
    const unsigned s = (x & 0x7f7f7f7fu) + K7;
    return (M & s & x) | (~M & (s | x));
}
__device__ __forceinline__ unsigned quick_bytes4(unsigned wl, unsigned wc, unsigned wr, unsigned wt, unsigned wb, unsigned K7, unsigned M) {
    const unsigned pl = __funnelshift_r(wl, wc, 8);      // bytes x-3 .. x
    const unsigned pr = __funnelshift_r(wc, wr, 24);     // bytes x+3 .. x+6
    const unsigned v = gt_bytes(__vabsdiffu4(wc, wt), K7, M) | gt_bytes(__vabsdiffu4(wc, wb), K7, M);
    const unsigned h = gt_bytes(__vabsdiffu4(wc, pr), K7, M) | gt_bytes(__vabsdiffu4(wc, pl), K7, M);
    return v & h;
}

// One cell, one warp.  tile: row t = level row gy0 - 3 + t, column = level column - X0.  score: row s = interior row s - 1.
// Returns the number of NMS survivors parked in `park`.
// The score map has its own pitch SP and column origin: score column = tile column - sxo (k_fast_cell keeps only the cell's
// columns plus a zero column on each side, which is what lets 32 cells fit an SM).
template <int TP, int SP>
__device__ __forceinline__ int cell_pass(const uint8_t* __restrict__ tile, uint8_t* __restrict__ score, unsigned short* candS,
                                         unsigned short* cornS, unsigned short* allS, unsigned* __restrict__ park, int cx0, int cx1,
                                         int ih, int th, int lane, int sxo) {
    constexpr int PS = TP, TPW = TP / 4;
    const int wa = cx0 >> 2, nwc = max(((cx1 + 3) >> 2) - wa, 2);        // (>= 2 keeps the reciprocal in 32 bits; extra words are masked)
    const unsigned mInv = 0xffffffffu / (unsigned)nwc + 1u;
    const int items = ((ih + 1) >> 1) * nwc;
    const unsigned K7 = (unsigned)(0x7f - (th & 0x7f)) * 0x01010101u, M = th >= 128 ? 0xffffffffu : 0u;
    int nCand = 0, nCorn = 0, nAll = 0, base = 0;
    for (;;) {
        const bool aDone = base >= items;
        if (nCorn >= 32 || (aDone && nCand == 0 && nCorn > 0)) {
            // ---- C: exact score of up to 32 corners ----
            const int n = min(nCorn, 32);
            nCorn -= n;
            if (lane < n) {
                const unsigned rec = cornS[nCorn + lane];
                const int pos = rec & 0x7fff;
                const uint8_t* q = tile + pos;
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
                const int t = pos / TP;                              // tile row = interior row + 3, score row = interior row + 1
                const int sp = (t - 2) * SP + (pos - t * TP) - sxo;
                score[sp] = (uint8_t)(M - 1);
                if (nAll + lane < FB_NC) allS[nAll + lane] = (unsigned short)sp;
            }
            nAll += n;
            __syncwarp();
            continue;
        }
        if (nCand >= 32 || (aDone && nCand > 0)) {
            // ---- B: ring test of up to 32 candidates ----
            const int n = min(nCand, 32);
            nCand -= n;
            bool isCorner = false;
            unsigned rec = 0;
            if (lane < n) {
                const unsigned pos = candS[nCand + lane];
                const uint8_t* q = tile + pos;
                const int v = q[0], hi = v + th, lo = v - th;
                unsigned mb = 0, md = 0;
#define ORBB_RING(off)                                       \
    {                                                        \
        const int p = q[off];                                \
        mb = __funnelshift_l((unsigned)(p - lo), mb, 1);     \
        md = __funnelshift_l((unsigned)(hi - p), md, 1);     \
    }
                ORBB_RING(3 * PS) ORBB_RING(3 * PS + 1) ORBB_RING(2 * PS + 2) ORBB_RING(PS + 3)
                ORBB_RING(3) ORBB_RING(-PS + 3) ORBB_RING(-2 * PS + 2) ORBB_RING(-3 * PS + 1)
                ORBB_RING(-3 * PS) ORBB_RING(-3 * PS - 1) ORBB_RING(-2 * PS - 2) ORBB_RING(-PS - 3)
                ORBB_RING(-3) ORBB_RING(PS - 3) ORBB_RING(2 * PS - 2) ORBB_RING(3 * PS - 1)
#undef ORBB_RING
                const bool cb = arc9(mb & 0xffffu), cd = arc9(md & 0xffffu);
                isCorner = cb | cd;
                rec = pos | (cd ? 0x8000u : 0u);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, isCorner);
            if (isCorner) cornS[nCorn + __popc(bal & ((1u << lane) - 1))] = (unsigned short)rec;
            nCorn += __popc(bal);
            __syncwarp();
            continue;
        }
        if (aDone) break;
        // ---- A: quick reject, one word column x 2 rows per lane ----
        {
            const int i = base + lane;
            unsigned cand = 0;
            int pos0 = 0;
            if (i < items) {
                const int s = (int)__umulhi((unsigned)i, mInv);
                const int w = wa + (i - s * nwc);
                const int r0 = 2 * s;
                const int xb0 = 4 * w;
                const int lo = max(cx0 - xb0, 0), hi = max(min(cx1 - xb0, 4), 0);
                const unsigned xm1 = hi > lo ? ((0x01010101u >> (8 * (4 - hi))) >> (8 * lo)) << (8 * lo) : 0u;      // 0x01 in the bytes of columns [cx0, cx1)
                const unsigned* q = reinterpret_cast<const unsigned*>(tile) + (r0 + 3) * TPW + w;
                cand = (quick_bytes4(q[-1], q[0], q[1], q[3 * TPW], q[-3 * TPW], K7, M) >> 7) & xm1;      // bit 8k: pixel k of row r0
                if (r0 + 1 < ih)                                                                              // bit 8k + 1: row r0 + 1
                    cand |= ((quick_bytes4(q[TPW - 1], q[TPW], q[TPW + 1], q[4 * TPW], q[-2 * TPW], K7, M) >> 7) & xm1) << 1;
                pos0 = (r0 + 3) * TP + xb0;
            }
            const int cnt = __popc(cand);
            const int inc = warp_incl_scan(cnt, lane);
            int o = nCand + inc - cnt;
            nCand += __shfl_sync(0xffffffffu, inc, 31);
            while (cand) {
                const int b = __ffs(cand) - 1;
                cand &= cand - 1;
                candS[o++] = (unsigned short)(pos0 + (b & 1) * TP + (b >> 3));
            }
            base += 32;
            __syncwarp();
        }
    }
    // ---- D: NMS inside the cell (strict '>' against the 8 neighbours; outside the cell counts as 0) ----
    int nSurv = 0;
    auto nms = [&](bool valid, int sp, int sc) {                      // sp = score-map position
        bool keep = false;
        unsigned rec = 0;
        if (valid) {
            const int r1 = sp / SP, x = sp - r1 * SP + sxo;
            const uint8_t* q = score + sp;
            int m = max((int)q[-SP], (int)q[SP]);
            if (x > cx0) m = max(m, max3i((int)q[-SP - 1], (int)q[-1], (int)q[SP - 1]));
            if (x + 1 < cx1) m = max(m, max3i((int)q[-SP + 1], (int)q[1], (int)q[SP + 1]));
            keep = sc > m;
            rec = ((unsigned)(r1 - 1) << 16) | ((unsigned)(x - cx0) << 8) | (unsigned)sc;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) park[nSurv + __popc(bal & ((1u << lane) - 1))] = rec;
        nSurv += __popc(bal);
    };
    if (nAll <= FB_NC) {
        for (int b0 = 0; b0 < nAll; b0 += 32) {
            const bool valid = b0 + lane < nAll;
            const int sp = valid ? allS[b0 + lane] : 0;
            nms(valid, sp, valid ? score[sp] : 0);
        }
    } else {                                                          // very dense cell: walk its score map
        const int wc = cx1 - cx0;
        for (int b0 = 0; b0 < ih * wc; b0 += 32) {
            const int i = b0 + lane;
            const int r = i / wc, x = cx0 + i - r * wc;
            const int sp = (r + 1) * SP + x - sxo;
            const int sc = i < ih * wc ? score[sp] : 0;
            nms(sc > 0, sp, sc);
        }
    }
    return nSurv;
}

__global__ void __launch_bounds__(FB_THREADS, 10) k_fast_band(const Plan* __restrict__ P, Bufs B, const BandDesc* __restrict__ bands) {
    extern __shared__ __align__(128) uint8_t fbSmem[];
    __shared__ __align__(8) unsigned long long sBar;
    const BandDesc bd = bands[blockIdx.x];
    const int frame = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const LevelPlan& L = P->lv[bd.level];
    const int ncell = bd.c1 - bd.c0;
    const int cellRow0 = bd.ci * L.nCols + bd.c0;                 // first cell of the segment within the level
    int* cellCount = B.cellCount + (size_t)frame * P->cellsTotal + L.cellBase + cellRow0;
    const int ih = bd.ih;
    if (ih <= 0) {                                                // cell row skipped by the reference (:810)
        if (tid < ncell) cellCount[tid] = 0;
        return;
    }
    const int rowsT = ih + 6;
    uint8_t* tile = fbSmem;
    uint8_t* score = fbSmem + ((rowsT * FB_TP + 127) & ~127);
    unsigned short* stacks = reinterpret_cast<unsigned short*>(score + (((ih + 2) * FB_TP + 127) & ~127)) + warp * (FB_WARP_SMEM / 2);
    // ---- stage: one bulk copy per tile row; the score map is cleared while the copies are in flight ----
    const uint8_t* roi = B.pyr + (size_t)frame * P->pyrStride + L.roiOff;
    if (tid == 0) {
        mbar_init(&sBar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sBar, rowsT * FB_TP);
    }
    __syncthreads();
    if (tid < rowsT) tma_bulk_g2s(tile + tid * FB_TP, roi + (ptrdiff_t)(bd.gy0 - 3 + tid) * L.pitch + bd.X0, FB_TP, &sBar);
    for (int i = tid; i < (ih + 2) * (FB_TP / 16); i += FB_THREADS) reinterpret_cast<uint4*>(score)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();                                              // score map cleared
    if (warp >= ncell) return;
    mbar_wait(&sBar, 0);

    // this warp's cell: interior columns [19 + j*wCell, min(19 + (j+1)*wCell, w - 19)) of the level, as tile columns
    const int j = bd.c0 + warp;
    const int cx0 = kEdge + j * L.wCell - bd.X0;
    const int cx1 = min(cx0 + L.wCell, L.w - kEdge - bd.X0);
    int nSurv = 0;
    // survivors are parked (unordered) in the quadtree's second key buffer, which k_octree only uses later
    unsigned* park = reinterpret_cast<unsigned*>(B.keys + ((size_t)frame * 2 + 1) * P->rawStride + L.rawBase + (size_t)(cellRow0 + warp) * L.cellCap);
    if (cx1 > cx0) {
        unsigned short* candS = stacks;
        unsigned short* cornS = candS + FB_CANDS;
        unsigned short* allS = cornS + FB_CORNS;
        nSurv = cell_pass<FB_TP, FB_TP>(tile, score, candS, cornS, allS, park, cx0, cx1, ih, min(max(P->iniTh, 0), 255), lane, 0);
        if (nSurv == 0)                                           // :833-846 (scores do not depend on the threshold: the map stays valid)
            nSurv = cell_pass<FB_TP, FB_TP>(tile, score, candS, cornS, allS, park, cx0, cx1, ih, min(max(P->minTh, 0), 255), lane, 0);
    }
    // ---- E: raster order by rank counting; keys are relative to the 16-px border (:865-866) ----
    __syncwarp();
    u64* keysOut = B.cellKeys + (size_t)frame * P->cellKeyStride + L.cellKeyBase + (size_t)(cellRow0 + warp) * L.cellCap;
    for (int i = lane; i < nSurv; i += 32) {
        const unsigned rec = __ldcg(park + i);
        int rank = 0;
        for (int k = 0; k < nSurv; k++) rank += __ldcg(park + k) < rec;
        const int x = kEdge + j * L.wCell + (int)((rec >> 8) & 0xffu) - kMinBorder;
        const int y = bd.gy0 + (int)(rec >> 16) - kMinBorder;
        keysOut[rank] = (u64)(unsigned)x | ((u64)(unsigned)y << 16) | ((u64)(rec & 0xffu) << 32);
    }
    if (lane == 0) cellCount[warp] = nSurv;
}


// One warp = one CTA = one cell with its own tile (pitch TP: 16-byte aligned window of the cell + its 3-px margin), so a
// finished cell frees its resources at once and up to 32 cells are in flight per SM, none waiting for a slower neighbour
// (in k_fast_band the CTA lives as long as its slowest cell, e.g. one that needs the minThFAST retry).
template <int TP>
__global__ void __launch_bounds__(32, 32) k_fast_cell(const Plan* __restrict__ P, Bufs B) {
    constexpr int SP = TP == 64 ? 48 : 80;                    // score pitch: cell width (<= 43 / 71) + a zero column on each side
    extern __shared__ __align__(128) uint8_t fcSmem[];
    __shared__ __align__(8) unsigned long long sBar;
    const int gcell = blockIdx.x, frame = blockIdx.y, lane = threadIdx.x;
    const CellDesc cd = B.cellDesc[gcell];
    int* cellCount = B.cellCount + (size_t)frame * P->cellsTotal + gcell;
    const int ih = cd.gy1 - cd.gy0;
    if (cd.gx1 <= cd.gx0 || ih <= 0) {                        // cell skipped by the reference (:810,:819) or smaller than 7 px
        if (lane == 0) *cellCount = 0;
        return;
    }
    const LevelPlan& L = P->lv[cd.level];
    const int rowsT = ih + 6;
    const int X0 = (cd.gx0 - 3) & ~15;
    uint8_t* tile = fcSmem;
    uint8_t* score = fcSmem + rowsT * TP;                     // TP is a multiple of 16
    unsigned short* candS = reinterpret_cast<unsigned short*>(score + (ih + 2) * SP);
    unsigned short* cornS = candS + FB_CANDS;
    unsigned short* allS = cornS + FB_CORNS;
    const uint8_t* roi = B.pyr + (size_t)frame * P->pyrStride + L.roiOff;
    if (lane == 0) {
        mbar_init(&sBar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sBar, rowsT * TP);
    }
    __syncwarp();
    for (int r = lane; r < rowsT; r += 32) tma_bulk_g2s(tile + r * TP, roi + (ptrdiff_t)(cd.gy0 - 3 + r) * L.pitch + X0, TP, &sBar);
    for (int i = lane; i < (ih + 2) * (SP / 16); i += 32) reinterpret_cast<uint4*>(score)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    mbar_wait(&sBar, 0);
    const int cx0 = cd.gx0 - X0, cx1 = cd.gx1 - X0;
    const int sxo = cx0 - 1;                                  // score column 0 = the zero column left of the cell
    // survivors are parked (unordered) in the quadtree's second key buffer, which k_octree only uses later
    unsigned* park = reinterpret_cast<unsigned*>(B.keys + ((size_t)frame * 2 + 1) * P->rawStride + cd.outOff);
    int nSurv = cell_pass<TP, SP>(tile, score, candS, cornS, allS, park, cx0, cx1, ih, min(max(P->iniTh, 0), 255), lane, sxo);
    if (nSurv == 0)                                           // :833-846 (scores do not depend on the threshold: the map stays valid)
        nSurv = cell_pass<TP, SP>(tile, score, candS, cornS, allS, park, cx0, cx1, ih, min(max(P->minTh, 0), 255), lane, sxo);
    // ---- E: raster order by rank counting; keys are relative to the 16-px border (:865-866) ----
    __syncwarp();
    u64* keysOut = B.cellKeys + (size_t)frame * P->cellKeyStride + cd.outOff;
    for (int i = lane; i < nSurv; i += 32) {
        const unsigned rec = __ldcg(park + i);
        int rank = 0;
        for (int k = 0; k < nSurv; k++) rank += __ldcg(park + k) < rec;
        const int x = cd.gx0 + (int)((rec >> 8) & 0xffu) - kMinBorder;
        const int y = cd.gy0 + (int)(rec >> 16) - kMinBorder;
        keysOut[rank] = (u64)(unsigned)x | ((u64)(unsigned)y << 16) | ((u64)(rec & 0xffu) << 32);
    }
    if (lane == 0) *cellCount = nSurv;
}

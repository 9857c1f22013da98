// K2: per-cell FAST-9/16 + non-maximum suppression + the iniThFAST -> minThFAST fallback (reference ORBextractor.cc:805-872:
// cv::FAST on each 35-px cell at iniThFAST, again at minThFAST when the cell stays empty).   (included INSIDE namespace orbb)
//
// One warp = one CTA = one cell with its own tile, so a finished cell frees its resources at once and up to 32 cells are in
// flight per SM.  The cell interiors tile the level and cv::FAST's NMS never looks across a cell border, so after the tile is in
// shared memory the warp needs nobody else:
//
//   stage   the cell's pixels (+3 px ring margin, 16-byte aligned window) by ONE tensor-map TMA copy (cp.async.bulk.tensor.3d:
//           x, y, frame of the level's bordered slab; SASS UTMALDG) -- or one bulk copy per row when no tensor map is available
//   per cell, a small state machine over one stack in shared memory keeps the expensive step at 32 busy lanes:
//     A     quick reject, 4 px x 2 rows per lane, byte-SIMD: a 9-arc always contains ring pixel 0 or 8 AND 4 or 12, so a
//           corner needs (|v-p0| > t or |v-p8| > t) and (|v-p4| > t or |v-p12| > t).  Per word: 4 VABSDIFF4, 4 adds, 4 LOP3.
//           The test may pass a non-corner (it is a filter; step S is exact), it never rejects a corner.  -> candidate stack
//     S     pops 32 candidates: EXACT score of both polarities at once.  Ring pixel p_k becomes ONE multiply-add
//               R_k = p_k * (1 - 2^16) + v * (2^16 - 1)   =   (p_k - v)  in the low 16 bits,  (v - p_k) - [p_k < v]  in the high 16
//           and the 9-arc minimum / maximum-over-arcs network runs on both halves with VIMNMX3.S16x2 (40 instructions):
//           low half = max over arcs of min(p_k - v) (bright), high half = the same for v - p_k, one short where positive --
//           the map d -> d - [d > 0] is monotone, so it commutes with min / max and is undone at the end.  score = max - 1
//           (cv::FAST: largest threshold for which the pixel is still a corner), corner <=> max > t.  -> score byte map + the
//           cell's corner list
//     D     NMS over the corner list inside the cell (strict '>' against the 8 neighbours; outside the cell counts as 0)
//     retry a cell without survivors runs A-D again at minThFAST on the pixels that are still in shared memory
//     E     raster order by rank counting -> cellKeys / cellCount  (= vToDistributeKeys, in the reference's order)
#pragma once

constexpr int FC_CANDS = 32 + 256;         // candidate stack: < 32 left over + one round of phase A (32 lanes x 8 px)
constexpr int FC_NC = 128;                 // corners of one cell kept for the list-driven NMS (more: NMS scans the score map)
constexpr int FC_STACK_BYTES = 2 * (FC_CANDS + FC_NC);

__device__ __forceinline__ int max3i(int a, int b, int c) { return max(max(a, b), c); }

// Quick reject of 4 pixels (one word): bit 7 of byte k is set when pixel k can still be a corner.
//   EXACT = false (both thresholds < 128, every real configuration): per byte x = |v - p|:  x > t  <=>  bit 7 of
//     ((x + (127 - t)) | x).  The adds run on whole words; a carry out of one byte can only turn the byte above it into a false
//     POSITIVE (x' + K + 1 reaches bit 7 for x' == t; if it wraps, x' >= 128 and the OR keeps bit 7), which step S sorts out.
//   EXACT = true (any threshold): s = (x & 0x7f) + (0x7f - (t & 0x7f)) carries into bit 7 iff low7(x) > low7(t); for t < 128 the
//     answer is s | x, for t >= 128 it is s & x.  M = t >= 128 ? ~0u : 0.
template <bool EXACT>
__device__ __forceinline__ unsigned gt_bytes(unsigned x, unsigned K7, unsigned M) {
    if (!EXACT) return (x + K7) | x;
    const unsigned s = (x & 0x7f7f7f7fu) + K7;
    return (M & s & x) | (~M & (s | x));
}
template <bool EXACT>
__device__ __forceinline__ unsigned quick_bytes4(unsigned wl, unsigned wc, unsigned wr, unsigned wt, unsigned wb, unsigned K7, unsigned M) {
    const unsigned pl = __funnelshift_r(wl, wc, 8);      // bytes x-3 .. x
    const unsigned pr = __funnelshift_r(wc, wr, 24);     // bytes x+3 .. x+6
    const unsigned v = gt_bytes<EXACT>(__vabsdiffu4(wc, wt), K7, M) | gt_bytes<EXACT>(__vabsdiffu4(wc, wb), K7, M);
    const unsigned h = gt_bytes<EXACT>(__vabsdiffu4(wc, pr), K7, M) | gt_bytes<EXACT>(__vabsdiffu4(wc, pl), K7, M);
    return v & h;
}

// One cell, NW warps (1 in batches; 4 in a call with a few frames, where the detector's duration is that of its slowest cell and
// the GPU is otherwise idle: the warps split the work items of phase A -- each with its own candidate stack -- and share the score
// map, the corner list and the survivor list through two counters in shared memory, sCnt = {corners, survivors}, zeroed by the
// caller).  tile: row t = level row gy0 - 3 + t, column = level column - X0 (pitch TP).  score: row s = interior row s - 1,
// column = tile column - sxo (pitch SP: the cell's columns plus a zero column on each side).
// Returns the number of NMS survivors parked in `park`.
template <int TP, int SP, bool EXACT, int NW>
__device__ __forceinline__ int cell_pass(const uint8_t* __restrict__ tile, uint8_t* __restrict__ score, unsigned short* candS,
                                         unsigned short* allS, unsigned* park, const CellDesc& cd, int ih, int th, unsigned K7,
                                         int lane, int warp = 0, int* sCnt = nullptr, const unsigned** parkUsed = nullptr) {
    constexpr int PS = TP, TPW = TP / 4;
    const int cx0 = cd.cx0, cx1 = cd.cx1, sxo = cx0 - 1;                 // score column 0 = the zero column left of the cell
    const int wa = cd.wa, wLast = cd.wLast, nwc = cd.nwc, items = cd.items;
    const unsigned mInv = cd.mInv, mF7 = cd.mF7, mL7 = cd.mL7;
    const unsigned M = th >= 128 ? 0xffffffffu : 0u;
    const unsigned* tileW = reinterpret_cast<const unsigned*>(tile) + wa;
    const unsigned lt = (1u << lane) - 1;
    int nCand = 0, nAll = 0, base = 32 * warp;
    if (NW > 1) candS += warp * FC_CANDS;
    for (;;) {
        const bool aDone = base >= items;
        if (nCand >= 32 || (aDone && nCand > 0)) {
            // ---- S: exact score of up to 32 candidates ----
            const int n = min(nCand, 32);
            nCand -= n;
            bool isCorner = false;
            int sp = 0;
            if (lane < n) {
                const unsigned rec = candS[nCand + lane];                // row << 8 | column (tile coordinates)
                const int row = rec >> 8, col = rec & 0xff;
                const uint8_t* q = tile + row * PS + col;
                const unsigned C = (unsigned)q[0] * 65535u;
#define ORBB_RING(off) ((unsigned)q[off] * 0xFFFF0001u + C)
                unsigned d[16];
                d[0] = ORBB_RING(3 * PS);       d[1] = ORBB_RING(3 * PS + 1);   d[2] = ORBB_RING(2 * PS + 2);   d[3] = ORBB_RING(PS + 3);
                d[4] = ORBB_RING(3);            d[5] = ORBB_RING(-PS + 3);      d[6] = ORBB_RING(-2 * PS + 2);  d[7] = ORBB_RING(-3 * PS + 1);
                d[8] = ORBB_RING(-3 * PS);      d[9] = ORBB_RING(-3 * PS - 1);  d[10] = ORBB_RING(-2 * PS - 2); d[11] = ORBB_RING(-PS - 3);
                d[12] = ORBB_RING(-3);          d[13] = ORBB_RING(PS - 3);      d[14] = ORBB_RING(2 * PS - 2);  d[15] = ORBB_RING(3 * PS - 1);
#undef ORBB_RING
                unsigned m3[16];
#pragma unroll
                for (int k = 0; k < 16; k++) m3[k] = __vimin3_s16x2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
                unsigned e[16];                                          // e[k] = min over the arc k .. k+8
#pragma unroll
                for (int k = 0; k < 16; k++) e[k] = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
                const unsigned t0 = __vimax3_s16x2(e[0], e[1], e[2]), t1 = __vimax3_s16x2(e[3], e[4], e[5]), t2 = __vimax3_s16x2(e[6], e[7], e[8]);
                const unsigned t3 = __vimax3_s16x2(e[9], e[10], e[11]), t4 = __vimax3_s16x2(e[12], e[13], e[14]);
                const unsigned u0 = __vimax3_s16x2(t0, t1, t2), u1 = __vimax3_s16x2(t3, t4, e[15]);
                const unsigned X = __vimax3_s16x2(u0, u1, u1);
                const int bright = (int)(short)(X & 0xffffu);
                int dark = (int)X >> 16;
                dark += dark > 0;                                        // undo d -> d - [d > 0]
                const int M = max(bright, dark);
                isCorner = M > th;
                sp = (row - 2) * SP + col - sxo;                         // tile row = interior row + 3, score row = interior row + 1
                if (isCorner) score[sp] = (uint8_t)(M - 1);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, isCorner);
            if (NW > 1) {                                            // (the corner list is shared by the warps of the cell)
                int first = 0;
                if (lane == 0 && bal) first = atomicAdd(&sCnt[0], __popc(bal));
                nAll = __shfl_sync(0xffffffffu, first, 0);
            }
            const int slot = nAll + __popc(bal & lt);
            if (isCorner && slot < FC_NC) allS[slot] = (unsigned short)sp;
            nAll += __popc(bal);
            __syncwarp();
            continue;
        }
        if (aDone) break;
        // ---- A: quick reject, one word column x 2 rows per lane ----
        {
            const int i = base + lane;
            unsigned cand = 0, basev = 0;
            if (i < items) {
                const int s = (int)__umulhi((unsigned)i, mInv);
                const int w = i - s * nwc;
                const int r0 = 2 * s;
                unsigned m7 = w == 0 ? mF7 : 0x80808080u;
                if (w == wLast) m7 &= mL7;
                if (w > wLast) m7 = 0u;
                const unsigned m7b = r0 + 1 < ih ? m7 : 0u;              // (the row below the cell is readable: it is masked, not skipped)
                const unsigned* q = tileW + (r0 + 3) * TPW + w;
                const unsigned q0 = quick_bytes4<EXACT>(q[-1], q[0], q[1], q[3 * TPW], q[-3 * TPW], K7, M);
                const unsigned q1 = quick_bytes4<EXACT>(q[TPW - 1], q[TPW], q[TPW + 1], q[4 * TPW], q[-2 * TPW], K7, M);
                cand = (q0 & m7) | ((q1 & m7b) >> 1);                    // bit 8k+7: pixel k of row r0; bit 8k+6: row r0 + 1
                basev = ((unsigned)(r0 + 3) << 8) | (unsigned)(4 * (wa + w));
            }
            // exclusive prefix of the per-lane candidate counts (0..8) from four ballots of the count's bits: independent votes
            // instead of the five dependent shuffle steps of a scan (13 % of the kernel's stall samples sat on that chain)
            const int cnt = __popc(cand);
            const unsigned v0 = __ballot_sync(0xffffffffu, cnt & 1), v1 = __ballot_sync(0xffffffffu, cnt & 2);
            const unsigned v2 = __ballot_sync(0xffffffffu, cnt & 4), v3 = __ballot_sync(0xffffffffu, cnt & 8);
            const int excl = __popc(v0 & lt) + 2 * __popc(v1 & lt) + 4 * __popc(v2 & lt) + 8 * __popc(v3 & lt);
            unsigned short* o = candS + (nCand + excl);
            nCand += __popc(v0) + 2 * __popc(v1) + 4 * __popc(v2) + 8 * __popc(v3);
#pragma unroll
            for (int k = 0; k < 4; k++) {                                // fixed predicated sequence (the order on the stack is irrelevant)
                if (cand & (0x80u << (8 * k))) *o++ = (unsigned short)(basev + k);
                if (cand & (0x40u << (8 * k))) *o++ = (unsigned short)(basev + 256 + k);
            }
            base += 32 * NW;
            __syncwarp();
        }
    }
    if (NW > 1) {
        __syncthreads();                                             // every warp's scores and corners are in
        nAll = sCnt[0];
    }
    // ---- D: NMS inside the cell (strict '>' against the 8 neighbours; outside the cell counts as 0) ----
    // one warp per cell: the survivors of an ordinary cell (<= FC_NC corners) are parked on the candidate stack, which is free by now
    // (phase E reads every record nSurv times: shared memory instead of L2 round trips); dense cells use the global park
    if (NW == 1 && nAll <= FC_NC) park = reinterpret_cast<unsigned*>(candS);
    if (parkUsed) *parkUsed = park;
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
        if (NW > 1) {
            int first = 0;
            if (lane == 0 && bal) first = atomicAdd(&sCnt[1], __popc(bal));
            nSurv = __shfl_sync(0xffffffffu, first, 0);
        }
        if (keep) park[nSurv + __popc(bal & lt)] = rec;
        nSurv += __popc(bal);
    };
    if (nAll <= FC_NC) {
        for (int b0 = 32 * warp; b0 < nAll; b0 += 32 * NW) {
            const bool valid = b0 + lane < nAll;
            const int sp = valid ? allS[b0 + lane] : 0;
            nms(valid, sp, valid ? score[sp] : 0);
        }
    } else {                                                          // very dense cell: walk its score map
        const int wc = cx1 - cx0;
        for (int b0 = 32 * warp; b0 < ih * wc; b0 += 32 * NW) {
            const int i = b0 + lane;
            const int r = i / wc, x = cx0 + i - r * wc;
            const int sp = (r + 1) * SP + x - sxo;
            const int sc = i < ih * wc ? score[sp] : 0;
            nms(sc > 0, sp, sc);
        }
    }
    if (NW > 1) {
        __syncthreads();                                             // all survivors are parked (global memory, same CTA)
        nSurv = sCnt[1];
    }
    return nSurv;
}

// TP: tile pitch in bytes (64 when every cell + 6 px margin + 15 px alignment slop fits, else 96).  TMAP: the tile comes in by
// one tensor-map TMA copy (tmaps = one CUtensorMap per level, a kernel parameter: dims {pitch, rows, frames} of the level's bordered
// slab; frame0 = index of this launch's first frame within the slab) instead of one bulk copy per row.
// (Two independent cells per 64-thread CTA -- shared memory instead of the 32-CTA limit bounding the cells in flight, 39 instead of 32
// per SM -- measured slower: 0.515 vs 0.477 ms per 256 frames.)
template <int TP, bool TMAP>
__global__ void __launch_bounds__(32, 32) k_fast_cell(const Plan* __restrict__ P, Bufs B, const __grid_constant__ TmapTable tmaps, int frame0, int cell0) {
    constexpr int SP = TP == 64 ? 48 : 80;                    // score pitch: cell width (<= 43 / 71) + a zero column on each side
    extern __shared__ __align__(128) uint8_t fcSmem[];
    __shared__ __align__(8) unsigned long long sBar;
    const int gcell = blockIdx.x + cell0, frame = blockIdx.y, lane = threadIdx.x;
    pdl_launch_dependents();
    const CellDesc cd = B.cellDesc[gcell];                    // (plan data, not a product of the previous kernel)
    int* cellCount = B.cellCount + (size_t)frame * P->cellsTotal + gcell;
    const int ih = cd.gy1 - cd.gy0;
    if (cd.gx1 <= cd.gx0 || ih <= 0) {                        // cell skipped by the reference (:810,:819) or smaller than 7 px
        pdl_wait();                                           // (the quadtree of the previous batch read this count)
        if (lane == 0) *cellCount = 0;
        return;
    }
    const LevelPlan& L = P->lv[cd.level];
    const int rowsT = ih + 6;
    const int X0 = (cd.gx0 - 3) & ~15;
    uint8_t* tile = fcSmem;
    const int tileRows = TMAP ? P->cellRows : rowsT;          // (the tensor-map box has a fixed height: the tallest cell's)
    uint8_t* score = fcSmem + tileRows * TP;                  // TP is a multiple of 16
    unsigned short* candS = reinterpret_cast<unsigned short*>(score + (ih + 2) * SP);
    unsigned short* allS = candS + FC_CANDS;
    if (lane == 0) {
        mbar_init(&sBar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sBar, tileRows * TP);
    }
    pdl_wait();                                               // the pyramid is complete from here on
    if (lane == 0) {
        if (TMAP) tma_tensor3d_g2s(tile, &tmaps.m[cd.level], kRoiX + X0, kEdge + cd.gy0 - 3, frame0 + frame, &sBar);
    }
    __syncwarp();
    if (!TMAP) {
        const uint8_t* roi = B.pyr + (size_t)frame * P->pyrStride + L.roiOff;
        for (int r = lane; r < rowsT; r += 32) tma_bulk_g2s(tile + r * TP, roi + (ptrdiff_t)(cd.gy0 - 3 + r) * L.pitch + X0, TP, &sBar);
    }
    for (int i = lane; i < (ih + 2) * (SP / 16); i += 32) reinterpret_cast<uint4*>(score)[i] = make_uint4(0, 0, 0, 0);
    __syncwarp();
    mbar_wait(&sBar, 0);
    // survivors are parked (unordered) in the quadtree's second key buffer, which k_octree only uses later
    unsigned* park = reinterpret_cast<unsigned*>(B.keys + ((size_t)frame * 2 + 1) * P->rawStride + cd.outOff);
    const int iniTh = P->iniTh, minTh = P->minTh;
    int nSurv;
    const unsigned* parked = park;
    if (iniTh < 128 && minTh < 128) {
        nSurv = cell_pass<TP, SP, false, 1>(tile, score, candS, allS, park, cd, ih, iniTh, P->k7Ini, lane, 0, nullptr, &parked);
        if (nSurv == 0)                                       // :833-846 (scores do not depend on the threshold: the map stays valid)
            nSurv = cell_pass<TP, SP, false, 1>(tile, score, candS, allS, park, cd, ih, minTh, P->k7Min, lane, 0, nullptr, &parked);
    } else {                                                  // thresholds >= 128: exact byte compare in the quick reject
        nSurv = cell_pass<TP, SP, true, 1>(tile, score, candS, allS, park, cd, ih, iniTh, P->k7Ini, lane, 0, nullptr, &parked);
        if (nSurv == 0) nSurv = cell_pass<TP, SP, true, 1>(tile, score, candS, allS, park, cd, ih, minTh, P->k7Min, lane, 0, nullptr, &parked);
    }
    // ---- E: raster order by rank counting; keys are relative to the 16-px border (:865-866) ----
    __syncwarp();
    u64* keysOut = B.cellKeys + (size_t)frame * P->cellKeyStride + cd.outOff;
    const bool parkedShared = parked != park;
    for (int i = lane; i < nSurv; i += 32) {
        const unsigned rec = parkedShared ? parked[i] : __ldcg(park + i);
        int rank = 0;
        if (parkedShared) for (int k = 0; k < nSurv; k++) rank += parked[k] < rec;
        else for (int k = 0; k < nSurv; k++) rank += __ldcg(park + k) < rec;
        const int x = cd.gx0 + (int)((rec >> 8) & 0xffu) - kMinBorder;
        const int y = cd.gy0 + (int)(rec >> 16) - kMinBorder;
        keysOut[rank] = (u64)(unsigned)x | ((u64)(unsigned)y << 16) | ((u64)(rec & 0xffu) << 32);
    }
    if (lane == 0) *cellCount = nSurv;
}

// The same cell with FC_MW warps (see cell_pass): for calls with a few frames.  Results are identical -- the order in which
// candidates, corners and survivors are found does not matter (NMS reads the finished score map, phase E sorts by rank).
constexpr int FC_MW = 4;
template <int TP, bool TMAP>
__global__ void __launch_bounds__(32 * FC_MW) k_fast_cell_mw(const Plan* __restrict__ P, Bufs B, const __grid_constant__ TmapTable tmaps, int frame0, int cell0) {
    constexpr int SP = TP == 64 ? 48 : 80;
    extern __shared__ __align__(128) uint8_t fcSmem[];
    __shared__ __align__(8) unsigned long long sBar;
    __shared__ int sCnt[2];
    const int gcell = blockIdx.x + cell0, frame = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    const CellDesc cd = B.cellDesc[gcell];
    int* cellCount = B.cellCount + (size_t)frame * P->cellsTotal + gcell;
    const int ih = cd.gy1 - cd.gy0;
    pdl_wait();                                               // the pyramid is complete from here on
    if (cd.gx1 <= cd.gx0 || ih <= 0) {
        if (tid == 0) *cellCount = 0;
        return;
    }
    const LevelPlan& L = P->lv[cd.level];
    const int rowsT = ih + 6;
    const int X0 = (cd.gx0 - 3) & ~15;
    uint8_t* tile = fcSmem;
    const int tileRows = TMAP ? P->cellRows : rowsT;
    uint8_t* score = fcSmem + tileRows * TP;
    unsigned short* candS = reinterpret_cast<unsigned short*>(score + (ih + 2) * SP);
    unsigned short* allS = candS + FC_MW * FC_CANDS;
    if (tid == 0) {
        mbar_init(&sBar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sBar, tileRows * TP);
        if (TMAP) tma_tensor3d_g2s(tile, &tmaps.m[cd.level], kRoiX + X0, kEdge + cd.gy0 - 3, frame0 + frame, &sBar);
        sCnt[0] = sCnt[1] = 0;
    }
    __syncthreads();
    if (!TMAP) {
        const uint8_t* roi = B.pyr + (size_t)frame * P->pyrStride + L.roiOff;
        for (int r = tid; r < rowsT; r += 32 * FC_MW) tma_bulk_g2s(tile + r * TP, roi + (ptrdiff_t)(cd.gy0 - 3 + r) * L.pitch + X0, TP, &sBar);
    }
    for (int i = tid; i < (ih + 2) * (SP / 16); i += 32 * FC_MW) reinterpret_cast<uint4*>(score)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    mbar_wait(&sBar, 0);
    unsigned* park = reinterpret_cast<unsigned*>(B.keys + ((size_t)frame * 2 + 1) * P->rawStride + cd.outOff);
    const int iniTh = P->iniTh, minTh = P->minTh;
    const bool exact = iniTh >= 128 || minTh >= 128;
    int nSurv = exact ? cell_pass<TP, SP, true, FC_MW>(tile, score, candS, allS, park, cd, ih, iniTh, P->k7Ini, lane, warp, sCnt)
                      : cell_pass<TP, SP, false, FC_MW>(tile, score, candS, allS, park, cd, ih, iniTh, P->k7Ini, lane, warp, sCnt);
    if (nSurv == 0) {                                         // :833-846 (uniform: nSurv comes out of shared memory behind a barrier)
        __syncthreads();
        if (tid == 0) sCnt[0] = sCnt[1] = 0;
        __syncthreads();
        nSurv = exact ? cell_pass<TP, SP, true, FC_MW>(tile, score, candS, allS, park, cd, ih, minTh, P->k7Min, lane, warp, sCnt)
                      : cell_pass<TP, SP, false, FC_MW>(tile, score, candS, allS, park, cd, ih, minTh, P->k7Min, lane, warp, sCnt);
    }
    // ---- E: raster order by rank counting ----
    u64* keysOut = B.cellKeys + (size_t)frame * P->cellKeyStride + cd.outOff;
    for (int i = tid; i < nSurv; i += 32 * FC_MW) {
        const unsigned rec = __ldcg(park + i);
        int rank = 0;
        for (int k = 0; k < nSurv; k++) rank += __ldcg(park + k) < rec;
        const int x = cd.gx0 + (int)((rec >> 8) & 0xffu) - kMinBorder;
        const int y = cd.gy0 + (int)(rec >> 16) - kMinBorder;
        keysOut[rank] = (u64)(unsigned)x | ((u64)(unsigned)y << 16) | ((u64)(rec & 0xffu) << 32);
    }
    if (tid == 0) *cellCount = nSurv;
}

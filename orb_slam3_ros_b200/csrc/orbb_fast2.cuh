// K2, level-wide formulation (production path): the FAST score of a pixel does not depend on the 35-px cell it falls
// in -- only the non-maximum suppression and the "no corner in this cell -> retry with minThFAST" decision do
// (reference ORBextractor.cc:805-872).  So:
//
//   k_fast_score   regular tiles over each level, no cell geometry: quick reject (4 px / thread / row, sliding 7-row
//                  register window, byte-SIMD) -> per-CTA candidate list -> 16-pixel ring test -> exact score.
//                  Writes a score byte map (0 = not a corner at iniThFAST, else response = arc score - 1).
//   k_fast_cells   one CTA per cell: loads the cell interior of the score map (zero outside: NMS never looks across a
//                  cell border), NMS, raster-ordered output.  Cells without any survivor are marked -1 ...
//   k_fast(mode 1) ... and only those run the per-cell pipeline of orbb_fast.cuh again at minThFAST.
//
// (included INSIDE namespace orbb, after orbb_fast.cuh)
#pragma once

constexpr int FS_THREADS = 128, FS_R = 8;
constexpr int FS_CAP = FS_THREADS * FS_R * 4;       // pixels per CTA

__device__ __forceinline__ unsigned quick_mask4(unsigned wl, unsigned wc, unsigned wr, unsigned wt, unsigned wb, unsigned K) {
    const unsigned pl = __funnelshift_r(wl, wc, 8);      // bytes x-3 .. x
    const unsigned pr = __funnelshift_r(wc, wr, 24);     // bytes x+3 .. x+6
    const unsigned a0 = __vabsdiffu4(wc, wt), a8 = __vabsdiffu4(wc, wb);
    const unsigned a4 = __vabsdiffu4(wc, pr), a12 = __vabsdiffu4(wc, pl);
    const unsigned me = umin16x2(umax16x2(prmt(a0, 0, 0x4240), prmt(a8, 0, 0x4240)), umax16x2(prmt(a4, 0, 0x4240), prmt(a12, 0, 0x4240)));
    const unsigned mo = umin16x2(umax16x2(prmt(a0, 0, 0x4341), prmt(a8, 0, 0x4341)), umax16x2(prmt(a4, 0, 0x4341), prmt(a12, 0, 0x4341)));
    const unsigned te = me + K, to = mo + K;             // bit 15 / 31 of a 16-bit lane set  <=>  value > th
    return ((te >> 15) & 1u) | ((to >> 14) & 2u) | ((te >> 29) & 4u) | ((to >> 28) & 8u);
}

// tile of one CTA: 32 word-columns with 16 halo bytes on each side (so every tile row starts 16-byte aligned in global
// memory and can be fetched by one TMA bulk copy) x (4 strips x FS_R rows + 6 halo rows)
constexpr int FS_TP = 128 + 32;                     // tile pitch in bytes
constexpr int FS_XOFF = 16;                         // tile byte of the CTA's first image column
constexpr int FS_TROWS = 4 * FS_R + 6;

__global__ void __launch_bounds__(FS_THREADS) k_fast_score(const Plan* __restrict__ P, Bufs B) {
    __shared__ __align__(128) uint8_t sTile[FS_TROWS * FS_TP];
    __shared__ __align__(8) unsigned long long sBar;
    __shared__ unsigned short sCand[FS_CAP];         // tile position of each candidate
    __shared__ unsigned short sCorner[FS_CAP];       // tile position | polarity << 15
    __shared__ int sCnt[2];
    const int frame = blockIdx.y;
    int level = 0;
    while (level + 1 < P->nlevels && (int)blockIdx.x >= P->lv[level + 1].fsBase) level++;
    const LevelPlan& L = P->lv[level];
    const int t = blockIdx.x - L.fsBase;
    const int gy = t / L.fsTilesX, gx = t - gy * L.fsTilesX;
    const int tid = threadIdx.x, lane = tid & 31, ty = tid >> 5;
    const int xt0 = gx * 128 - FS_XOFF;              // level column of tile byte 0
    const int yt0 = kEdge + gy * 4 * FS_R - 3;       // level row of tile row 0
    const int xlo = kEdge, xhi = L.w - kEdge, yhi = L.h - kEdge;      // union of the cell interiors: [19, w-19) x [19, h-19)
    const int th = min(max(P->iniTh, 0), 255);
    const unsigned K = (unsigned)(0x7fff - th) * 0x00010001u;
    const uint8_t* roi = B.pyr + (size_t)frame * P->pyrStride + L.roiOff;
    uint8_t* score = B.score + (size_t)frame * P->blurStride + L.blurOff;
    if (tid < 2) sCnt[tid] = 0;
    // ---- stage the tile: one 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) per tile row, all completing on one mbarrier.
    // Rows below the bottom apron are clamped (never used); bytes right of the apron come from the row's alignment
    // padding / the next row (never used either).
    if (tid == 0) {
        mbar_init(&sBar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sBar, FS_TROWS * FS_TP);
    }
    __syncthreads();
    if (tid < FS_TROWS) {
        const int gyy = min(yt0 + tid, L.h + kEdge - 1);
        tma_bulk_g2s(sTile + tid * FS_TP, roi + (ptrdiff_t)gyy * L.pitch + xt0, FS_TP, &sBar);
    }
    mbar_wait(&sBar, 0);

    // ---- A: quick reject over FS_R rows, 4 pixels per thread and row ----
    const int wcol = gx * 32 + lane;
    const int y0 = kEdge + (gy * 4 + ty) * FS_R;
    unsigned candAll = 0;
    if (wcol * 4 < L.w && y0 < yhi) {
        unsigned xmask = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) xmask |= (unsigned)(wcol * 4 + k >= xlo && wcol * 4 + k < xhi) << k;
        const unsigned* tw = reinterpret_cast<const unsigned*>(sTile) + (ty * FS_R) * (FS_TP / 4) + lane + FS_XOFF / 4;
        unsigned* srow = reinterpret_cast<unsigned*>(score + (size_t)y0 * L.bpitch + 4 * wcol);
#pragma unroll
        for (int r = 0; r < FS_R; r++) {
            const unsigned* c = tw + (r + 3) * (FS_TP / 4);
            const unsigned m = quick_mask4(c[-1], c[0], c[1], c[3 * (FS_TP / 4)], c[-3 * (FS_TP / 4)], K) & xmask;
            if (y0 + r < yhi) {
                srow[(size_t)r * (L.bpitch / 4)] = 0u;
                candAll |= m << (4 * r);
            }
        }
    }
    {
        const int cnt = __popc(candAll);
        const int inc = warp_incl_scan(cnt, lane);
        const int wtot = __shfl_sync(0xffffffffu, inc, 31);
        int wbase = 0;
        if (lane == 31 && wtot) wbase = atomicAdd(&sCnt[0], wtot);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        int o = wbase + inc - cnt;
        const int pos0 = (ty * FS_R + 3) * FS_TP + lane * 4 + FS_XOFF;
        while (candAll) {
            const int b = __ffs(candAll) - 1;
            candAll &= candAll - 1;
            sCand[o++] = (unsigned short)(pos0 + (b >> 2) * FS_TP + (b & 3));
        }
    }
    __syncthreads();

    // ---- B: full ring test on the candidates ----
    constexpr int PS = FS_TP;
    const int nCand = sCnt[0];
    for (int base = 0; base < nCand; base += FS_THREADS) {
        const int i = base + tid;
        bool corner = false;
        unsigned rec = 0;
        if (i < nCand) {
            const unsigned pos = sCand[i];
            const uint8_t* q = &sTile[pos];
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
            corner = cb | cd;
            rec = pos | (cd ? 0x8000u : 0u);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, corner);
        int wbase = 0;
        if (lane == 0 && bal) wbase = atomicAdd(&sCnt[1], __popc(bal));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (corner) sCorner[wbase + __popc(bal & ((1u << lane) - 1))] = (unsigned short)rec;
    }
    __syncthreads();

    // ---- C: exact score of the corners -> score map ----
    const int nCorner = sCnt[1];
    for (int i = tid; i < nCorner; i += FS_THREADS) {
        const unsigned rec = sCorner[i];
        const int pos = rec & 0x7fff;
        const uint8_t* q = &sTile[pos];
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
        const int r = pos / PS, col = pos - r * PS;
        score[(size_t)(yt0 + r) * L.bpitch + (xt0 + col)] = (uint8_t)(M - 1);
    }
}

constexpr int FC_THREADS = 64;
constexpr int FC_CAP = 1408;                        // >= max NMS survivors of one cell (37 x 37)

// One CTA per cell; 5.6 KB of shared memory so that 32 CTAs fit an SM (the kernel is a chain of dependent loads, not
// arithmetic).  Corners are sparse (~1 % of the pixels): the cell's part of the score map is scanned straight from
// global memory one word (4 px) per thread; non-zero bytes go to a small list, and only those are compared with their 8
// neighbours (read from the map again, zero outside the cell interior: cv::FAST ran on the cell alone).  Survivors are
// parked in the cell's output segment, then reloaded, ranked by raster position and written back in order.
__global__ void __launch_bounds__(FC_THREADS) k_fast_cells(Bufs B, int cellsTotal, u64 scoreStride, unsigned cellKeyStride) {
    __shared__ unsigned sList[FC_CAP];              // (row << 16 | col << 8 | score)
    __shared__ int sCnt[2];
    const int frame = blockIdx.y;
    const int gcell = blockIdx.x;
    const int tid = threadIdx.x;
    const CellDesc cd = B.cellDesc[gcell];
    int* cellCount = B.cellCount + (size_t)frame * cellsTotal + gcell;
    if (cd.gx1 <= cd.gx0) {                         // :810,:819
        if (tid == 0) *cellCount = 0;
        return;
    }
    const int gx0 = cd.gx0, gx1 = cd.gx1, gy0 = cd.gy0, gy1 = cd.gy1;     // cell interior, level coordinates
    const int ih = gy1 - gy0, bp = cd.bpitch;
    const int a0 = gx0 & ~3;
    const int nw = (gx1 - a0 + 3) >> 2;             // words per interior row (<= 20)
    const unsigned mw = (65536u + nw - 1) / nw;
    const uint8_t* score = B.score + (size_t)frame * scoreStride + cd.scoreOff;
    u64* out = B.cellKeys + (size_t)frame * cellKeyStride + cd.outOff;
    unsigned* park = reinterpret_cast<unsigned*>(out);      // survivors before ordering (the segment holds >= 37*37 keys)
    if (tid < 2) sCnt[tid] = 0;
    const int bandRows = max(1, FC_CAP / (4 * nw));          // a band can never overflow the list (one band for normal cells)
    for (int rb = 0; rb < ih; rb += bandRows) {
        const int rows = min(bandRows, ih - rb);
        __syncthreads();
        if (tid == 0) sCnt[0] = 0;
        __syncthreads();
        for (int i = tid; i < rows * nw; i += FC_THREADS) {
            const int rr = (int)(((unsigned)i * mw) >> 16), w = i - rr * nw;
            const int r = rb + rr;
            const int col0 = a0 + 4 * w;
            const int lo = max(gx0 - col0, 0), hi = min(gx1 - col0, 4);
            unsigned v = __ldg(reinterpret_cast<const unsigned*>(score + (size_t)(gy0 + r) * bp + col0));
            v &= (0xffffffffu >> (8 * (4 - hi))) & (0xffffffffu << (8 * lo));
            while (v) {
                const int k = (__ffs(v) - 1) >> 3;
                const unsigned sc = (v >> (8 * k)) & 0xffu;
                v &= ~(0xffu << (8 * k));
                sList[atomicAdd(&sCnt[0], 1)] = ((unsigned)r << 16) | ((unsigned)(col0 + k - gx0) << 8) | sc;
            }
        }
        __syncthreads();
        // ---- NMS: strict '>' against the 8 neighbours (cv::FAST, fast.cpp) ----
        const int nCorner = sCnt[0];
        for (int i = tid; i < nCorner; i += FC_THREADS) {
            const unsigned rec = sList[i];
            const int r = (int)(rec >> 16), cx = (int)((rec >> 8) & 0xffu), sc = (int)(rec & 0xffu);
            const int x = gx0 + cx, y = gy0 + r;
            const uint8_t* q = score + (size_t)y * bp + x;
            const bool left = x > gx0, right = x + 1 < gx1, up = y > gy0, down = y + 1 < gy1;
            bool keep = true;
            if (left) keep &= sc > q[-1];
            if (right) keep &= sc > q[1];
            if (up) {
                keep &= sc > q[-bp];
                if (left) keep &= sc > q[-bp - 1];
                if (right) keep &= sc > q[-bp + 1];
            }
            if (down) {
                keep &= sc > q[bp];
                if (left) keep &= sc > q[bp - 1];
                if (right) keep &= sc > q[bp + 1];
            }
            if (keep) park[atomicAdd(&sCnt[1], 1)] = rec;    // rec orders by (row, col): the raster key
        }
    }
    __syncthreads();
    // ---- raster order by rank counting; keys are relative to the 16-px border (:865-866) ----
    const int nSurv = sCnt[1];
    for (int i = tid; i < nSurv; i += FC_THREADS) sList[i] = park[i];
    __syncthreads();
    for (int i = tid; i < nSurv; i += FC_THREADS) {
        const unsigned rec = sList[i];
        int rank = 0;
        for (int j = 0; j < nSurv; j++) rank += sList[j] < rec;
        const int x = gx0 + (int)((rec >> 8) & 0xffu) - kMinBorder, y = gy0 + (int)(rec >> 16) - kMinBorder;
        out[rank] = (u64)(unsigned)x | ((u64)(unsigned)y << 16) | ((u64)(rec & 0xffu) << 32);
    }
    if (tid == 0) {
        *cellCount = nSurv;
        if (nSurv == 0) {                           // retry this cell with minThFAST (k_fast, mode 1)
            const int slot = atomicAdd(&B.fbCount[frame], 1);
            B.fbList[(size_t)frame * cellsTotal + slot] = gcell;
        }
    }
}

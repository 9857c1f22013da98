// K1, production pyramid kernels (ComputePyramid, reference ORBextractor.cc:1170-1195).  (included INSIDE namespace orbb)
//
//   k_pyr_level0     copy of the source into its bordered slab, 16 bytes per thread.
//   k_pyr_resize_t   (production) = k_pyr_resize_s with the CTA's source rows staged in shared memory by TMA bulk copies
//   k_pyr_resize_s   cv::resize INTER_LINEAR for scale steps up to ~1.3 (the reference's 1.2): a thread owns FOUR destination columns and walks
//                    down PR_ROWS destination rows.  cv::resize's horizontal pass (HResizeLinear) of a source row is
//                    computed once and reused by the next destination row (consecutive destination rows share a source
//                    row at scale 1.2): three aligned word loads cover the <= 9 source bytes of the four columns, the two
//                    taps of a column come out of them with one funnel shift, the tap sum is one IDP.2A
//                    (u16 coefficients x u8 pixels).  The vertical pass is two IMAD.HI per pixel.
//   k_pyr_apron16    reflect-101 apron of every level: one thread per 16-byte chunk that touches the apron (on demand only: when
//                    the pyramid is handed out as mvImagePyramid; no kernel of the pipeline reads the apron).
//
// The generic k_pyr_resize (orbb_extract.cu) stays as the fallback for larger scale steps.
#pragma once

#ifndef ORBB_PR_ROWS
#define ORBB_PR_ROWS 16
#endif
constexpr int PR_ROWS = ORBB_PR_ROWS;  // destination rows per thread (measured per 256 frames: 8 rows 0.267 ms, 12: 0.258, 16: 0.258 -- the prologue of a thread is a fifth of its instructions at 8)
constexpr int PR_ROWS_LATENCY = 2;    // ... in calls with a few frames (k_pyr_resize_t only)
constexpr int PR_THREADS = 128;       // 32 word-columns x 4 row strips

__global__ void __launch_bounds__(256) k_pyr_level0_v(const Plan* __restrict__ P, Bufs B, const uint8_t* __restrict__ src,
                                                      size_t rowStride, size_t frameStride, int aligned16) {
    const LevelPlan& L = P->lv[0];
    const int chunk = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int frame = blockIdx.z;
    const int x0 = chunk * 16;
    if (x0 >= L.w || y >= L.h) return;
    const uint8_t* s = src + (size_t)frame * frameStride + (size_t)y * rowStride + x0;
    uint8_t* d = B.pyr + (size_t)frame * P->pyrStride + L.roiOff + (size_t)y * L.pitch + x0;      // roiOff, pitch: 16-byte aligned
    if (aligned16 && x0 + 16 <= L.w) {
        *reinterpret_cast<uint4*>(d) = __ldg(reinterpret_cast<const uint4*>(s));
    } else {
        const int n = min(16, L.w - x0);
        for (int k = 0; k < n; k++) d[k] = __ldg(s + k);
    }
}

// per-thread constants of one destination column: packed coefficients, byte shift inside the word pair, and a mask that
// selects the word pair (words 0/1 or 1/2 of the three loaded ones)
struct ResizeCol { unsigned coef, sh, m; };

__device__ __forceinline__ void hresize4(const unsigned* __restrict__ rp, const ResizeCol (&c)[4], int (&h)[4]) {
    const unsigned wa = __ldg(rp), wb = __ldg(rp + 1), wc = __ldg(rp + 2);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const unsigned lo = (wa & ~c[k].m) | (wb & c[k].m), hi = (wb & ~c[k].m) | (wc & c[k].m);      // one LOP3 each
        h[k] = (int)(__dp2a_lo(c[k].coef, __funnelshift_r(lo, hi, c[k].sh), 0u) >> 4);      // (S[sx]*a0 + S[sx+1]*a1) >> 4
    }
}

// PRMT = the taps of the first three columns lie in the first two words for every thread of the level (LevelPlan::prmtTaps,
// true at the reference's scale 1.2): one byte permute with a per-thread selector (held in ResizeCol::sh) picks both taps
// instead of two mask-selects and a funnel shift.
template <bool PRMT>
__device__ __forceinline__ void hresize4s(const unsigned* rp, const ResizeCol (&c)[4], int (&h)[4]) {      // rp: shared memory
    const unsigned wa = rp[0], wb = rp[1], wc = rp[2];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        unsigned taps;
        if (PRMT && k < 3) {
            taps = __byte_perm(wa, wb, c[k].sh);
        } else {
            const unsigned lo = (wa & ~c[k].m) | (wb & c[k].m), hi = (wb & ~c[k].m) | (wc & c[k].m);
            taps = __funnelshift_r(lo, hi, c[k].sh);
        }
        h[k] = (int)(__dp2a_lo(c[k].coef, taps, 0u) >> 4);
    }
}

__global__ void __launch_bounds__(PR_THREADS, 16) k_pyr_resize_s(const Plan* __restrict__ P, Bufs B, int level) {
    const LevelPlan& L = P->lv[level];
    const LevelPlan& S = P->lv[level - 1];
    const int word = blockIdx.x * 32 + (threadIdx.x & 31);
    const int dy0 = (blockIdx.y * (PR_THREADS / 32) + (threadIdx.x >> 5)) * PR_ROWS;
    const int frame = blockIdx.z;
    const int Lw = L.w, Lh = L.h, Lpitch = L.pitch, Sh1 = S.h - 1, Spitch = S.pitch;
    pdl_launch_dependents();
    pdl_wait();
    if (word * 4 >= Lw || dy0 >= Lh) return;
    const int4* tx = reinterpret_cast<const int4*>(B.tab + L.tabX + word * 4);      // 4 entries (sx, a0 | a1 << 16), padded
    const int4 t01 = __ldg(tx), t23 = __ldg(tx + 1);
    const int wbase = t01.x >> 2;
    ResizeCol c[4];
    {
        const int sx[4] = {t01.x, t01.z, t23.x, t23.z};
        const int cf[4] = {t01.y, t01.w, t23.y, t23.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int o = sx[k] - 4 * wbase;                 // 0 .. 7 (checked on the host: LevelPlan::fastResize)
            c[k].coef = (unsigned)cf[k];
            c[k].sh = 8u * (o & 3);
            c[k].m = o >= 4 ? 0xffffffffu : 0u;
            // keep the three as live registers: ptxas otherwise re-derives them from `o` in every row
            asm volatile("" : "+r"(c[k].sh), "+r"(c[k].m));
        }
    }
    const uint8_t* sroi = B.pyr + (size_t)frame * P->pyrStride + S.roiOff + 4 * wbase;
    uint8_t* d = B.pyr + (size_t)frame * P->pyrStride + L.roiOff + (size_t)dy0 * Lpitch + 4 * word;
    const int2* ty = B.tab + L.tabY + dy0;
    const int rows = min(PR_ROWS, Lh - dy0);
    int cached = -1, hc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int r = 0; r < PR_ROWS; r++) {
        if (r < rows) {
            const int2 t = __ldg(ty + r);
            const int r0 = min(max(t.x, 0), Sh1), r1 = min(max(t.x + 1, 0), Sh1);
            const int b0 = (int)(t.y << 16), b1 = (int)(t.y & 0xffff0000);      // coefficient << 16: mulhi == (b * h) >> 16
            int h0[4];
            if (r0 == cached) {
#pragma unroll
                for (int k = 0; k < 4; k++) h0[k] = hc[k];
            } else {
                hresize4(reinterpret_cast<const unsigned*>(sroi + (size_t)r0 * Spitch), c, h0);
            }
            if (r1 != r0) hresize4(reinterpret_cast<const unsigned*>(sroi + (size_t)r1 * Spitch), c, hc);
            else {
#pragma unroll
                for (int k = 0; k < 4; k++) hc[k] = h0[k];
            }
            cached = r1;
            unsigned out = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int v = (__mulhi(b0, h0[k]) + __mulhi(b1, hc[k]) + 2) >> 2;      // 0 .. 255 (coefficients sum to 2048)
                out |= (unsigned)v << (8 * k);
            }
            *reinterpret_cast<unsigned*>(d + (size_t)r * Lpitch) = out;
        }
    }
}

// The same kernel with the source rows of the CTA (32 destination rows x 128 destination columns -> about 40 rows x 176 bytes)
// staged in shared memory by one TMA bulk copy per source row: the three words per source row then come from shared
// memory (short scoreboard) instead of L1/L2 (long scoreboard, which is what the kernel above waits on most).
constexpr int PR_SROWS = PR_ROWS <= 8 ? 48 : PR_ROWS * 5 + 8, PR_SPITCH = 208;

__device__ __forceinline__ unsigned vresize4(const int (&h0)[4], const int (&h1)[4], int b0, int b1) {
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int v = (__mulhi(b0, h0[k]) + __mulhi(b1, h1[k]) + 2) >> 2;      // 0 .. 255 (coefficients sum to 2048)
        out |= (unsigned)v << (8 * k);
    }
    return out;
}

// The same with the two products of a pixel pair packed into one word: (b * h) >> 16 is the upper half of a 32-bit product (b <= 2048,
// h < 2^15), so one PRMT packs two of them, the sums of the two rows and the rounding constant add as halves (<= 1022 each) and the
// four result bytes are byte 0 of each half after the shift: 8 IMAD + 5 PRMT + 2 IADD3 + 2 SHF instead of 8 IMAD.HI + 15.  b0, b1 are
// the plain coefficients here (not shifted by 16).
__device__ __forceinline__ unsigned vresize4p(const int (&h0)[4], const int (&h1)[4], unsigned b0, unsigned b1) {
    const unsigned p01 = __byte_perm(b0 * (unsigned)h0[0], b0 * (unsigned)h0[1], 0x7632), p23 = __byte_perm(b0 * (unsigned)h0[2], b0 * (unsigned)h0[3], 0x7632);
    const unsigned q01 = __byte_perm(b1 * (unsigned)h1[0], b1 * (unsigned)h1[1], 0x7632), q23 = __byte_perm(b1 * (unsigned)h1[2], b1 * (unsigned)h1[3], 0x7632);
    const unsigned s01 = (p01 + q01 + 0x00020002u) >> 2, s23 = (p23 + q23 + 0x00020002u) >> 2;
    return __byte_perm(s01, s23, 0x6420);
}

// ROWS = destination rows per thread: PR_ROWS for batches; PR_ROWS_LATENCY for a call with a few frames, where a level is one
// short dependent kernel on the critical path and more, shorter threads finish it sooner.
template <int ROWS, bool PRMT>
__global__ void __launch_bounds__(PR_THREADS, 12) k_pyr_resize_t(const Plan* __restrict__ P, Bufs B, int level) {
    __shared__ __align__(128) uint8_t sSrc[PR_SROWS * PR_SPITCH];
    __shared__ __align__(16) int4 sRow[(PR_THREADS / 32) * ROWS];      // per destination row of the CTA: {r0, r1, b0, b1}
    __shared__ __align__(8) unsigned long long sBar;
    const LevelPlan& L = P->lv[level];
    const LevelPlan& S = P->lv[level - 1];
    const int tid = threadIdx.x;
    const int word = blockIdx.x * 32 + (tid & 31);
    constexpr int rowsPerCta = (PR_THREADS / 32) * ROWS;
    const int dyc = blockIdx.y * rowsPerCta;                              // first destination row of the CTA
    const int dy0 = dyc + (tid >> 5) * ROWS;
    const int frame = blockIdx.z;
    const int Lw = L.w, Lh = L.h, Lpitch = L.pitch, Sh1 = S.h - 1, Spitch = S.pitch;
    pdl_launch_dependents();
    // ---- stage: source rows rs0..rs1, bytes [xs0, xs0 + PR_SPITCH); the row table of the CTA ----
    const int2* tyc = B.tab + L.tabY;
    const int rs0 = min(max(__ldg(tyc + dyc).x, 0), Sh1);
    const int rs1 = min(max(__ldg(tyc + min(dyc + rowsPerCta, Lh) - 1).x + 1, 0), Sh1);
    const int xs0 = __ldg(B.tab + L.tabX + blockIdx.x * 128).x & ~15;
    const uint8_t* sbase = B.pyr + (size_t)frame * P->pyrStride + S.roiOff + xs0;
    if (tid == 0) {
        mbar_init(&sBar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&sBar, (rs1 - rs0 + 1) * PR_SPITCH);
    }
    if (tid < rowsPerCta && dyc + tid < Lh) {
        const int2 t = __ldg(tyc + dyc + tid);
        sRow[tid] = make_int4(min(max(t.x, 0), Sh1) - rs0, min(max(t.x + 1, 0), Sh1) - rs0, (int)(t.y & 0xffff), (int)((unsigned)t.y >> 16));
    }
    __syncthreads();
    pdl_wait();                                                           // (everything above reads the plan's tables only)
    if (tid <= rs1 - rs0) tma_bulk_g2s(sSrc + tid * PR_SPITCH, sbase + (size_t)(rs0 + tid) * Spitch, PR_SPITCH, &sBar);
    if (word * 4 >= Lw || dy0 >= Lh) return;
    const int4* tx = reinterpret_cast<const int4*>(B.tab + L.tabX + word * 4);      // 4 entries (sx, a0 | a1 << 16), padded
    const int4 t01 = __ldg(tx), t23 = __ldg(tx + 1);
    const int wbase = t01.x >> 2;
    ResizeCol c[4];
    {
        const int sx[4] = {t01.x, t01.z, t23.x, t23.z};
        const int cf[4] = {t01.y, t01.w, t23.y, t23.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int o = sx[k] - 4 * wbase;
            c[k].coef = (unsigned)cf[k];
            if (PRMT && k < 3) {
                c[k].sh = (unsigned)o | ((unsigned)(o + 1) << 4);      // byte-permute selector: bytes o, o + 1 of {wa, wb}
                c[k].m = 0u;
                asm volatile("" : "+r"(c[k].sh));
            } else {
                c[k].sh = 8u * (o & 3);
                c[k].m = o >= 4 ? 0xffffffffu : 0u;
                asm volatile("" : "+r"(c[k].sh), "+r"(c[k].m));
            }
        }
    }
    uint8_t* d = B.pyr + (size_t)frame * P->pyrStride + L.roiOff + (size_t)dy0 * Lpitch + 4 * word;
    const int4* rowt = sRow + (tid >> 5) * ROWS;
    const int rows = min(ROWS, Lh - dy0);
    mbar_wait(&sBar, 0);
    // Walk down the SOURCE rows of the strip: every source row gets its horizontal pass exactly once (into ha / hb in turn),
    // and the destination row whose lower source row it is gets emitted right after (the source row index grows with every
    // destination row, so that is at most one).  All branches are warp-uniform.
    int r = 0;
    int4 cur = rowt[0];                                       // {r0, r1, b0, b1} of the next destination row to emit
    int s = cur.x;                                            // newest source row computed
    const uint8_t* sp = sSrc + (4 * wbase - xs0) + s * PR_SPITCH;
    int ha[4], hb[4];
    hresize4s<PRMT>(reinterpret_cast<const unsigned*>(sp), c, ha);
    // emit every destination row whose lower source row is the newest one (`nw`; `pv` = the row before it)
#define PR_EMIT(nw, pv)                                                                                         \
    while (cur.y == s) {                                                                                        \
        *reinterpret_cast<unsigned*>(d) = cur.x == s ? vresize4p(nw, nw, cur.z, cur.w) : vresize4p(pv, nw, cur.z, cur.w); \
        d += Lpitch;                                                                                            \
        if (++r == rows) return;                                                                                \
        cur = rowt[r];                                                                                          \
    }
    while (true) {
        PR_EMIT(ha, hb)
        sp += PR_SPITCH; s++;
        hresize4s<PRMT>(reinterpret_cast<const unsigned*>(sp), c, hb);
        PR_EMIT(hb, ha)
        sp += PR_SPITCH; s++;
        hresize4s<PRMT>(reinterpret_cast<const unsigned*>(sp), c, ha);
    }
#undef PR_EMIT
}

// Apron of every level in one launch.  Work item = one 16-byte chunk of a bordered row that contains apron bytes: all
// chunks of the 19 rows above / below the image, and the chunks at the two ends of every image row.  Each item computes
// its 16 bytes by reflect-101 from final image pixels (copyMakeBorder BORDER_REFLECT_101; |offset| <= 19 < size: one
// fold) and stores them with one 16-byte store; chunk bytes outside the apron (row padding) are don't-care, image
// bytes inside an end chunk are rewritten with their own value.  Words left of the image are a byte-reversed pair of
// image words (one PRMT); words inside the image are copied; only the words that straddle the right edge go byte by byte.
struct ApronTable { int base[ORBB_MAX_LEVELS + 1]; };        // first item of every level (kernel parameter: constant bank)

__global__ void __launch_bounds__(256) k_pyr_apron16(const Plan* __restrict__ P, Bufs B, const ApronTable T, int nlevels) {
    int item = blockIdx.x * 256 + threadIdx.x;
    if (item >= T.base[nlevels]) return;
    const int frame = blockIdx.y;
    int level = 0;
#pragma unroll
    for (int l = 1; l < ORBB_MAX_LEVELS; l++) level += (l < nlevels && item >= T.base[l]);
    const LevelPlan& L = P->lv[level];
    const ApronLevel A = P->apron[level];
    const int w = L.w, h = L.h, pitch = L.pitch;
    item -= A.itemBase;
    // three item classes per level, each row-major (coalesced) and each taking ONE path below in all lanes of a warp:
    // chunks inside the image of the 2 x A.rows apron rows (copy), the left chunks of every bordered row (reflection by PRMT),
    // the right-edge chunks of every bordered row (bytes)
    int by, chunk;                                          // bordered row (0 = top apron row), chunk within the row
    const int nI = 2 * A.rows * A.interiorChunks, nL = A.nLeft * (h + 2 * A.rows);
    if (item < nI) {
        by = A.interiorChunks == 1 ? item : (int)__umulhi((unsigned)item, A.invIC);
        chunk = 2 + item - by * A.interiorChunks;
        if (by >= A.rows) by += h;
    } else if (item < nI + nL) {
        item -= nI;
        by = A.nLeft == 2 ? item >> 1 : item;
        chunk = A.nLeft == 2 ? item & 1 : A.leftChunk0;
    } else {
        item -= nI + nL;
        by = A.nRight == 1 ? item : (int)__umulhi((unsigned)item, A.invNR);
        chunk = A.rightChunk0 + item - by * A.nRight;
    }
    const int iy = by - A.rows;
    const int sy = iy < 0 ? -iy : (iy >= h ? 2 * h - 2 - iy : iy);
    uint8_t* roi = B.pyr + (size_t)frame * P->pyrStride + L.roiOff;
    const uint8_t* srow = roi + (ptrdiff_t)sy * pitch;
    const int x0 = chunk * 16 - kRoiX;                      // image column of the chunk's first byte
    uint4 v;
    if (x0 >= 0 && x0 + 16 <= w) {
        v = *reinterpret_cast<const uint4*>(srow + x0);
    } else {
        unsigned o[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int X = x0 + 4 * q;
            unsigned acc = 0;
            if (X >= 0 && X + 4 <= w) {
                acc = *reinterpret_cast<const unsigned*>(srow + X);
            } else if (X + 3 < 0) {
                if (X >= -20) {                             // bytes X..X+3 <- image columns -X, -X-1, -X-2, -X-3
                    const unsigned* p = reinterpret_cast<const unsigned*>(srow - X);
                    acc = __byte_perm(p[-1], p[0], 0x1234);
                }
            } else if (X < w + kEdge) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int x = X + k;
                    if (x < w + kEdge) acc |= (unsigned)srow[x >= w ? 2 * w - 2 - x : x] << (8 * k);
                }
            }
            o[q] = acc;
        }
        v = make_uint4(o[0], o[1], o[2], o[3]);
    }
    *reinterpret_cast<uint4*>(roi + (ptrdiff_t)iy * pitch + x0) = v;
}

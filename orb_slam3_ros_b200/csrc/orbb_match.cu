// liborbb200.so -- Hamming matching and rectified-stereo matching for B200 (sm_100a).
//
//   k_knn2_partial / k_knn2_merge   cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2)   reference Frame.cc:1144
//   k_ratio_test                    Lowe ratio                                       Frame.cc:1151
//   k_best2_csr                     best / second-best candidate scans               ORBmatcher.cc:77-120 (and siblings)
//   k_stereo_match / k_stereo_cut   Frame::ComputeStereoMatches                      Frame.cc:811-981
//
// The 2-NN kernel is integer work on the CUDA cores (XOR + POPC over 32-bit lanes, no tensor cores): queries live in
// registers (3 per thread), database tiles are staged in shared memory with 1-D TMA bulk copies (cp.async.bulk +
// mbarrier, 3 stages) and every thread of a warp reads the same database row (shared-memory broadcast).  The direct
// form (8 POPC per pair) is bound by the POPC pipe (16 lanes/clk/SM); a carry-save compressor in LOP3s first
// (hamming_c332: 5 POPC + 14 LOP3 per pair) balances the POPC and integer-ALU pipes and is 1.56x faster.
#include <limits.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "orbb_internal.cuh"

struct orbb_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    std::string err;
    // scratch
    unsigned* partial = nullptr; size_t partialCap = 0;     // [chunk][nq][2] packed (dist<<22 | local index)
    void* scratch[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratchCap[4] = {0, 0, 0, 0};
    void* frame[ORBB_FRAME_SLOTS] = {};                     // orbb_frame_upload: frames kept on the device between scans
    size_t frameCap[ORBB_FRAME_SLOTS] = {};
};

namespace orbb {

static int m_err(orbb_matcher* m, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (m) m->err = buf;
    g_lastError = buf;
    return code;
}

#define ORBM_CUDA(m, call)                                                                                   \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return orbb::m_err(m, ORBB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// brute-force Hamming 2-NN
// ------------------------------------------------------------------------------------------------
constexpr int KNN_THREADS = 256;
constexpr int KNN_ROWS = 256;                      // database rows per shared-memory tile (8 KB)
constexpr int KNN_STAGES = 3;
constexpr unsigned KNN_IDX_BITS = 22;              // packed key = dist << 22 | chunk-local index
constexpr unsigned KNN_NONE = 0xffffffffu;

__device__ __forceinline__ void top2_insert(unsigned& k1, unsigned& k2, unsigned k) {
    const unsigned hi = max(k1, k);
    k1 = min(k1, k);
    k2 = min(k2, hi);
}

// Hamming distance, the direct way: 8 XOR + 8 POPC (+ adds).  POPC issues at 16 lanes/clk/SM on B200, a quarter of
// the integer-ALU rate, so a kernel made only of this is POPC-pipe bound at 2 pairs/clk/SM.
__device__ __forceinline__ int hamming_popc(const unsigned (&q)[8], const uint4 a, const uint4 b) {
    return __popc(q[0] ^ a.x) + __popc(q[1] ^ a.y) + __popc(q[2] ^ a.z) + __popc(q[3] ^ a.w) + __popc(q[4] ^ b.x) +
           __popc(q[5] ^ b.y) + __popc(q[6] ^ b.z) + __popc(q[7] ^ b.w);
}

// The same distance through a carry-save adder tree (Harley-Seal): the eight XOR words are compressed bit-slice-wise
// into ones / twos / fours / eights words with 3-input LOP3s (sum = a^b^c, carry = majority), so only 4 POPC remain:
// d = popc(ones) + 2 popc(twos) + 4 popc(fours) + 8 popc(eights).  It trades POPC-pipe work for ALU-pipe work; mixing
// both forms per thread balances the two pipes (about two CSA pairs per direct pair).
__device__ __forceinline__ unsigned maj3(unsigned a, unsigned b, unsigned c) {      // carry of a full adder: one LOP3
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ unsigned xor3(unsigned a, unsigned b, unsigned c) {      // sum of a full adder: one LOP3
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ int hamming_csa(const unsigned (&q)[8], const uint4 a, const uint4 b) {
    const unsigned x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x2 = q[2] ^ a.z, x3 = q[3] ^ a.w;
    const unsigned x4 = q[4] ^ b.x, x5 = q[5] ^ b.y, x6 = q[6] ^ b.z, x7 = q[7] ^ b.w;
    const unsigned o1 = xor3(x0, x1, x2), t1 = maj3(x0, x1, x2);
    const unsigned o2 = xor3(x3, x4, x5), t2 = maj3(x3, x4, x5);
    const unsigned o3 = xor3(o1, o2, x6), t3 = maj3(o1, o2, x6);
    const unsigned ones = o3 ^ x7, t4 = o3 & x7;
    const unsigned w1 = xor3(t1, t2, t3), f1 = maj3(t1, t2, t3);
    const unsigned twos = w1 ^ t4, f2 = w1 & t4;
    const unsigned fours = f1 ^ f2, eights = f1 & f2;
    return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
}

// Shallower compressors: fewer LOP3s per pair at the price of one or two more POPCs.
//   hamming_c71: 7 words -> 3 bit-slices with 4 full adders, the 8th word counted directly (16 LOP3, 4 POPC)
//   hamming_c332: full adders on (x0,x1,x2), (x3,x4,x5), (s1,s2,x6); x7 direct (14 LOP3, 5 POPC) -- close to the
//                 ALU : POPC = 4 : 1 balance of the two pipes on its own
__device__ __forceinline__ int hamming_c71(const unsigned (&q)[8], const uint4 a, const uint4 b) {
    const unsigned x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x2 = q[2] ^ a.z, x3 = q[3] ^ a.w;
    const unsigned x4 = q[4] ^ b.x, x5 = q[5] ^ b.y, x6 = q[6] ^ b.z, x7 = q[7] ^ b.w;
    const unsigned s1 = xor3(x0, x1, x2), c1 = maj3(x0, x1, x2);
    const unsigned s2 = xor3(x3, x4, x5), c2 = maj3(x3, x4, x5);
    const unsigned ones = xor3(s1, s2, x6), c3 = maj3(s1, s2, x6);
    const unsigned twos = xor3(c1, c2, c3), fours = maj3(c1, c2, c3);
    return __popc(ones) + __popc(x7) + 2 * __popc(twos) + 4 * __popc(fours);
}
__device__ __forceinline__ int hamming_c332(const unsigned (&q)[8], const uint4 a, const uint4 b) {
    const unsigned x0 = q[0] ^ a.x, x1 = q[1] ^ a.y, x2 = q[2] ^ a.z, x3 = q[3] ^ a.w;
    const unsigned x4 = q[4] ^ b.x, x5 = q[5] ^ b.y, x6 = q[6] ^ b.z, x7 = q[7] ^ b.w;
    const unsigned s1 = xor3(x0, x1, x2), c1 = maj3(x0, x1, x2);
    const unsigned s2 = xor3(x3, x4, x5), c2 = maj3(x3, x4, x5);
    const unsigned ones = xor3(s1, s2, x6), c3 = maj3(s1, s2, x6);
    return __popc(ones) + __popc(x7) + 2 * (__popc(c1) + __popc(c2) + __popc(c3));
}
template <int METHOD>
__device__ __forceinline__ int hamming_by(const unsigned (&q)[8], const uint4 a, const uint4 b) {
    return METHOD == 0 ? hamming_popc(q, a, b) : METHOD == 1 ? hamming_csa(q, a, b) : METHOD == 2 ? hamming_c71(q, a, b) : hamming_c332(q, a, b);
}

// grid (query tiles, database chunks).  A chunk is < 2^22 rows; packed keys make "lowest index wins ties" the
// natural order of an unsigned min.  NC queries per thread use the CSA form, NS the direct form.
template <int NC, int NS, int MC = 1, int MS = 0>
__global__ void __launch_bounds__(KNN_THREADS) k_knn2_partial(const uint4* __restrict__ q, int nq, const uint4* __restrict__ db,
                                                             long long nd, int chunkRows, unsigned* __restrict__ partial) {
    constexpr int QPT = NC + NS;
    constexpr int QTILE = KNN_THREADS * QPT;
    __shared__ __align__(128) uint4 sDb[KNN_STAGES][KNN_ROWS * 2];
    __shared__ __align__(8) unsigned long long sBar[KNN_STAGES];
    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.y * chunkRows;
    const int rows = (int)min((long long)chunkRows, nd - row0);
    const int ntiles = (rows + KNN_ROWS - 1) / KNN_ROWS;
    if (tid == 0) {
        for (int s = 0; s < KNN_STAGES; s++) mbar_init(&sBar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int s = t % KNN_STAGES;
        const int r = min(KNN_ROWS, rows - t * KNN_ROWS);
        mbar_expect_tx(&sBar[s], (unsigned)r * 32u);
        tma_bulk_g2s(&sDb[s][0], db + (row0 + (long long)t * KNN_ROWS) * 2, (unsigned)r * 32u, &sBar[s]);
    };
    if (tid == 0)
        for (int t = 0; t < KNN_STAGES && t < ntiles; t++) issue(t);

    unsigned qa[QPT][8];
#pragma unroll
    for (int k = 0; k < QPT; k++) {
        const int qi = min(blockIdx.x * QTILE + k * KNN_THREADS + tid, nq - 1);
        const uint4 a = __ldg(q + (size_t)qi * 2), b = __ldg(q + (size_t)qi * 2 + 1);
        qa[k][0] = a.x; qa[k][1] = a.y; qa[k][2] = a.z; qa[k][3] = a.w;
        qa[k][4] = b.x; qa[k][5] = b.y; qa[k][6] = b.z; qa[k][7] = b.w;
    }
    unsigned k1[QPT], k2[QPT];
#pragma unroll
    for (int k = 0; k < QPT; k++) k1[k] = k2[k] = KNN_NONE;

    for (int t = 0; t < ntiles; t++) {
        const int s = t % KNN_STAGES;
        mbar_wait(&sBar[s], (unsigned)(t / KNN_STAGES) & 1u);
        const int r = min(KNN_ROWS, rows - t * KNN_ROWS);
        const uint4* tile = &sDb[s][0];
        const unsigned jbase = (unsigned)(t * KNN_ROWS);
#pragma unroll 2
        for (int j = 0; j < r; j++) {
            const uint4 a = tile[2 * j], b = tile[2 * j + 1];
#pragma unroll
            for (int k = 0; k < QPT; k++) {
                const int d = k < NC ? hamming_by<MC>(qa[k], a, b) : hamming_by<MS>(qa[k], a, b);
                top2_insert(k1[k], k2[k], ((unsigned)d << KNN_IDX_BITS) + (jbase + (unsigned)j));
            }
        }
        __syncthreads();                               // everyone is done with stage s
        if (tid == 0 && t + KNN_STAGES < ntiles) issue(t + KNN_STAGES);
    }
#pragma unroll
    for (int k = 0; k < QPT; k++) {
        const int qi = blockIdx.x * QTILE + k * KNN_THREADS + tid;
        if (qi < nq) {
            unsigned* o = partial + ((size_t)blockIdx.y * nq + qi) * 2;
            o[0] = k1[k];
            o[1] = k2[k];
        }
    }
}

// merge per-chunk packed results into (idx, dist) pairs; chunks are visited in ascending database order so a strict
// '<' keeps the lowest global index among equal distances.
__global__ void k_knn2_merge_chunks(const unsigned* __restrict__ partial, int nchunks, int nq, int chunkRows, int indexBase,
                                    int* __restrict__ idx2, int* __restrict__ dist2) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    int d1 = INT_MAX, d2 = INT_MAX, i1 = -1, i2 = -1;
    for (int c = 0; c < nchunks; c++) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const unsigned p = partial[((size_t)c * nq + qi) * 2 + k];
            if (p == KNN_NONE) continue;
            const int d = (int)(p >> KNN_IDX_BITS);
            const int i = indexBase + c * chunkRows + (int)(p & ((1u << KNN_IDX_BITS) - 1));
            if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = i; }
            else if (d < d2) { d2 = d; i2 = i; }
        }
    }
    idx2[2 * qi] = i1; idx2[2 * qi + 1] = i2;
    dist2[2 * qi] = d1; dist2[2 * qi + 1] = d2;
}

// merge across database shards: input [shard][nq][2] (idx, dist), lexicographic (dist, idx) order
// (shardStride = ints between the blocks of consecutive shards: 2 * nq for separate idx / dist arrays, 4 * nq for the packed
// [idx][dist] blocks that orbb_knn2_sharded all-gathers)
__global__ void k_knn2_merge_shards(const int* __restrict__ idxSh, const int* __restrict__ distSh, size_t shardStride, int nshards, int nq,
                                    int* __restrict__ idx2, int* __restrict__ dist2) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    int d1 = INT_MAX, d2 = INT_MAX, i1 = -1, i2 = -1;
    for (int s = 0; s < nshards; s++) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int i = idxSh[(size_t)s * shardStride + (size_t)qi * 2 + k];
            const int d = distSh[(size_t)s * shardStride + (size_t)qi * 2 + k];
            if (i < 0) continue;
            if (d < d1 || (d == d1 && i < i1)) { d2 = d1; i2 = i1; d1 = d; i1 = i; }
            else if (d < d2 || (d == d2 && i < i2)) { d2 = d; i2 = i; }
        }
    }
    idx2[2 * qi] = i1; idx2[2 * qi + 1] = i2;
    dist2[2 * qi] = d1; dist2[2 * qi + 1] = d2;
}

__global__ void k_ratio_test(const int* __restrict__ idx2, const int* __restrict__ dist2, int nq, double ratio, uint8_t* __restrict__ keep) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    // (*it).size() >= 2 && (*it)[0].distance < (*it)[1].distance * 0.7   (DMatch::distance is a float)
    const bool two = idx2[2 * qi] >= 0 && idx2[2 * qi + 1] >= 0;
    keep[qi] = two && (double)(float)dist2[2 * qi] < (double)(float)dist2[2 * qi + 1] * ratio;
}

// ------------------------------------------------------------------------------------------------
// best / second-best over candidate lists, one warp per query
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
           __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(256) k_best2_csr(const uint4* __restrict__ q, int nq, const uint4* __restrict__ train,
                                                  const int* __restrict__ cand, const int* __restrict__ rowptr, int init,
                                                  int* __restrict__ out4) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const uint4 a0 = __ldg(q + (size_t)qi * 2), a1 = __ldg(q + (size_t)qi * 2 + 1);
    const int beg = rowptr[qi], end = rowptr[qi + 1];
    unsigned long long k1 = ~0ull, k2 = ~0ull;            // dist << 32 | position in the list
    for (int c = beg + lane; c < end; c += 32) {
        const int t = cand[c];
        const int d = hamming256(a0, a1, __ldg(train + (size_t)t * 2), __ldg(train + (size_t)t * 2 + 1));
        if (d < init) {
            const unsigned long long k = ((unsigned long long)(unsigned)d << 32) | (unsigned)(c - beg);
            const unsigned long long hi = max(k1, k);
            k1 = min(k1, k);
            k2 = min(k2, hi);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, off), o2 = __shfl_xor_sync(0xffffffffu, k2, off);
        const unsigned long long lo = min(k1, o1), hi = max(k1, o1);
        k2 = min(min(k2, o2), hi);
        k1 = lo;
    }
    if (lane == 0) {
        int* o = out4 + (size_t)qi * 4;
        o[0] = k1 == ~0ull ? init : (int)(k1 >> 32);
        o[1] = k1 == ~0ull ? -1 : cand[beg + (int)(k1 & 0xffffffffu)];
        o[2] = k2 == ~0ull ? init : (int)(k2 >> 32);
        o[3] = k2 == ~0ull ? -1 : cand[beg + (int)(k2 & 0xffffffffu)];
    }
}

// ------------------------------------------------------------------------------------------------
// Frame::GetFeaturesInArea (Frame.cc:657-723) fused with the best / second-best scan of ORBmatcher::SearchByProjection
// (ORBmatcher.cc:71-120): one warp per query.  No grid is materialised: a keypoint's grid cell (PosInGrid, :725-735) is a
// pure function of its coordinates, so every lane tests keypoints directly against the query's cell window, level range
// and radius.  The reference visits candidates cell by cell (ix, then iy, then keypoint index) and keeps the FIRST minimum;
// the same tie rule falls out of an unsigned min over  dist << 32 | ix << 26 | iy << 20 | index.
// ------------------------------------------------------------------------------------------------
// K = candidates kept per query, in the order of the reference's scan (distance, then first seen): K = 2 is the best / second-best
// pair of :77-120; K = 4 lets the host apply the reference's in-order decisions (a key point matched by an earlier map point of
// the same call is skipped by the later ones, :88-90) without going back to the device -- dropping taken candidates from a longer
// sorted list leaves the same first two.  Key point coordinates / octaves are read with byte strides, so the arrays may be the
// extractor's own orbb_keypoint records (x, y at offset 0, octave at offset 20, stride 24).
template <int K>
__global__ void __launch_bounds__(256) k_search_area_topk(const uint8_t* __restrict__ kps, size_t kpsStride, const uint8_t* __restrict__ oct,
                                                         size_t octStride, const uint4* __restrict__ train, int n, float minX, float minY,
                                                         float invW, float invH, const float4* __restrict__ queries,
                                                         const int2* __restrict__ qlev, const uint4* __restrict__ qdesc, int nq,
                                                         const uint8_t* __restrict__ skip, const float* __restrict__ uRight, int init,
                                                         int* __restrict__ out) {
    const int qi = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const float4 q = __ldg(queries + qi);                      // x, y, r, projXR
    const int2 lev = __ldg(qlev + qi);
    const float x = q.x, y = q.y, r = q.z;
    unsigned long long top[K];                                 // this lane's K smallest keys, ascending
#pragma unroll
    for (int k = 0; k < K; k++) top[k] = ~0ull;
    const int nMinCellX = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, minX), r), invW)));
    const int nMaxCellX = min(63, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, minX), r), invW)));
    const int nMinCellY = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, minY), r), invH)));
    const int nMaxCellY = min(47, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, minY), r), invH)));
    if (nMinCellX < 64 && nMaxCellX >= 0 && nMinCellY < 48 && nMaxCellY >= 0) {
        const bool bCheckLevels = lev.x > 0 || lev.y >= 0;
        const uint4 a0 = __ldg(qdesc + (size_t)qi * 2), a1 = __ldg(qdesc + (size_t)qi * 2 + 1);
        for (int j = lane; j < n; j += 32) {
            const float2 kp = __ldg(reinterpret_cast<const float2*>(kps + (size_t)j * kpsStride));
            const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, minX), invW)), py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, minY), invH));
            if (px < nMinCellX || px > nMaxCellX || py < nMinCellY || py > nMaxCellY) continue;   // also drops cells outside the grid
            if (bCheckLevels) {
                const int o = __ldg(reinterpret_cast<const int*>(oct + (size_t)j * octStride));
                if (o < lev.x || (lev.y >= 0 && o > lev.y)) continue;
            }
            if (!(fabsf(__fsub_rn(kp.x, x)) < r && fabsf(__fsub_rn(kp.y, y)) < r)) continue;
            if (skip && skip[j]) continue;
            if (uRight) {
                const float ur = __ldg(uRight + j);
                if (ur > 0 && fabsf(__fsub_rn(q.w, ur)) > r) continue;
            }
            const int d = hamming256(a0, a1, __ldg(train + (size_t)j * 2), __ldg(train + (size_t)j * 2 + 1));
            if (d < init) {
                unsigned long long k = ((unsigned long long)(unsigned)d << 32) | ((unsigned long long)px << 26) |
                                       ((unsigned long long)py << 20) | (unsigned)j;
#pragma unroll
                for (int t = 0; t < K; t++) {                  // insertion into the sorted K
                    const unsigned long long lo = min(top[t], k);
                    k = max(top[t], k);
                    top[t] = lo;
                }
            }
        }
    }
    // K rounds: the smallest head of all lanes is the next entry; its owner advances (keys are unique: they contain the index)
    int* o = out + (size_t)qi * 2 * K;
#pragma unroll
    for (int round = 0; round < K; round++) {
        unsigned long long m = top[0];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, off));
        if (top[0] == m && m != ~0ull) {
#pragma unroll
            for (int t = 0; t + 1 < K; t++) top[t] = top[t + 1];
            top[K - 1] = ~0ull;
        }
        if (lane == 0) {
            o[2 * round] = m == ~0ull ? init : (int)(m >> 32);
            o[2 * round + 1] = m == ~0ull ? -1 : (int)(m & 0xfffffu);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:329-403): one warp per map point.  Lane i owns row i of the
// N x N distance matrix (rows beyond 32 in further passes); the median sorted[(size_t)(0.5*(N-1))] of a row is found by
// bisection on the distance value (count of entries <= v), so no matrix and no sort are needed.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_distinctive(const uint4* __restrict__ desc, const int* __restrict__ rowptr, int ngroups,
                                                    int* __restrict__ best) {
    const int g = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= ngroups) return;
    const int beg = rowptr[g], N = rowptr[g + 1] - beg;
    if (N <= 0) {
        if (lane == 0) best[g] = -1;
        return;
    }
    const uint4* d = desc + (size_t)beg * 2;
    const int kth = (int)(0.5 * (N - 1));                     // :392
    unsigned bestKey = 0xffffffffu;                           // median << 16 | row  (first minimum = smallest key)
    for (int base = 0; base < N; base += 32) {
        const int i = base + lane;
        if (i < N) {
            const uint4 a0 = __ldg(d + 2 * i), a1 = __ldg(d + 2 * i + 1);
            int lo = 0, hi = 256;                             // smallest v with #{j : d_ij <= v} > kth
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                int cnt = 0;
                for (int j = 0; j < N; j++) cnt += hamming256(a0, a1, __ldg(d + 2 * j), __ldg(d + 2 * j + 1)) <= mid;
                if (cnt > kth) hi = mid; else lo = mid + 1;
            }
            bestKey = min(bestKey, ((unsigned)lo << 16) | (unsigned)i);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) bestKey = min(bestKey, __shfl_xor_sync(0xffffffffu, bestKey, off));
    if (lane == 0) best[g] = (int)(bestKey & 0xffffu);
}

// ------------------------------------------------------------------------------------------------
// Frame::ComputeStereoMatches (Frame.cc:811-981): one warp per left keypoint
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// Rotation-consistency filter of the match scans (reference ORBmatcher.cc:345-352 and siblings + ComputeThreeMaxima,
// :2012-2053): every match votes for the bin round((angle_a - angle_b [+360]) * (1/HISTO_LENGTH)) of a 30-bin histogram, the three
// fullest bins are found by the reference's single pass (later bins do not displace earlier equal ones; second / third are
// dropped when below 10 % of the first) and the matches of all other bins are discarded.  One CTA per match set.
__global__ void __launch_bounds__(256) k_rotation_check(const float* __restrict__ angA, const float* __restrict__ angB,
                                                        const int* __restrict__ rowptr, uint8_t* __restrict__ keep, int* __restrict__ ind3) {
    __shared__ int sHist[30];
    __shared__ int sInd[3];
    const int set = blockIdx.x, tid = threadIdx.x;
    const int beg = rowptr[set], end = rowptr[set + 1];
    if (tid < 30) sHist[tid] = 0;
    __syncthreads();
    const float factor = 1.0f / 30;                              // :236
    auto bin_of = [&](int i) {
        float rot = __fsub_rn(angA[i], angB[i]);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, factor));
        if (bin == 30) bin = 0;
        return min(max(bin, 0), 29);                             // (the reference asserts the range)
    };
    for (int i = beg + tid; i < end; i += 256) atomicAdd(&sHist[bin_of(i)], 1);
    __syncthreads();
    if (tid == 0) {
        int max1 = 0, max2 = 0, max3 = 0, i1 = -1, i2 = -1, i3 = -1;
        for (int i = 0; i < 30; i++) {
            const int sz = sHist[i];
            if (sz > max1) { max3 = max2; max2 = max1; max1 = sz; i3 = i2; i2 = i1; i1 = i; }
            else if (sz > max2) { max3 = max2; max2 = sz; i3 = i2; i2 = i; }
            else if (sz > max3) { max3 = sz; i3 = i; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { i2 = -1; i3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) i3 = -1;
        sInd[0] = i1; sInd[1] = i2; sInd[2] = i3;
        ind3[3 * set] = i1; ind3[3 * set + 1] = i2; ind3[3 * set + 2] = i3;
    }
    __syncthreads();
    const int i1 = sInd[0], i2 = sInd[1], i3 = sInd[2];
    for (int i = beg + tid; i < end; i += 256) {
        const int b = bin_of(i);
        keep[i] = b == i1 || b == i2 || b == i3;
    }
}

// "Next" row: Frame::UndistortKeyPoints (reference orb_slam3/src/Frame.cc:747-780) = cv::undistortPoints(pts, pts, K, D,
// noArray(), K) over the N keypoints: OpenCV's cvUndistortPointsInternal with its default criteria (5 fixed-point
// iterations, no epsilon test), in double precision without fused multiply-adds (this file is built with -fmad=false),
// result rounded to float.  k[0..13] = distortion coefficients (k1 k2 p1 p2 k3 k4 k5 k6 s1 s2 s3 s4 taux tauy), zero-filled;
// the tilt terms must be zero (ORB-SLAM3's pinhole model has 4 or 5 coefficients).
struct UndistortParams { double fx, fy, cx, cy, nfx, nfy, ncx, ncy, k[12]; };

__device__ __forceinline__ float2 undistort_point(float px, float py, const UndistortParams& p) {
    const double u = (double)px, v = (double)py;
    const double ifx = 1. / p.fx, ify = 1. / p.fy;
    double x = (u - p.cx) * ifx, y = (v - p.cy) * ify;
    const double x0 = x, y0 = y;
    const double* k = p.k;
    for (int j = 0; j < 5; j++) {
        const double r2 = x * x + y * y;
        const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) { x = (u - p.cx) * ifx; y = (v - p.cy) * ify; break; }
        const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
        const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
        x = (x0 - deltaX) * icdist;
        y = (y0 - deltaY) * icdist;
    }
    // RR = P (new camera matrix), R = identity: xx = fx'*x + 0*y + cx', ww = 1 / (0*x + 0*y + 1)
    const double xx = p.nfx * x + 0. * y + p.ncx, yy = 0. * x + p.nfy * y + p.ncy, ww = 1. / (0. * x + 0. * y + 1.);
    return make_float2((float)(xx * ww), (float)(yy * ww));
}

__global__ void __launch_bounds__(256) k_undistort(const float2* __restrict__ in, float2* __restrict__ out, int n, const UndistortParams p) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const float2 pt = in[i];
    out[i] = undistort_point(pt.x, pt.y, p);
}

// "Next" row (RGB-D counterpart of S1): Frame::ComputeStereoFromRGBD (reference orb_slam3/src/Frame.cc:984-1005) over the
// keypoints of the extractor's last batch: d = imDepth.at<float>(v, u) at the (distorted) keypoint, mvDepth = d and
// mvuRight = kpU.x - mbf / d when d > 0, else -1 / -1; kpU = the undistorted keypoint (Frame::UndistortKeyPoints, :747-780;
// dist[0] == 0: kpU = kp).  A 16-bit depth map is scaled on the fly like Tracking::GrabImageRGBD does with
// imDepth.convertTo(imDepth, CV_32F, mDepthMapFactor) (Tracking.cc:1576-1577: float(src) * float(alpha), one rounding).
__global__ void __launch_bounds__(256) k_rgbd_stereo(const Plan* __restrict__ P, Bufs B, const uint8_t* __restrict__ depth, int isU16,
                                                     float factor, size_t rowStride, size_t frameStride, const UndistortParams p,
                                                     int undist, float mbf) {
    const int frame = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int n = B.outCount[frame * 2];
    if (i >= n) return;
    const size_t fo = (size_t)frame * P->kpCap;
    const orbb_keypoint kp = B.kps[fo + i];
    const int u = (int)kp.x, v = (int)kp.y;                  // Mat::at<float>(float, float): the indices convert to int
    const uint8_t* row = depth + (size_t)frame * frameStride + (size_t)v * rowStride;
    const float d = isU16 ? __fmul_rn((float)reinterpret_cast<const unsigned short*>(row)[u], factor) : reinterpret_cast<const float*>(row)[u];
    float ur = -1.f, dp = -1.f;
    if (d > 0) {
        const float xu = undist ? undistort_point(kp.x, kp.y, p).x : kp.x;
        dp = d;
        ur = __fsub_rn(xu, __fdiv_rn(mbf, d));
    }
    B.uRight[fo + i] = ur;
    B.depth[fo + i] = dp;
}

// Right-image keypoints bucketed by row (counting sort on (int)y, one CTA per frame): the reference's vRowIndices table
// (Frame.cc:824-838) lists for every row the right keypoints whose band [y - r, y + r] covers it; here a left keypoint
// scans the buckets of the rows within the largest band radius of its own row and applies the exact band test per
// candidate.  The candidate set is the same, and the packed (distance, index) minimum does not depend on the visiting order.
__global__ void __launch_bounds__(256) k_stereo_rows(const Plan* __restrict__ P, Bufs BR) {
    extern __shared__ int sRow[];                            // [H + 1] counts -> starts, [H + 1] running positions
    const int frame = blockIdx.x, tid = threadIdx.x;
    const int H = P->H, nR = BR.outCount[frame * 2];
    int* cnt = sRow;
    int* pos = sRow + H + 1;
    __shared__ int sPart[256];
    const size_t fo = (size_t)frame * P->kpCap;
    const orbb_keypoint* kR = BR.kps + fo;
    for (int i = tid; i <= H; i += 256) cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < nR; i += 256) atomicAdd(&cnt[min(max((int)kR[i].y, 0), H - 1)], 1);
    __syncthreads();
    // exclusive scan of cnt[0..H]: each thread owns a contiguous chunk
    const int chunk = (H + 1 + 255) / 256, c0 = tid * chunk, c1 = min(c0 + chunk, H + 1);
    int sum = 0;
    for (int i = c0; i < c1; i++) sum += cnt[i];
    sPart[tid] = sum;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        const int v = tid >= off ? sPart[tid - off] : 0;
        __syncthreads();
        sPart[tid] += v;
        __syncthreads();
    }
    int run = sPart[tid] - sum;
    int* rowStart = BR.stRowStart + (size_t)frame * (H + 2);
    for (int i = c0; i < c1; i++) {
        const int c = cnt[i];
        pos[i] = run;
        rowStart[i] = run;
        run += c;
    }
    if (tid == 255) rowStart[H + 1] = sPart[255];
    __syncthreads();
    for (int i = tid; i < nR; i += 256) {
        const orbb_keypoint kp = kR[i];
        const float r = __fmul_rn(2.0f, P->lv[kp.octave].scale);                           // :832
        const int maxr = (int)ceilf(__fadd_rn(kp.y, r)), minr = (int)floorf(__fsub_rn(kp.y, r));
        const int slot = atomicAdd(&pos[min(max((int)kp.y, 0), H - 1)], 1);
        BR.stRec[fo + slot] = make_uint4(__float_as_uint(kp.x), ((unsigned)minr & 0xffffu) | ((unsigned)maxr << 16), (unsigned)kp.octave, (unsigned)i);
    }
}

__global__ void __launch_bounds__(256) k_stereo_match(const Plan* __restrict__ P, Bufs BL, Bufs BR, float mbf, float mb, int useRows) {
    const int frame = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int iL = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int nL = BL.outCount[frame * 2], nR = BR.outCount[frame * 2];
    if (iL >= nL) return;
    const size_t fo = (size_t)frame * P->kpCap;
    float* uRightOut = BL.uRight + fo;
    float* depthOut = BL.depth + fo;
    int* bestROut = BL.bestR + fo;
    int* sadOut = BL.sad + fo;
    float resU = -1.f, resD = -1.f;
    int resSad = -1, resR = -1;

    const orbb_keypoint kpL = BL.kps[fo + iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const float minZ = mb, minD = 0.f, maxD = __fdiv_rn(mbf, minZ);                       // :841-843
    const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);
    const int rowL = (int)vL;                                                              // vRowIndices[vL] :856
    const int nRows = P->lv[0].h;
    bool alive = !(maxU < 0) && rowL >= 0 && rowL < nRows;                                 // :864

    // ---- best Hamming candidate, ascending iR, strict '<', start at TH_HIGH (:867-894) ----
    unsigned best = (100u << 16) | 0xffffu;          // dist << 16 | iR  (iR < 65535)
    if (alive) {
        const uint4* dl = reinterpret_cast<const uint4*>(BL.desc + (fo + iL) * 32);
        const uint4 a0 = dl[0], a1 = dl[1];
        if (useRows) {
            // row buckets within the largest band radius (+2: the ceil / floor of the band ends) of rowL
            const int R = (int)ceilf(2.0f * P->lv[P->nlevels - 1].scale) + 2;
            const int* rowStart = BR.stRowStart + (size_t)frame * (nRows + 2);
            const int beg = rowStart[max(rowL - R, 0)], end = rowStart[min(rowL + R, nRows - 1) + 1];
            const uint4* rec = BR.stRec + fo;
            for (int base = beg; base < end; base += 32) {
                if (base + lane < end) {
                    const uint4 rc = __ldg(rec + base + lane);
                    const int minr = (int)(short)(rc.y & 0xffffu), maxr = (int)rc.y >> 16, oct = (int)rc.z;
                    const float x = __uint_as_float(rc.x);
                    if (rowL >= minr && rowL <= maxr && oct >= levelL - 1 && oct <= levelL + 1 && x >= minU && x <= maxU) {
                        const uint4* dr = reinterpret_cast<const uint4*>(BR.desc + (fo + rc.w) * 32);
                        const int d = hamming256(a0, a1, dr[0], dr[1]);
                        if (d < 100) best = min(best, ((unsigned)d << 16) | rc.w);
                    }
                }
            }
        } else {
        const orbb_keypoint* kR = BR.kps + fo;
        for (int base = 0; base < nR; base += 32) {
            const int iR = base + lane;
            if (iR < nR) {
                const orbb_keypoint kp = kR[iR];
                const float r = __fmul_rn(2.0f, P->lv[kp.octave].scale);                   // :832
                const int maxr = (int)ceilf(__fadd_rn(kp.y, r)), minr = (int)floorf(__fsub_rn(kp.y, r));
                if (rowL >= minr && rowL <= maxr && kp.octave >= levelL - 1 && kp.octave <= levelL + 1 && kp.x >= minU &&
                    kp.x <= maxU) {
                    const uint4* dr = reinterpret_cast<const uint4*>(BR.desc + (fo + iR) * 32);
                    const int d = hamming256(a0, a1, dr[0], dr[1]);
                    if (d < 100) best = min(best, ((unsigned)d << 16) | (unsigned)iR);
                }
            }
        }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
    const int bestDist = (int)(best >> 16);
    const int bestIdxR = (int)(best & 0xffffu);

    if (alive && bestDist < 75) {                                                          // thOrbDist :816,:897
        resR = bestIdxR;
        const float uR0 = BR.kps[fo + bestIdxR].x;
        const LevelPlan& L = P->lv[levelL];
        const float sf = 1.0f / L.scale;                                                   // mvInvScaleFactors
        const float scaleduL = roundf(__fmul_rn(kpL.x, sf));
        const float scaledvL = roundf(__fmul_rn(kpL.y, sf));
        const float scaleduR0 = roundf(__fmul_rn(uR0, sf));
        const int w = 5, LL = 5;
        const float iniu = scaleduR0 + LL - w, endu = scaleduR0 + LL + w + 1;
        if (!(iniu < 0 || endu >= (float)L.w)) {                                           // :918
            const uint8_t* pl = BL.pyr + (size_t)frame * P->pyrStride + L.roiOff;
            const uint8_t* pr = BR.pyr + (size_t)frame * P->pyrStride + L.roiOff;
            const int y0 = (int)(scaledvL - w), xl0 = (int)(scaleduL - w), xr00 = (int)(scaleduR0 - w);
            int lv[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int p = lane + 32 * k;
                lv[k] = p < 121 ? pl[(size_t)(y0 + p / 11) * L.pitch + xl0 + p % 11] : 0;
            }
            int bestSad = INT_MAX, bestinc = 0;
            float dists[11];
#pragma unroll
            for (int inc = -5; inc <= 5; inc++) {                                          // :921-933
                int sad = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int p = lane + 32 * k;
                    if (p < 121) sad += abs(lv[k] - (int)pr[(size_t)(y0 + p / 11) * L.pitch + xr00 + inc + p % 11]);
                }
                sad = warp_sum_i(sad);
                const float dist = (float)sad;
                if (dist < (float)bestSad) { bestSad = sad; bestinc = inc; }
                dists[inc + 5] = dist;
            }
            if (!(bestinc == -LL || bestinc == LL)) {                                      // :935
                float d1 = 0, d2 = 0, d3 = 0;
#pragma unroll
                for (int k = 1; k < 10; k++)
                    if (k == bestinc + 5) { d1 = dists[k - 1]; d2 = dists[k]; d3 = dists[k + 1]; }
                const float deltaR = __fdiv_rn(__fsub_rn(d1, d3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2))));   // :943
                if (!(deltaR < -1 || deltaR > 1)) {
                    float bestuR = __fmul_rn(L.scale, __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));   // :949
                    float disparity = __fsub_rn(uL, bestuR);
                    if (disparity >= minD && disparity < maxD) {                           // :953
                        if (disparity <= 0) {
                            disparity = 0.01f;                                             // float(0.01)
                            bestuR = (float)((double)uL - 0.01);
                        }
                        resD = __fdiv_rn(mbf, disparity);
                        resU = bestuR;
                        resSad = bestSad;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        uRightOut[iL] = resU;
        depthOut[iL] = resD;
        bestROut[iL] = resR;
        sadOut[iL] = resSad;
    }
}

// median-based outlier cut (:967-980).  The sorted (SAD, iL) vector is only used for the value at index size/2 and
// for "everything >= thDist from the back", i.e. an order statistic + a filter: two-pass radix select on the
// 15-bit SAD.  One CTA per frame.
__global__ void __launch_bounds__(256) k_stereo_cut(const Plan* __restrict__ P, Bufs BL) {
    const int frame = blockIdx.x;
    const int tid = threadIdx.x;
    const size_t fo = (size_t)frame * P->kpCap;
    const int nL = BL.outCount[frame * 2];
    int* sad = BL.sad + fo;
    __shared__ int hist[256];
    __shared__ int sBin, sRank, sCount, sMedian;
    hist[tid] = 0;
    if (tid == 0) sCount = 0;
    __syncthreads();
    int local = 0;
    for (int i = tid; i < nL; i += 256) {
        const int s = sad[i];
        if (s >= 0) { atomicAdd(&hist[min(s >> 7, 255)], 1); local++; }
    }
    atomicAdd(&sCount, local);
    __syncthreads();
    const int count = sCount;
    if (count == 0) return;                      // the reference reads vDistIdx[0] of an empty vector here (UB)
    const int kth = count / 2;
    if (tid == 0) {
        int acc = 0, b = 0;
        for (; b < 256; b++) { if (acc + hist[b] > kth) break; acc += hist[b]; }
        sBin = b; sRank = kth - acc;
    }
    __syncthreads();
    const int bin = sBin;
    __syncthreads();
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < nL; i += 256) {
        const int s = sad[i];
        if (s >= 0 && min(s >> 7, 255) == bin) atomicAdd(&hist[s & 127], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0, b = 0;
        for (; b < 128; b++) { if (acc + hist[b] > sRank) break; acc += hist[b]; }
        sMedian = (bin << 7) | b;
    }
    __syncthreads();
    const float median = (float)sMedian;
    const float thDist = __fmul_rn(1.5f * 1.4f, median);                                   // :969
    for (int i = tid; i < nL; i += 256) {
        const int s = sad[i];
        if (s >= 0 && !((float)s < thDist)) {                                              // :973-978
            BL.uRight[fo + i] = -1.f;
            BL.depth[fo + i] = -1.f;
        }
    }
}

static int ensure_scratch(orbb_matcher* m, int slot, size_t bytes) {
    if (m->scratchCap[slot] >= bytes) return ORBB_OK;
    if (m->scratch[slot]) cudaFree(m->scratch[slot]);
    m->scratch[slot] = nullptr; m->scratchCap[slot] = 0;
    ORBM_CUDA(m, cudaMalloc(&m->scratch[slot], bytes));
    m->scratchCap[slot] = bytes;
    return ORBB_OK;
}

// ---- what orbb_nccl.cu (the sharded 2-NN) needs from this translation unit ----
int matcher_device(const orbb_matcher* m) { return m->device; }
cudaStream_t matcher_stream(const orbb_matcher* m) { return m->stream; }
int matcher_scratch(orbb_matcher* m, int slot, size_t bytes, void** out) {
    const int rc = ensure_scratch(m, slot, bytes);
    *out = rc ? nullptr : m->scratch[slot];
    return rc;
}
int matcher_error(orbb_matcher* m, int code, const char* msg) { return m_err(m, code, "%s", msg); }
int launch_merge_shards(orbb_matcher* m, const int32_t* idx_sh, const int32_t* dist_sh, size_t shard_stride, int nshards, int nq, int32_t* idx2,
                        int32_t* dist2) {
    k_knn2_merge_shards<<<(nq + 255) / 256, 256, 0, m->stream>>>(idx_sh, dist_sh, shard_stride, nshards, nq, idx2, dist2);
    m->launches++;
    ORBM_CUDA(m, cudaGetLastError());
    return ORBB_OK;
}

}  // namespace orbb

using namespace orbb;

extern "C" {

int orbb_hamming_distance(const uint8_t* a, const uint8_t* b) {
    // ORBmatcher::DescriptorDistance (ORBmatcher.cc:2058-2074): a single 32-byte pair stays on the host
    unsigned long long x[4], y[4];
    memcpy(x, a, 32);
    memcpy(y, b, 32);
    return __builtin_popcountll(x[0] ^ y[0]) + __builtin_popcountll(x[1] ^ y[1]) + __builtin_popcountll(x[2] ^ y[2]) +
           __builtin_popcountll(x[3] ^ y[3]);
}

int orbb_matcher_create(int device, orbb_matcher** out) {
    if (!out) return m_err(nullptr, ORBB_ERR_ARG, "null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return m_err(nullptr, ORBB_ERR_CUDA, "no CUDA device (%s): liborbb200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return m_err(nullptr, ORBB_ERR_ARG, "device %d out of range", device);
    orbb_matcher* m = new orbb_matcher();
    m->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete m;
        return m_err(nullptr, ORBB_ERR_CUDA, "cannot create stream on device %d", device);
    }
    *out = m;
    return ORBB_OK;
}

void orbb_matcher_destroy(orbb_matcher* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaStreamSynchronize(m->stream);
    if (m->partial) cudaFree(m->partial);
    for (int i = 0; i < 4; i++) if (m->scratch[i]) cudaFree(m->scratch[i]);
    for (int i = 0; i < ORBB_FRAME_SLOTS; i++) if (m->frame[i]) cudaFree(m->frame[i]);
    cudaStreamDestroy(m->stream);
    delete m;
}

const char* orbb_matcher_last_error(const orbb_matcher* m) { return m ? m->err.c_str() : g_lastError.c_str(); }
long long orbb_matcher_launch_count(const orbb_matcher* m) { return m ? m->launches : 0; }
void* orbb_matcher_stream(orbb_matcher* m) { return m ? (void*)m->stream : nullptr; }

int orbb_knn2_dev(orbb_matcher* m, const uint8_t* q_dev, int nq, const uint8_t* db_dev, int64_t nd, int32_t index_base,
                  int32_t* idx2_dev, int32_t* dist2_dev) {
    if (!m || !idx2_dev || !dist2_dev || nq < 0 || nd < 0) return m_err(m, ORBB_ERR_ARG, "bad argument");
    if (nq == 0) return ORBB_OK;
    if (((uintptr_t)q_dev | (uintptr_t)db_dev) & 15) return m_err(m, ORBB_ERR_ARG, "descriptor arrays must be 16-byte aligned");
    if (nd + (int64_t)index_base > INT_MAX) return m_err(m, ORBB_ERR_ARG, "database too large for 32-bit indices");
    ORBM_CUDA(m, cudaSetDevice(m->device));
    // (CSA, direct) queries per thread; ORBB_KNN_MIX=c,s overrides for experiments
    // default: 3 queries per thread, all through the 3-3-2 compressor (measured best: 820 Gpairs/s vs 526 direct)
    static int mixC = -1, mixS = -1, mixM = 3;
    if (mixC < 0) {
        mixC = 3; mixS = 0;
        if (const char* e = getenv("ORBB_KNN_MIX")) sscanf(e, "%d,%d,%d", &mixC, &mixS, &mixM);
    }
    const int qpt = mixC + mixS;
    const int qtile = KNN_THREADS * qpt;
    const int qtiles = (nq + qtile - 1) / qtile;
    // database chunks: enough CTAs to fill 148 SMs several times over, chunk a multiple of the tile, < 2^22 rows
    int nchunks = 1;
    if (nd > 0) {
        const int want = std::max(1, (148 * 8 + qtiles - 1) / qtiles);
        long long rowsPer = (nd + want - 1) / want;
        rowsPer = std::max<long long>(KNN_ROWS * 4, (rowsPer + KNN_ROWS - 1) / KNN_ROWS * KNN_ROWS);
        rowsPer = std::min<long long>(rowsPer, (1ll << KNN_IDX_BITS) - KNN_ROWS);
        nchunks = (int)((nd + rowsPer - 1) / rowsPer);
        const size_t need = (size_t)nchunks * nq * 2 * sizeof(unsigned);
        if (m->partialCap < need) {
            if (m->partial) cudaFree(m->partial);
            m->partial = nullptr; m->partialCap = 0;
            ORBM_CUDA(m, cudaMalloc((void**)&m->partial, need));
            m->partialCap = need;
        }
#define ORBB_KNN_LAUNCH(C, S)                                                                                                  \
    k_knn2_partial<C, S><<<dim3(qtiles, nchunks), KNN_THREADS, 0, m->stream>>>((const uint4*)q_dev, nq, (const uint4*)db_dev, nd, \
                                                                              (int)rowsPer, m->partial)
#define ORBB_KNN_LAUNCH_M(C, S, M)                                                                                                \
    k_knn2_partial<C, S, M, 0><<<dim3(qtiles, nchunks), KNN_THREADS, 0, m->stream>>>((const uint4*)q_dev, nq, (const uint4*)db_dev, \
                                                                                    nd, (int)rowsPer, m->partial)
        if (mixM == 3 && mixC == 3 && mixS == 0) ORBB_KNN_LAUNCH_M(3, 0, 3);          // production
        else if (mixM == 3 && mixC == 4 && mixS == 0) ORBB_KNN_LAUNCH_M(4, 0, 3);     // A/B variants, each covered by tests/test_gpu_extract.py::test_every_shipped_switch_keeps_parity
        else if (mixC == 0 && mixS == 4) ORBB_KNN_LAUNCH(0, 4);                        // no carry-save compression: 8 POPC per pair
        else if (mixC == 2 && mixS == 2) ORBB_KNN_LAUNCH(2, 2);
        else return m_err(m, ORBB_ERR_ARG, "unsupported ORBB_KNN_MIX=%d,%d", mixC, mixS);
#undef ORBB_KNN_LAUNCH
#undef ORBB_KNN_LAUNCH_M
        m->launches++;
        k_knn2_merge_chunks<<<(nq + 255) / 256, 256, 0, m->stream>>>(m->partial, nchunks, nq, (int)rowsPer, index_base, idx2_dev, dist2_dev);
        m->launches++;
    } else {
        k_knn2_merge_chunks<<<(nq + 255) / 256, 256, 0, m->stream>>>(nullptr, 0, nq, 0, index_base, idx2_dev, dist2_dev);
        m->launches++;
    }
    ORBM_CUDA(m, cudaGetLastError());
    return ORBB_OK;
}

int orbb_knn2(orbb_matcher* m, const uint8_t* q, int nq, const uint8_t* db, int64_t nd, int32_t* idx2, int32_t* dist2) {
    if (!m || !idx2 || !dist2 || nq < 0 || nd < 0) return m_err(m, ORBB_ERR_ARG, "bad argument");
    if (nq == 0) return ORBB_OK;
    ORBM_CUDA(m, cudaSetDevice(m->device));
    int rc;
    if ((rc = ensure_scratch(m, 0, (size_t)nq * 32)) || (rc = ensure_scratch(m, 1, std::max<size_t>((size_t)nd * 32, 32))) ||
        (rc = ensure_scratch(m, 2, (size_t)nq * 2 * sizeof(int))) || (rc = ensure_scratch(m, 3, (size_t)nq * 2 * sizeof(int))))
        return rc;
    ORBM_CUDA(m, cudaMemcpyAsync(m->scratch[0], q, (size_t)nq * 32, cudaMemcpyHostToDevice, m->stream));
    if (nd > 0) ORBM_CUDA(m, cudaMemcpyAsync(m->scratch[1], db, (size_t)nd * 32, cudaMemcpyHostToDevice, m->stream));
    rc = orbb_knn2_dev(m, (const uint8_t*)m->scratch[0], nq, (const uint8_t*)m->scratch[1], nd, 0, (int32_t*)m->scratch[2], (int32_t*)m->scratch[3]);
    if (rc) return rc;
    ORBM_CUDA(m, cudaMemcpyAsync(idx2, m->scratch[2], (size_t)nq * 2 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    ORBM_CUDA(m, cudaMemcpyAsync(dist2, m->scratch[3], (size_t)nq * 2 * sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
    return ORBB_OK;
}

int orbb_knn2_merge_dev(orbb_matcher* m, const int32_t* idx_sh_dev, const int32_t* dist_sh_dev, int nshards, int nq,
                        int32_t* idx2_dev, int32_t* dist2_dev) {
    if (!m || !idx_sh_dev || !dist_sh_dev || !idx2_dev || !dist2_dev || nshards < 1 || nq < 0) return m_err(m, ORBB_ERR_ARG, "bad argument");
    if (nq == 0) return ORBB_OK;
    ORBM_CUDA(m, cudaSetDevice(m->device));
    return launch_merge_shards(m, idx_sh_dev, dist_sh_dev, (size_t)nq * 2, nshards, nq, idx2_dev, dist2_dev);
}

int orbb_ratio_test_dev(orbb_matcher* m, const int32_t* idx2_dev, const int32_t* dist2_dev, int nq, double ratio, uint8_t* keep_dev) {
    if (!m || !idx2_dev || !dist2_dev || !keep_dev || nq < 0) return m_err(m, ORBB_ERR_ARG, "bad argument");
    if (nq == 0) return ORBB_OK;
    ORBM_CUDA(m, cudaSetDevice(m->device));
    k_ratio_test<<<(nq + 255) / 256, 256, 0, m->stream>>>(idx2_dev, dist2_dev, nq, ratio, keep_dev);
    m->launches++;
    ORBM_CUDA(m, cudaGetLastError());
    return ORBB_OK;
}

static int best2_csr_impl(orbb_matcher* m, const uint8_t* q, int nq, const uint8_t* train, int ntrain, bool trainOnDevice, const int32_t* cand,
                          const int32_t* rowptr, int init, int32_t* out4) {
    if (!m || !rowptr || !out4 || nq < 0 || ntrain < 0 || (nq > 0 && !q)) return m_err(m, ORBB_ERR_ARG, "bad argument");
    if (nq == 0) return ORBB_OK;
    // the arrays are host-resident: a bad candidate index would become an out-of-bounds device read that poisons the CUDA context
    if (rowptr[0] != 0) return m_err(m, ORBB_ERR_ARG, "rowptr[0] must be 0");
    for (int i = 0; i < nq; i++)
        if (rowptr[i + 1] < rowptr[i]) return m_err(m, ORBB_ERR_ARG, "rowptr decreases at query %d", i);
    const int ncand = rowptr[nq];
    if (ncand > 0 && (!cand || !train)) return m_err(m, ORBB_ERR_ARG, "candidate lists without cand / train arrays");
    for (int i = 0; i < ncand; i++)
        if (cand[i] < 0 || cand[i] >= ntrain) return m_err(m, ORBB_ERR_ARG, "candidate %d = %d outside [0, %d)", i, cand[i], ntrain);
    ORBM_CUDA(m, cudaSetDevice(m->device));
    const size_t bq = (size_t)nq * 32, bt = std::max<size_t>((size_t)ntrain * 32, 32), bc = std::max<size_t>((size_t)ncand * 4, 4);
    const size_t br = (size_t)(nq + 1) * 4, bo = (size_t)nq * 16;
    int rc;
    if ((rc = ensure_scratch(m, 0, bq + 256)) || (rc = ensure_scratch(m, 1, bt)) || (rc = ensure_scratch(m, 2, bc + br + 512)) ||
        (rc = ensure_scratch(m, 3, bo)))
        return rc;
    int* dCand = (int*)m->scratch[2];
    int* dRow = (int*)((char*)m->scratch[2] + (bc + 255) / 256 * 256);
    ORBM_CUDA(m, cudaMemcpyAsync(m->scratch[0], q, bq, cudaMemcpyHostToDevice, m->stream));
    if (ntrain > 0 && !trainOnDevice) ORBM_CUDA(m, cudaMemcpyAsync(m->scratch[1], train, (size_t)ntrain * 32, cudaMemcpyHostToDevice, m->stream));
    if (ncand > 0) ORBM_CUDA(m, cudaMemcpyAsync(dCand, cand, (size_t)ncand * 4, cudaMemcpyHostToDevice, m->stream));
    ORBM_CUDA(m, cudaMemcpyAsync(dRow, rowptr, br, cudaMemcpyHostToDevice, m->stream));
    k_best2_csr<<<(nq + 7) / 8, 256, 0, m->stream>>>((const uint4*)m->scratch[0], nq, trainOnDevice ? (const uint4*)train : (const uint4*)m->scratch[1], dCand, dRow,
                                                    init, (int*)m->scratch[3]);
    m->launches++;
    ORBM_CUDA(m, cudaGetLastError());
    ORBM_CUDA(m, cudaMemcpyAsync(out4, m->scratch[3], bo, cudaMemcpyDeviceToHost, m->stream));
    ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
    return ORBB_OK;
}

// frame-side arrays resident on the device for the matcher scans (orbb_frame_upload): one slot = one frame
int orbb_best2_csr(orbb_matcher* m, const uint8_t* q, int nq, const uint8_t* train, int ntrain, const int32_t* cand, const int32_t* rowptr,
                   int init, int32_t* out4) {
    return best2_csr_impl(m, q, nq, train, ntrain, false, cand, rowptr, init, out4);
}
int orbb_best2_csr_dev(orbb_matcher* m, const uint8_t* q, int nq, const uint8_t* train_dev, int ntrain, const int32_t* cand, const int32_t* rowptr,
                       int init, int32_t* out4) {
    return best2_csr_impl(m, q, nq, train_dev, ntrain, true, cand, rowptr, init, out4);
}

int orbb_search_area_topk(orbb_matcher* m, const orbb_frame_view* f, const float* grid4, const float* queries, const int32_t* qlev,
                          const uint8_t* qdesc, int nq, const uint8_t* skip, int init, int k, int32_t* out) {
    if (!m || !f || !grid4 || !out || f->n < 0 || nq < 0 || (k != 1 && k != 2 && k != 4 && k != 8) ||
        (f->n > 0 && (!f->kps_xy || !f->octaves || !f->desc || f->kps_stride < 8 || f->oct_stride < 4)) || (nq > 0 && (!queries || !qlev || !qdesc)))
        return m_err(m, ORBB_ERR_ARG, "bad argument");
    if (nq == 0) return ORBB_OK;
    const int n = f->n;
    if (n >= (1 << 20)) return m_err(m, ORBB_ERR_UNSUPPORTED, "more than 2^20 keypoints per frame");
    ORBM_CUDA(m, cudaSetDevice(m->device));
    // one staging buffer: [queries nq*16][qlev nq*8][qdesc nq*32][skip n][out nq*k*8] and, for a host-resident frame,
    // [kps n*stride][oct n*stride][train n*32][uRight n*4]; 256-B aligned parts
    const bool dev = f->on_device != 0;
    const size_t kb = dev ? 0 : (size_t)n * f->kps_stride, ob = dev ? 0 : (size_t)n * f->oct_stride;
    const bool sameArray = !dev && (const char*)f->octaves >= (const char*)f->kps_xy && (const char*)f->octaves < (const char*)f->kps_xy + f->kps_stride &&
                           f->oct_stride == f->kps_stride;      // octaves live inside the key point records: upload them once
    size_t off[10], cur = 0;
    const size_t sz[9] = {(size_t)nq * 16, (size_t)nq * 8, (size_t)nq * 32, skip ? (size_t)n : 0, (size_t)nq * k * 8, kb, sameArray ? 0 : ob,
                          dev ? 0 : (size_t)n * 32, (dev || !f->u_right) ? 0 : (size_t)n * 4};
    for (int i = 0; i < 9; i++) { off[i] = cur; cur += (sz[i] + 255) / 256 * 256; }
    int rc;
    if ((rc = ensure_scratch(m, 0, cur + 256))) return rc;
    char* base = (char*)m->scratch[0];
    const void* src[9] = {queries, qlev, qdesc, skip, nullptr, f->kps_xy, f->octaves, f->desc, f->u_right};
    for (int i = 0; i < 9; i++)
        if (sz[i] && src[i]) ORBM_CUDA(m, cudaMemcpyAsync(base + off[i], src[i], sz[i], cudaMemcpyHostToDevice, m->stream));
    const uint8_t* dK = dev ? (const uint8_t*)f->kps_xy : (const uint8_t*)(base + off[5]);
    const uint8_t* dO = dev ? (const uint8_t*)f->octaves : sameArray ? dK + ((const char*)f->octaves - (const char*)f->kps_xy) : (const uint8_t*)(base + off[6]);
    const uint4* dT = dev ? (const uint4*)f->desc : (const uint4*)(base + off[7]);
    const float* dU = !f->u_right ? nullptr : dev ? f->u_right : (const float*)(base + off[8]);
#define ORBB_TOPK(KK)                                                                                                                       \
    k_search_area_topk<KK><<<(nq + 7) / 8, 256, 0, m->stream>>>(dK, f->kps_stride, dO, f->oct_stride, dT, n, grid4[0], grid4[1], grid4[2], grid4[3], \
                                                               (const float4*)(base + off[0]), (const int2*)(base + off[1]),                   \
                                                               (const uint4*)(base + off[2]), nq, skip ? (const uint8_t*)(base + off[3]) : nullptr, \
                                                               dU, init, (int*)(base + off[4]))
    if (k == 1) ORBB_TOPK(1);
    else if (k == 2) ORBB_TOPK(2);
    else if (k == 4) ORBB_TOPK(4);
    else ORBB_TOPK(8);
#undef ORBB_TOPK
    m->launches++;
    ORBM_CUDA(m, cudaGetLastError());
    ORBM_CUDA(m, cudaMemcpyAsync(out, base + off[4], sz[4], cudaMemcpyDeviceToHost, m->stream));
    ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
    return ORBB_OK;
}

int orbb_search_area_best2(orbb_matcher* m, const float* kps_xy, const int32_t* octaves, const uint8_t* train, int n, const float* grid4,
                           const float* queries, const int32_t* qlev, const uint8_t* qdesc, int nq, const uint8_t* skip, const float* u_right,
                           int init, int32_t* out4) {
    orbb_frame_view f;
    f.kps_xy = kps_xy; f.kps_stride = 8; f.octaves = octaves; f.oct_stride = 4; f.desc = train; f.u_right = u_right; f.n = n; f.on_device = 0;
    return orbb_search_area_topk(m, &f, grid4, queries, qlev, qdesc, nq, skip, init, 2, out4);
}

int orbb_frame_upload(orbb_matcher* m, int slot, const orbb_frame_view* host, orbb_frame_view* dev) {
    if (!m || !host || !dev || slot < 0 || slot >= ORBB_FRAME_SLOTS || host->n < 0 || host->on_device ||
        (host->n > 0 && (!host->kps_xy || !host->octaves || !host->desc || host->kps_stride < 8 || host->oct_stride < 4)))
        return m_err(m, ORBB_ERR_ARG, "bad argument");
    ORBM_CUDA(m, cudaSetDevice(m->device));
    const int n = host->n;
    // packed on the device as {x, y} floats, int32 octaves, 32-byte descriptors, float u_right
    const size_t bk = ((size_t)n * 8 + 255) / 256 * 256, bo = ((size_t)n * 4 + 255) / 256 * 256, bd = ((size_t)n * 32 + 255) / 256 * 256;
    const size_t need = bk + bo + bd + bo + 256;
    if (m->frameCap[slot] < need) {
        if (m->frame[slot]) cudaFree(m->frame[slot]);
        m->frame[slot] = nullptr; m->frameCap[slot] = 0;
        ORBM_CUDA(m, cudaMalloc(&m->frame[slot], need));
        m->frameCap[slot] = need;
    }
    char* base = (char*)m->frame[slot];
    if (n > 0) {
        ORBM_CUDA(m, cudaMemcpy2DAsync(base, 8, host->kps_xy, host->kps_stride, 8, n, cudaMemcpyHostToDevice, m->stream));
        ORBM_CUDA(m, cudaMemcpy2DAsync(base + bk, 4, host->octaves, host->oct_stride, 4, n, cudaMemcpyHostToDevice, m->stream));
        ORBM_CUDA(m, cudaMemcpyAsync(base + bk + bo, host->desc, (size_t)n * 32, cudaMemcpyHostToDevice, m->stream));
        if (host->u_right) ORBM_CUDA(m, cudaMemcpyAsync(base + bk + bo + bd, host->u_right, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
    }
    dev->kps_xy = base; dev->kps_stride = 8; dev->octaves = base + bk; dev->oct_stride = 4; dev->desc = (const uint8_t*)(base + bk + bo);
    dev->u_right = host->u_right ? (const float*)(base + bk + bo + bd) : nullptr;
    dev->n = n; dev->on_device = 1;
    return ORBB_OK;                                        // (the copies are ordered before any later scan on the matcher's stream)
}

int orbb_distinctive_csr(orbb_matcher* m, const uint8_t* desc, int ntotal, const int32_t* rowptr, int ngroups, int32_t* best) {
    if (!m || !rowptr || !best || ngroups < 0 || ntotal < 0) return m_err(m, ORBB_ERR_ARG, "bad argument");
    if (ngroups == 0) return ORBB_OK;
    if (rowptr[ngroups] > ntotal) return m_err(m, ORBB_ERR_ARG, "rowptr exceeds the descriptor array");
    for (int g = 0; g < ngroups; g++)
        if (rowptr[g + 1] - rowptr[g] > 65535) return m_err(m, ORBB_ERR_UNSUPPORTED, "group %d has more than 65535 descriptors", g);
    ORBM_CUDA(m, cudaSetDevice(m->device));
    const size_t bd = std::max<size_t>((size_t)ntotal * 32, 32), br = (size_t)(ngroups + 1) * 4, bo = (size_t)ngroups * 4;
    int rc;
    if ((rc = ensure_scratch(m, 0, bd)) || (rc = ensure_scratch(m, 2, br)) || (rc = ensure_scratch(m, 3, bo))) return rc;
    if (ntotal > 0) ORBM_CUDA(m, cudaMemcpyAsync(m->scratch[0], desc, (size_t)ntotal * 32, cudaMemcpyHostToDevice, m->stream));
    ORBM_CUDA(m, cudaMemcpyAsync(m->scratch[2], rowptr, br, cudaMemcpyHostToDevice, m->stream));
    k_distinctive<<<(ngroups + 3) / 4, 128, 0, m->stream>>>((const uint4*)m->scratch[0], (const int*)m->scratch[2], ngroups, (int*)m->scratch[3]);
    m->launches++;
    ORBM_CUDA(m, cudaGetLastError());
    ORBM_CUDA(m, cudaMemcpyAsync(best, m->scratch[3], bo, cudaMemcpyDeviceToHost, m->stream));
    ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
    return ORBB_OK;
}

// ---- rotation-consistency filter of the match scans (ORBmatcher.cc:345-352, :405-423, ComputeThreeMaxima :2012-2053) -------
int orbb_rotation_check_csr(orbb_matcher* m, const float* angle_a, const float* angle_b, int total, const int32_t* rowptr, int nsets,
                            uint8_t* keep, int32_t* ind3) {
    if (!m || total < 0 || nsets < 0 || !rowptr || (total > 0 && (!angle_a || !angle_b || !keep)) || (nsets > 0 && !ind3))
        return m_err(m, ORBB_ERR_ARG, "bad argument");
    if (nsets == 0) return ORBB_OK;
    for (int s = 0; s < nsets; s++)
        if (rowptr[s + 1] < rowptr[s] || rowptr[s + 1] > total) return m_err(m, ORBB_ERR_ARG, "bad rowptr at set %d", s);
    ORBM_CUDA(m, cudaSetDevice(m->device));
    const size_t ba = std::max<size_t>(sizeof(float) * (size_t)total, 4), br = sizeof(int) * (size_t)(nsets + 1);
    int rc;
    if ((rc = ensure_scratch(m, 0, 2 * ba)) || (rc = ensure_scratch(m, 1, br)) || (rc = ensure_scratch(m, 2, std::max<size_t>(total, 1))) ||
        (rc = ensure_scratch(m, 3, sizeof(int) * 3 * (size_t)nsets)))
        return rc;
    float* dA = (float*)m->scratch[0];
    float* dB = dA + std::max(total, 1);
    if (total > 0) {
        ORBM_CUDA(m, cudaMemcpyAsync(dA, angle_a, sizeof(float) * total, cudaMemcpyHostToDevice, m->stream));
        ORBM_CUDA(m, cudaMemcpyAsync(dB, angle_b, sizeof(float) * total, cudaMemcpyHostToDevice, m->stream));
    }
    ORBM_CUDA(m, cudaMemcpyAsync(m->scratch[1], rowptr, br, cudaMemcpyHostToDevice, m->stream));
    k_rotation_check<<<nsets, 256, 0, m->stream>>>(dA, dB, (const int*)m->scratch[1], (uint8_t*)m->scratch[2], (int*)m->scratch[3]);
    m->launches++;
    ORBM_CUDA(m, cudaGetLastError());
    if (total > 0) ORBM_CUDA(m, cudaMemcpyAsync(keep, m->scratch[2], total, cudaMemcpyDeviceToHost, m->stream));
    ORBM_CUDA(m, cudaMemcpyAsync(ind3, m->scratch[3], sizeof(int) * 3 * (size_t)nsets, cudaMemcpyDeviceToHost, m->stream));
    ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
    return ORBB_OK;
}

// ---- Frame::UndistortKeyPoints ("next" row) -----------------------------------------------------------------------
int orbb_undistort_points(orbb_matcher* m, const float* xy, int n, const float* K4, const float* dist, int ndist, const float* newK4,
                          float* out_xy) {
    if (!m || n < 0 || !K4 || !newK4 || (n > 0 && (!xy || !out_xy)) || ndist < 0 || ndist > 12 || (ndist > 0 && !dist))
        return m_err(m, ORBB_ERR_ARG, "bad argument");
    if (n == 0) return ORBB_OK;
    if (ndist == 0 || dist[0] == 0.0f) {                   // Frame.cc:749-753: mDistCoef.at<float>(0) == 0 -> mvKeysUn = mvKeys
        memcpy(out_xy, xy, sizeof(float) * 2 * (size_t)n);
        return ORBB_OK;
    }
    ORBM_CUDA(m, cudaSetDevice(m->device));
    UndistortParams p;
    p.fx = K4[0]; p.fy = K4[1]; p.cx = K4[2]; p.cy = K4[3];
    p.nfx = newK4[0]; p.nfy = newK4[1]; p.ncx = newK4[2]; p.ncy = newK4[3];
    for (int i = 0; i < 12; i++) p.k[i] = i < ndist ? (double)dist[i] : 0.0;
    int rc;
    const size_t bytes = sizeof(float) * 2 * (size_t)n;
    if ((rc = ensure_scratch(m, 0, bytes)) || (rc = ensure_scratch(m, 1, bytes))) return rc;
    ORBM_CUDA(m, cudaMemcpyAsync(m->scratch[0], xy, bytes, cudaMemcpyHostToDevice, m->stream));
    k_undistort<<<(n + 255) / 256, 256, 0, m->stream>>>((const float2*)m->scratch[0], (float2*)m->scratch[1], n, p);
    m->launches++;
    ORBM_CUDA(m, cudaGetLastError());
    ORBM_CUDA(m, cudaMemcpyAsync(out_xy, m->scratch[1], bytes, cudaMemcpyDeviceToHost, m->stream));
    ORBM_CUDA(m, cudaStreamSynchronize(m->stream));
    return ORBB_OK;
}

// ---- stereo ----------------------------------------------------------------------------------------
int orbb_stereo_match_batch(orbb_extractor* hL, orbb_extractor* hR, int nframes, float bf, float b) {
    if (!hL || !hR) return ORBB_ERR_ARG;
    if (!hL->planValid || !hR->planValid || nframes < 1 || nframes > hL->lastFrames || nframes > hR->lastFrames)
        return set_err(hL, ORBB_ERR_ARG, "stereo: both extractors must hold a batch of >= %d frames", nframes);
    const Plan &A = hL->plan, &B = hR->plan;
    if (hL->device != hR->device || A.W != B.W || A.H != B.H || A.nlevels != B.nlevels || A.kpCap != B.kpCap || A.pyrStride != B.pyrStride)
        return set_err(hL, ORBB_ERR_ARG, "stereo: left/right extractors differ in geometry or parameters");
    if (A.kpCap >= 65535) return set_err(hL, ORBB_ERR_UNSUPPORTED, "stereo: more than 65534 keypoints per image");
    ORBB_CUDA(hL, cudaSetDevice(hL->device));
    // the left stream waits for the right extractor's work
    cudaEvent_t ev;
    ORBB_CUDA(hL, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ORBB_CUDA(hL, cudaEventRecord(ev, hR->stream));
    ORBB_CUDA(hL, cudaStreamWaitEvent(hL->stream, ev, 0));
    ORBB_CUDA(hL, cudaEventDestroy(ev));
    // right keypoints bucketed by row (dynamic shared memory: two ints per image row); taller images scan all candidates
    const size_t rowSmem = (size_t)(A.H + 1) * 2 * sizeof(int);
    const int useRows = rowSmem <= 40 * 1024 && A.H <= 32767;
    if (useRows) {
        k_stereo_rows<<<nframes, 256, rowSmem, hL->stream>>>(hR->dPlan, hR->b);
        hL->launches++;
    }
    k_stereo_match<<<dim3((A.kpCap + 7) / 8, nframes), 256, 0, hL->stream>>>(hL->dPlan, hL->b, hR->b, bf, b, useRows);
    k_stereo_cut<<<nframes, 256, 0, hL->stream>>>(hL->dPlan, hL->b);
    hL->launches += 2;
    ORBB_CUDA(hL, cudaGetLastError());
    // ... and the right extractor's next batch waits for the stereo kernels, which read ITS pyramid, key points, descriptors and
    // row buckets on the left stream (orbb_extract_batch is asynchronous: without this the next right-image extraction could
    // overwrite them under the running stereo kernels)
    ORBB_CUDA(hL, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ORBB_CUDA(hL, cudaEventRecord(ev, hL->stream));
    ORBB_CUDA(hL, cudaStreamWaitEvent(hR->stream, ev, 0));
    ORBB_CUDA(hL, cudaEventDestroy(ev));
    return ORBB_OK;
}

int orbb_rgbd_stereo_batch(orbb_extractor* h, const void* dev_depth, int depth_is_u16, float depth_factor, size_t row_stride,
                           size_t frame_stride, int nframes, const float* K4, const float* dist, int ndist, float bf) {
    if (!h) return ORBB_ERR_ARG;
    if (!h->planValid || nframes < 1 || nframes > h->lastFrames) return set_err(h, ORBB_ERR_ARG, "rgbd: the extractor must hold a batch of >= %d frames", nframes);
    if (!dev_depth || !K4 || ndist < 0 || ndist > 12 || (ndist > 0 && !dist)) return set_err(h, ORBB_ERR_ARG, "rgbd: bad argument");
    const Plan& P = h->plan;
    const size_t px = depth_is_u16 ? 2 : 4;
    if (row_stride < (size_t)P.W * px || frame_stride < row_stride * (size_t)P.H) return set_err(h, ORBB_ERR_ARG, "rgbd: depth strides smaller than the image");
    ORBB_CUDA(h, cudaSetDevice(h->device));
    UndistortParams p;
    p.fx = K4[0]; p.fy = K4[1]; p.cx = K4[2]; p.cy = K4[3];
    p.nfx = K4[0]; p.nfy = K4[1]; p.ncx = K4[2]; p.ncy = K4[3];          // Frame.cc:766: P = mK
    for (int i = 0; i < 12; i++) p.k[i] = i < ndist ? (double)dist[i] : 0.0;
    const int undist = ndist > 0 && dist[0] != 0.0f;                      // Frame.cc:749
    k_rgbd_stereo<<<dim3((P.kpCap + 255) / 256, nframes), 256, 0, h->stream>>>(h->dPlan, h->b, (const uint8_t*)dev_depth, depth_is_u16, depth_factor,
                                                                         row_stride, frame_stride, p, undist, bf);
    h->launches++;
    ORBB_CUDA(h, cudaGetLastError());
    return ORBB_OK;
}

int orbb_stereo_fetch(orbb_extractor* hL, int nframes, float* u_right, float* depth, int capacity) {
    if (!hL || !hL->planValid || nframes < 1 || nframes > hL->lastFrames) return ORBB_ERR_ARG;
    ORBB_CUDA(hL, cudaSetDevice(hL->device));
    const Plan& P = hL->plan;
    const int ncopy = std::min(capacity, P.kpCap);
    if (u_right) ORBB_CUDA(hL, cudaMemcpy2DAsync(u_right, sizeof(float) * capacity, hL->b.uRight, sizeof(float) * P.kpCap, sizeof(float) * ncopy, nframes, cudaMemcpyDeviceToHost, hL->stream));
    if (depth) ORBB_CUDA(hL, cudaMemcpy2DAsync(depth, sizeof(float) * capacity, hL->b.depth, sizeof(float) * P.kpCap, sizeof(float) * ncopy, nframes, cudaMemcpyDeviceToHost, hL->stream));
    ORBB_CUDA(hL, cudaStreamSynchronize(hL->stream));
    return ORBB_OK;
}

int orbb_stereo_match(orbb_extractor* hL, orbb_extractor* hR, int frame, float bf, float b, float* u_right, float* depth,
                      int32_t* best_r, int32_t* sad, int capacity, int* n_left) {
    if (!hL || !hR || frame < 0) return ORBB_ERR_ARG;
    int rc = orbb_stereo_match_batch(hL, hR, frame + 1, bf, b);
    if (rc) return rc;
    const Plan& P = hL->plan;
    int nL = 0;
    ORBB_CUDA(hL, cudaMemcpyAsync(&nL, hL->b.outCount + 2 * frame, sizeof(int), cudaMemcpyDeviceToHost, hL->stream));
    ORBB_CUDA(hL, cudaStreamSynchronize(hL->stream));
    if (n_left) *n_left = nL;
    if (nL > capacity) return set_err(hL, ORBB_ERR_CAPACITY, "stereo: %d left keypoints, capacity %d", nL, capacity);
    const size_t fo = (size_t)frame * P.kpCap;
    if (u_right) ORBB_CUDA(hL, cudaMemcpyAsync(u_right, hL->b.uRight + fo, sizeof(float) * nL, cudaMemcpyDeviceToHost, hL->stream));
    if (depth) ORBB_CUDA(hL, cudaMemcpyAsync(depth, hL->b.depth + fo, sizeof(float) * nL, cudaMemcpyDeviceToHost, hL->stream));
    if (best_r) ORBB_CUDA(hL, cudaMemcpyAsync(best_r, hL->b.bestR + fo, sizeof(int) * nL, cudaMemcpyDeviceToHost, hL->stream));
    if (sad) ORBB_CUDA(hL, cudaMemcpyAsync(sad, hL->b.sad + fo, sizeof(int) * nL, cudaMemcpyDeviceToHost, hL->stream));
    ORBB_CUDA(hL, cudaStreamSynchronize(hL->stream));
    return ORBB_OK;
}

}  // extern "C"

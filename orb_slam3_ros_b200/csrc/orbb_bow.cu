// liborbb200.so -- bag-of-words transform ("next" row of the scope table).
//
// Frame::ComputeBoW / KeyFrame::ComputeBoW (reference orb_slam3/src/Frame.cc:738-745) call
// DBoW2::TemplatedVocabulary<FORB>::transform(features, BowVector&, FeatureVector&, levelsup = 4)
// (orb_slam3/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1213 and :1230-1275): every descriptor descends the
// k-ary vocabulary tree (at each node the child with the smallest Hamming distance, first child wins ties,
// FORB::distance = FORB.cpp:81-101), the leaf's word gets the leaf's tf-idf weight added (BowVector::addWeight,
// BowVector.cpp:30-42), the ancestor `levelsup` levels above the leaves collects the feature index
// (FeatureVector::addFeature, FeatureVector.cpp:31-45), and the vector is L1/L2-normalised (BowVector.cpp:58-80).
//
//   k_bow_descend    one warp per descriptor; lanes = children of the current node (children are stored contiguously in
//                    visiting order), arg-min by unsigned min over  dist << 16 | child position.
//   k_bow_assemble   one CTA per descriptor set: ranks by (word, feature) and (node, feature), then reproduces the
//                    std::map semantics -- per word the weights are added in feature order, the norm is accumulated in
//                    ascending word order -- sequentially where the order of double additions is observable.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "orbb_internal.cuh"

struct orbb_vocab {
    int device = 0;
    int nnodes = 0, depth = 0;
    int* childBegin = nullptr;
    int* childCount = nullptr;
    int* childId = nullptr;
    uint4* childDesc = nullptr;
    double* weight = nullptr;
    int* wordId = nullptr;
    // scratch of the last transform
    void* scratch = nullptr;
    size_t scratchCap = 0;
    cudaStream_t stream = nullptr;
    long long launches = 0;
    std::string err;
};

namespace orbb {

static int v_err(orbb_vocab* v, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (v) v->err = buf;
    g_lastError = buf;
    return code;
}
#define ORBV_CUDA(v, call)                                                                                     \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess) return orbb::v_err(v, ORBB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

__device__ __forceinline__ int ham256(const uint4 a0, const uint4 a1, const uint4 b0, const uint4 b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
           __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)  (:1230-1275)
__global__ void __launch_bounds__(256) k_bow_descend(const uint4* __restrict__ desc, int n, const int* __restrict__ childBegin,
                                                    const int* __restrict__ childCount, const int* __restrict__ childId,
                                                    const uint4* __restrict__ childDesc, const double* __restrict__ weight,
                                                    const int* __restrict__ wordId, int nidLevel, int* __restrict__ outWord,
                                                    int* __restrict__ outNode, double* __restrict__ outWeight) {
    const int f = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= n) return;
    const uint4 a0 = __ldg(desc + (size_t)f * 2), a1 = __ldg(desc + (size_t)f * 2 + 1);
    int node = 0, level = 0, nid = 0;                        // nid_level <= 0 -> root (:1240)
    int cnt = childCount[0];
    while (cnt > 0) {                                        // do { ... } while(!isLeaf)
        ++level;
        const int beg = childBegin[node];
        unsigned best = 0xffffffffu;
        for (int c = lane; c < cnt; c += 32) {
            const int d = ham256(a0, a1, __ldg(childDesc + (size_t)(beg + c) * 2), __ldg(childDesc + (size_t)(beg + c) * 2 + 1));
            best = min(best, ((unsigned)d << 16) | (unsigned)c);     // strict '<' in visiting order == smallest (d, position)
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
        node = childId[beg + (int)(best & 0xffffu)];
        if (level == nidLevel) nid = node;
        cnt = childCount[node];
    }
    if (lane == 0) {
        outWord[f] = wordId[node];
        outNode[f] = nid;
        outWeight[f] = weight[node];
    }
}

// BowVector / FeatureVector of one descriptor set (features [beg, end) of the batch).
__global__ void __launch_bounds__(256) k_bow_assemble(const int* __restrict__ rowptr, const int* __restrict__ word, const int* __restrict__ node,
                                                     const double* __restrict__ weight, int norm, unsigned long long* __restrict__ keyW,
                                                     unsigned long long* __restrict__ keyN, int* __restrict__ bowId, double* __restrict__ bowVal,
                                                     int* __restrict__ fvNode, int* __restrict__ fvStart, int* __restrict__ fvFeat,
                                                     int* __restrict__ counts) {
    const int set = blockIdx.x, tid = threadIdx.x;
    const int beg = rowptr[set], n = rowptr[set + 1] - beg;
    __shared__ int sScan[256];
    __shared__ int sValid, sWords, sNodes;
    __shared__ double sNorm;
    // ---- 1. features with a positive weight ("not stopped", :1169), ranked by (word, feature) and by (node, feature) ----
    if (tid == 0) { sValid = 0; sWords = 0; sNodes = 0; }
    __syncthreads();
    int local = 0;
    for (int i = tid; i < n; i += 256) local += weight[beg + i] > 0;
    atomicAdd(&sValid, local);
    __syncthreads();
    const int nv = sValid;
    for (int i = tid; i < n; i += 256) {
        if (!(weight[beg + i] > 0)) continue;
        const unsigned long long kw = ((unsigned long long)(unsigned)word[beg + i] << 24) | (unsigned)i;
        const unsigned long long kn = ((unsigned long long)(unsigned)node[beg + i] << 24) | (unsigned)i;
        int rw = 0, rn = 0;
        for (int j = 0; j < n; j++) {
            if (!(weight[beg + j] > 0)) continue;
            rw += (((unsigned long long)(unsigned)word[beg + j] << 24) | (unsigned)j) < kw;
            rn += (((unsigned long long)(unsigned)node[beg + j] << 24) | (unsigned)j) < kn;
        }
        keyW[beg + rw] = kw;
        keyN[beg + rn] = kn;
    }
    __syncthreads();
    // ---- 2. BowVector: one entry per distinct word, weights added in feature order (addWeight, BowVector.cpp:30-42) ----
    int run = 0;
    for (int base = 0; base < nv; base += 256) {
        const int i = base + tid;
        const bool head = i < nv && (i == 0 || (keyW[beg + i] >> 24) != (keyW[beg + i - 1] >> 24));
        sScan[tid] = head;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {            // inclusive Hillis-Steele scan
            const int v = tid >= off ? sScan[tid - off] : 0;
            __syncthreads();
            sScan[tid] += v;
            __syncthreads();
        }
        if (head) {
            const int slot = run + sScan[tid] - 1;
            const unsigned long long w = keyW[beg + i] >> 24;
            double acc = 0.0;
            for (int j = i; j < nv && (keyW[beg + j] >> 24) == w; j++) acc += weight[beg + (int)(keyW[beg + j] & 0xffffffu)];
            bowId[beg + slot] = (int)w;
            bowVal[beg + slot] = acc;
        }
        run += sScan[255];
        __syncthreads();
    }
    if (tid == 0) sWords = run;
    // ---- 3. FeatureVector: one entry per distinct node, feature indices ascending (addFeature, FeatureVector.cpp:31-45) ----
    run = 0;
    for (int base = 0; base < nv; base += 256) {
        const int i = base + tid;
        const bool head = i < nv && (i == 0 || (keyN[beg + i] >> 24) != (keyN[beg + i - 1] >> 24));
        if (i < nv) fvFeat[beg + i] = (int)(keyN[beg + i] & 0xffffffu);
        sScan[tid] = head;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {
            const int v = tid >= off ? sScan[tid - off] : 0;
            __syncthreads();
            sScan[tid] += v;
            __syncthreads();
        }
        if (head) {
            const int slot = run + sScan[tid] - 1;
            fvNode[beg + slot] = (int)(keyN[beg + i] >> 24);
            fvStart[beg + slot] = i;
        }
        run += sScan[255];
        __syncthreads();
    }
    if (tid == 0) sNodes = run;
    __syncthreads();
    // ---- 4. normalisation in ascending word order (BowVector::normalize, BowVector.cpp:58-80) ----
    const int nw = sWords;
    if (tid == 0) {
        double acc = 0.0;
        if (norm == 1) for (int k = 0; k < nw; k++) acc += fabs(bowVal[beg + k]);
        else if (norm == 2) { for (int k = 0; k < nw; k++) acc += bowVal[beg + k] * bowVal[beg + k]; acc = sqrt(acc); }
        sNorm = acc;
    }
    __syncthreads();
    const double nrm = sNorm;
    if (norm != 0 && nrm > 0.0)
        for (int k = tid; k < nw; k += 256) bowVal[beg + k] = bowVal[beg + k] / nrm;
    if (norm == 0 && nw > 0)                                 // TemplatedVocabulary.h:1176-1182
        for (int k = tid; k < nw; k += 256) bowVal[beg + k] = bowVal[beg + k] / (double)nw;
    if (tid == 0) { counts[3 * set] = nw; counts[3 * set + 1] = sNodes; counts[3 * set + 2] = nv; }
}

}  // namespace orbb

using namespace orbb;

extern "C" {

int orbb_vocab_create(int device, int nnodes, const int32_t* child_begin, const int32_t* child_count, const int32_t* child_list,
                      int nchildren, const uint8_t* node_desc, const double* node_weight, const int32_t* node_word_id, int depth,
                      orbb_vocab** out) {
    if (!out || nnodes < 1 || !child_begin || !child_count || !node_desc || !node_weight || !node_word_id || depth < 1)
        return v_err(nullptr, ORBB_ERR_ARG, "bad vocabulary arguments");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return v_err(nullptr, ORBB_ERR_CUDA, "no CUDA device (%s): liborbb200 has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return v_err(nullptr, ORBB_ERR_ARG, "device %d out of range", device);
    for (int i = 0; i < nnodes; i++) {
        if (child_count[i] < 0 || child_count[i] > 65535 || child_begin[i] < 0 || child_begin[i] + child_count[i] > nchildren)
            return v_err(nullptr, ORBB_ERR_ARG, "node %d: bad child range", i);
    }
    for (int i = 0; i < nchildren; i++)
        if (child_list[i] <= 0 || child_list[i] >= nnodes) return v_err(nullptr, ORBB_ERR_ARG, "child list entry %d out of range", i);
    orbb_vocab* v = new orbb_vocab();
    v->device = device; v->nnodes = nnodes; v->depth = depth;
    // descriptors re-laid out in child order so that the children of one node are contiguous
    std::vector<uint8_t> cd((size_t)std::max(nchildren, 1) * 32);
    for (int i = 0; i < nchildren; i++) memcpy(&cd[(size_t)i * 32], node_desc + (size_t)child_list[i] * 32, 32);
#define VA(ptr, T, count, src)                                                                                        \
    if (cudaMalloc((void**)&ptr, std::max<size_t>((count), 1) * sizeof(T)) != cudaSuccess ||                           \
        cudaMemcpy(ptr, src, (size_t)(count) * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {                    \
        v_err(nullptr, ORBB_ERR_CUDA, "vocabulary upload failed: %s", cudaGetErrorString(cudaGetLastError()));        \
        orbb_vocab_destroy(v);                                                                                        \
        return ORBB_ERR_CUDA;                                                                                         \
    }
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete v;
        return v_err(nullptr, ORBB_ERR_CUDA, "cannot create stream on device %d", device);
    }
    VA(v->childBegin, int, nnodes, child_begin);
    VA(v->childCount, int, nnodes, child_count);
    VA(v->childId, int, nchildren, child_list);
    VA(v->childDesc, uint4, (size_t)nchildren * 2, cd.data());
    VA(v->weight, double, nnodes, node_weight);
    VA(v->wordId, int, nnodes, node_word_id);
#undef VA
    *out = v;
    return ORBB_OK;
}

void orbb_vocab_destroy(orbb_vocab* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    if (v->stream) { cudaStreamSynchronize(v->stream); cudaStreamDestroy(v->stream); }
    cudaFree(v->childBegin); cudaFree(v->childCount); cudaFree(v->childId); cudaFree(v->childDesc); cudaFree(v->weight);
    cudaFree(v->wordId); cudaFree(v->scratch);
    delete v;
}

const char* orbb_vocab_last_error(const orbb_vocab* v) { return v ? v->err.c_str() : g_lastError.c_str(); }
long long orbb_vocab_launch_count(const orbb_vocab* v) { return v ? v->launches : 0; }

int orbb_bow_transform(orbb_vocab* v, const uint8_t* desc, const int32_t* rowptr, int nsets, int levelsup, int norm, int32_t* bow_id,
                       double* bow_val, int32_t* fv_node, int32_t* fv_start, int32_t* fv_feat, int32_t* counts) {
    if (!v || !rowptr || !counts || nsets < 0 || (norm != 0 && norm != 1 && norm != 2)) return v_err(v, ORBB_ERR_ARG, "bad argument");
    if (nsets == 0) return ORBB_OK;
    const int total = rowptr[nsets];
    for (int s = 0; s < nsets; s++)
        if (rowptr[s + 1] < rowptr[s] || rowptr[s + 1] - rowptr[s] >= (1 << 24)) return v_err(v, ORBB_ERR_ARG, "bad rowptr at set %d", s);
    ORBV_CUDA(v, cudaSetDevice(v->device));
    // scratch layout (256-B aligned parts)
    const size_t T = (size_t)std::max(total, 1);
    const size_t sz[12] = {T * 32, (size_t)(nsets + 1) * 4, T * 4, T * 4, T * 8, T * 8, T * 8, T * 4, T * 8, T * 4, T * 4, T * 4};
    size_t off[13], cur = 0;
    for (int i = 0; i < 12; i++) { off[i] = cur; cur += (sz[i] + 255) / 256 * 256; }
    off[12] = cur; cur += ((size_t)nsets * 12 + 255) / 256 * 256;
    if (v->scratchCap < cur) {
        cudaFree(v->scratch);
        v->scratch = nullptr; v->scratchCap = 0;
        ORBV_CUDA(v, cudaMalloc(&v->scratch, cur));
        v->scratchCap = cur;
    }
    char* b = (char*)v->scratch;
    uint4* dDesc = (uint4*)(b + off[0]); int* dRow = (int*)(b + off[1]); int* dWord = (int*)(b + off[2]); int* dNode = (int*)(b + off[3]);
    double* dW = (double*)(b + off[4]); unsigned long long* dKW = (unsigned long long*)(b + off[5]);
    unsigned long long* dKN = (unsigned long long*)(b + off[6]); int* dBowId = (int*)(b + off[7]); double* dBowVal = (double*)(b + off[8]);
    int* dFvNode = (int*)(b + off[9]); int* dFvStart = (int*)(b + off[10]); int* dFvFeat = (int*)(b + off[11]); int* dCounts = (int*)(b + off[12]);
    if (total > 0) ORBV_CUDA(v, cudaMemcpyAsync(dDesc, desc, (size_t)total * 32, cudaMemcpyHostToDevice, v->stream));
    ORBV_CUDA(v, cudaMemcpyAsync(dRow, rowptr, (size_t)(nsets + 1) * 4, cudaMemcpyHostToDevice, v->stream));
    if (total > 0) {
        k_bow_descend<<<(total + 7) / 8, 256, 0, v->stream>>>(dDesc, total, v->childBegin, v->childCount, v->childId, v->childDesc, v->weight,
                                                             v->wordId, v->depth - levelsup, dWord, dNode, dW);
        v->launches++;
    }
    k_bow_assemble<<<nsets, 256, 0, v->stream>>>(dRow, dWord, dNode, dW, norm, dKW, dKN, dBowId, dBowVal, dFvNode, dFvStart, dFvFeat, dCounts);
    v->launches++;
    ORBV_CUDA(v, cudaGetLastError());
    if (total > 0) {
        if (bow_id) ORBV_CUDA(v, cudaMemcpyAsync(bow_id, dBowId, (size_t)total * 4, cudaMemcpyDeviceToHost, v->stream));
        if (bow_val) ORBV_CUDA(v, cudaMemcpyAsync(bow_val, dBowVal, (size_t)total * 8, cudaMemcpyDeviceToHost, v->stream));
        if (fv_node) ORBV_CUDA(v, cudaMemcpyAsync(fv_node, dFvNode, (size_t)total * 4, cudaMemcpyDeviceToHost, v->stream));
        if (fv_start) ORBV_CUDA(v, cudaMemcpyAsync(fv_start, dFvStart, (size_t)total * 4, cudaMemcpyDeviceToHost, v->stream));
        if (fv_feat) ORBV_CUDA(v, cudaMemcpyAsync(fv_feat, dFvFeat, (size_t)total * 4, cudaMemcpyDeviceToHost, v->stream));
    }
    ORBV_CUDA(v, cudaMemcpyAsync(counts, dCounts, (size_t)nsets * 12, cudaMemcpyDeviceToHost, v->stream));
    ORBV_CUDA(v, cudaStreamSynchronize(v->stream));
    return ORBB_OK;
}

}  // extern "C"

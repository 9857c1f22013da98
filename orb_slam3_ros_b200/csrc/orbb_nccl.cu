// liborbb200.so -- the one exchange step of the path: brute-force Hamming 2-NN against a database sharded over the GPUs of a box
// (BASELINE config 4; the semantics are those of cv::BFMatcher::knnMatch(q, t, 2) at reference orb_slam3/src/Frame.cc:1144-1151).
//
// Every rank scans all queries against its contiguous database shard (k_knn2_partial / k_knn2_merge_chunks, orbb_match.cu), ONE
// ncclAllGather moves the per-shard top-2 -- idx and dist packed into one [4 * nq] int32 block per rank, 3.2 MB at 200 k queries --
// over NVLink / NVSwitch, and k_knn2_merge_shards picks the global top-2 by lexicographic (distance, index) on every rank; with
// contiguous shards that reproduces BFMatcher's lowest-index tie rule exactly.
//
// NCCL is bound at run time (dlopen of the libnccl.so.2 the process already holds -- e.g. the one PyTorch ships -- else the system
// one), so the library has no link-time NCCL dependency and a caller's ncclComm_t, created with that same NCCL, can be passed in.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "orbb_internal.cuh"

struct orbb_matcher;
namespace orbb {
int matcher_device(const orbb_matcher* m);
cudaStream_t matcher_stream(const orbb_matcher* m);
int matcher_scratch(orbb_matcher* m, int slot, size_t bytes, void** out);
int matcher_error(orbb_matcher* m, int code, const char* msg);
void matcher_count_launches(orbb_matcher* m, int n);
int launch_merge_shards(orbb_matcher* m, const int32_t* idx_sh, const int32_t* dist_sh, size_t shard_stride, int nshards, int nq, int32_t* idx2,
                        int32_t* dist2);
}  // namespace orbb

namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string err;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // an NCCL that is already mapped into the process wins (the caller's communicators belong to it)
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names)
            if (!api.lib) api.lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        for (const char* n : names)
            if (!api.lib) api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) { api.err = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : ""); return; }
#define ORBB_SYM(field, name)                                              \
    api.field = (decltype(api.field))dlsym(api.lib, name);                  \
    if (!api.field) { api.err = std::string("missing symbol ") + name; api.lib = nullptr; return; }
        ORBB_SYM(GetUniqueId, "ncclGetUniqueId")
        ORBB_SYM(CommInitRank, "ncclCommInitRank")
        ORBB_SYM(CommDestroy, "ncclCommDestroy")
        ORBB_SYM(CommCount, "ncclCommCount")
        ORBB_SYM(CommUserRank, "ncclCommUserRank")
        ORBB_SYM(AllGather, "ncclAllGather")
        ORBB_SYM(GetErrorString, "ncclGetErrorString")
        ORBB_SYM(GetVersion, "ncclGetVersion")
#undef ORBB_SYM
    });
    return api;
}

int nccl_fail(orbb_matcher* m, const char* what, ncclResult_t r) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s failed: %s", what, nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
    return orbb::matcher_error(m, ORBB_ERR_CUDA, buf);
}

}  // namespace

extern "C" {

int orbb_nccl_version(void) {
    NcclApi& a = nccl();
    int v = 0;
    if (!a.lib || a.GetVersion(&v) != 0) return 0;
    return v;
}

int orbb_nccl_unique_id(void* id128) {
    NcclApi& a = nccl();
    if (!id128) return orbb::matcher_error(nullptr, ORBB_ERR_ARG, "null id buffer");
    if (!a.lib) return orbb::matcher_error(nullptr, ORBB_ERR_UNSUPPORTED, a.err.c_str());
    ncclUniqueId id;
    const ncclResult_t r = a.GetUniqueId(&id);
    if (r != 0) return nccl_fail(nullptr, "ncclGetUniqueId", r);
    memcpy(id128, &id, sizeof id);
    return ORBB_OK;
}

int orbb_nccl_comm_create(int device, int nranks, int rank, const void* id128, void** comm_out) {
    NcclApi& a = nccl();
    if (!id128 || !comm_out || nranks < 1 || rank < 0 || rank >= nranks) return orbb::matcher_error(nullptr, ORBB_ERR_ARG, "bad argument");
    if (!a.lib) return orbb::matcher_error(nullptr, ORBB_ERR_UNSUPPORTED, a.err.c_str());
    if (cudaSetDevice(device) != cudaSuccess) return orbb::matcher_error(nullptr, ORBB_ERR_CUDA, "cudaSetDevice failed");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t c = nullptr;
    const ncclResult_t r = a.CommInitRank(&c, nranks, id, rank);
    if (r != 0) return nccl_fail(nullptr, "ncclCommInitRank", r);
    *comm_out = c;
    return ORBB_OK;
}

void orbb_nccl_comm_destroy(void* comm) {
    if (comm && nccl().lib) nccl().CommDestroy((ncclComm_t)comm);
}

int orbb_knn2_sharded(orbb_matcher* m, void* nccl_comm, const uint8_t* q_dev, int nq, const uint8_t* db_shard_dev, int64_t nd_shard,
                      int32_t index_base, int32_t* idx2_dev, int32_t* dist2_dev) {
    if (!m || !nccl_comm || !idx2_dev || !dist2_dev || nq < 0 || nd_shard < 0) return orbb::matcher_error(m, ORBB_ERR_ARG, "bad argument");
    NcclApi& a = nccl();
    if (!a.lib) return orbb::matcher_error(m, ORBB_ERR_UNSUPPORTED, a.err.c_str());
    if (nq == 0) return ORBB_OK;
    ncclComm_t comm = (ncclComm_t)nccl_comm;
    int nranks = 0;
    ncclResult_t r = a.CommCount(comm, &nranks);
    if (r != 0 || nranks < 1) return nccl_fail(m, "ncclCommCount", r);
    if (cudaSetDevice(orbb::matcher_device(m)) != cudaSuccess) return orbb::matcher_error(m, ORBB_ERR_CUDA, "cudaSetDevice failed");
    // send block of this rank: [idx nq*2][dist nq*2]; receive area: nranks such blocks
    const size_t blockInts = (size_t)nq * 4;
    void *send = nullptr, *recv = nullptr;
    int rc;
    if ((rc = orbb::matcher_scratch(m, 2, blockInts * sizeof(int32_t), &send)) || (rc = orbb::matcher_scratch(m, 3, blockInts * sizeof(int32_t) * nranks, &recv)))
        return rc;
    int32_t* sIdx = (int32_t*)send;
    int32_t* sDist = sIdx + (size_t)nq * 2;
    if ((rc = orbb_knn2_dev(m, q_dev, nq, db_shard_dev, nd_shard, index_base, sIdx, sDist))) return rc;
    if (nranks == 1) return orbb::launch_merge_shards(m, sIdx, sDist, blockInts, 1, nq, idx2_dev, dist2_dev);
    r = a.AllGather(send, recv, blockInts, ncclInt32, comm, orbb::matcher_stream(m));
    if (r != 0) return nccl_fail(m, "ncclAllGather", r);
    const int32_t* rIdx = (const int32_t*)recv;
    return orbb::launch_merge_shards(m, rIdx, rIdx + (size_t)nq * 2, blockInts, nranks, nq, idx2_dev, dist2_dev);
}

}  // extern "C"

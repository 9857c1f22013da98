// Internal declarations shared by the extractor and matcher translation units of liborbb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/orbb200.h"

namespace orbb {

typedef unsigned long long u64;

constexpr int kEdge = 19;          // EDGE_THRESHOLD        (reference ORBextractor.cc:73)
constexpr int kRoiX = 32;          // column of the first image pixel inside a bordered row (>= kEdge, 16B aligned)
constexpr int kMinBorder = 16;     // EDGE_THRESHOLD - 3    (:789)
constexpr int kCellPix = 80;       // max side of one FAST cell incl. its 6-px apron
constexpr int kMaxIni = 16;        // max root nodes of the quadtree (image aspect ratio <= 16.5)

// One pyramid level: geometry + where its data lives inside the per-frame slabs.
struct LevelPlan {
    int w, h, pitch;               // image size, bytes per bordered row
    unsigned pyrOff, roiOff;       // byte offsets in the frame's pyramid slab: apron origin / first image pixel
    int bpitch;
    unsigned blurOff;              // blurred level (no apron)
    int tabX, tabY;                // offsets (int2 units) of the resize tables of this level
    int nCols, nRows, wCell, hCell, maxBX, maxBY;   // FAST cell grid (:789-803)
    int cellBase, cellCap;         // first cell index; key capacity per cell
    unsigned cellKeyBase;
    int nFeat, nIni;               // mnFeaturesPerLevel[l]; root nodes (:559)
    float hX;                      // :561
    unsigned rawBase;              // quadtree key ping-pong buffers
    int rawCap;
    unsigned nodeBase;
    int maxNodes;
    unsigned selBase;
    int selCap;
    float scale, kpSize;
    int blurTileBase, blurEdgeBase, blurTilesX, blurTilesY;      // k_blur: first CTA of the level in the interior / edge-column launch; word columns, 32-row strips
    int fastResize;                    // 1: the 4 source taps of every column group fit 3 aligned words (k_pyr_resize_s)
    int prmtTaps;                      // 1: the taps of the first three columns of every group lie in the first two of those words
};

// k_pyr_apron16 work decomposition of one level (16-byte chunks that contain apron bytes): `rows` apron rows above and below
// the image, `nLeft` chunks from chunk `leftChunk0` at the left end of every bordered row, `nRight` from `rightChunk0` at the right
struct ApronLevel { int itemBase, interiorChunks, rightChunk0, nRight, rows, leftChunk0, nLeft; unsigned invIC, invNR; };     // inv* = 2^32 / n + 1

struct Plan {
    int nlevels, W, H;
    int cellsTotal, blurTilesTotal, blurEdgeTotal;
    int cellTp, cellSmem, cellRows;   // k_fast_cell: tile pitch (64 / 96), dynamic shared memory per CTA, tile rows of the tallest cell
    int kpCap;
    int iniTh, minTh;              // clamped to 0..255
    unsigned k7Ini, k7Min;         // (0x7f - (th & 0x7f)) * 0x01010101: the byte-wise compare constant of the quick reject
    u64 pyrStride, blurStride;                                  // bytes per frame
    unsigned cellKeyStride, rawStride, nodeStride, selStride;   // entries per frame
    int umax[16];
    LevelPlan lv[ORBB_MAX_LEVELS];
    ApronLevel apron[ORBB_MAX_LEVELS];      // the 19-px apron of mvImagePyramid (built on demand: nothing in the pipeline reads it)
    int apronItems;
};

struct QNode {                     // quadtree node: UL=(x0,y0) BR=(x1,y1); keys = segment of a ping-pong buffer
    short x0, y0, x1, y1;
    int start;
    int cntbuf;                    // count | (buffer index << 31)
};

struct WorkItem { int level, x, y, pos; };

// FAST cell (host-built, shared by all frames): interior [gx0,gx1) x [gy0,gy1) in level coordinates (empty when the
// reference skips the cell, ORBextractor.cc:810,:819), first key slot of the cell inside the frame's cellKeys.
// The second half holds what the detector's state machine needs in every round, precomputed here because ptxas otherwise
// re-derives these loop invariants (as uniform-datapath instructions, but issue slots all the same) in every round: tile columns
// [cx0, cx1) of the interior (tile column 0 = level column (gx0 - 3) & ~15), first word `wa`, last word offset `wLast`, words per
// row pair `nwc` (>= 2) with its reciprocal, work items of phase A, byte masks of the first / last word.
struct __align__(16) CellDesc {
    short gx0, gx1, gy0, gy1; int level; unsigned outOff;
    short cx0, cx1, wa, wLast; int nwc, items; unsigned mInv, mF7, mL7, pad;
};

// One tensor map per pyramid level (k_fast_cell's tile loads): dims {pitch, bordered rows, frames} of the level's slab
struct TmapTable { CUtensorMap m[ORBB_MAX_LEVELS]; };

// Device buffers of one extractor handle, sized for `capacity` frames.
struct Bufs {
    uint8_t* pyr;
    uint8_t* blur;
    int2* tab;                     // resize tables (shared by all frames)
    const CellDesc* cellDesc;      // [cellsTotal]
    int* cellCount;
    int* cellOff;
    u64* cellKeys;                 // key = x | y<<16 | score<<32 (x,y relative to the 16-px border)
    u64* keys;                     // [frame][2][rawStride]
    QNode* nodes;                  // [frame][2][nodeStride]
    u64* rec;                      // [frame][nodeStride] sort records / scratch
    int4* cnt4;                    // [frame][nodeStride]
    int* pend;                     // [frame][2][nodeStride]
    int* elist;                    // [frame][nodeStride]
    uint8_t* erased;               // [frame][nodeStride]
    int* sortTmp;                  // [frame][3][nodeStride] scratch of the parallel std::sort emulation (large levels)
    u64* sel;                      // [frame][selStride] selected keys (level coordinates)
    int* selCount;                 // [frame][ORBB_MAX_LEVELS]
    WorkItem* work;                // [frame][kpCap]
    orbb_keypoint* kps;            // [frame][kpCap]
    uint8_t* desc;                 // [frame][kpCap][32]
    int* outCount;                 // [frame][2] {n, monoIndex}
    int* status;                   // [frame] error flags
    float* uRight;                 // [frame][kpCap]   stereo outputs (left handle only)
    float* depth;
    int* bestR;
    int* sad;
    uint4* stRec;                  // [frame][kpCap] stereo: this image's keypoints bucketed by row {x, minr | maxr << 16, octave, index}
    int* stRowStart;               // [frame][H + 2] first record of every row bucket
};

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy wrappers (PTX ISA 8.x, sm_90+; SASS: SYNCS.*, UBLKCP)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Programmatic dependent launch (PTX griddepcontrol, sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization
// may have its CTAs scheduled while the previous kernel of the stream is still draining its last wave.  pdl_launch_dependents() at the
// top of a kernel lets the NEXT kernel be scheduled as soon as every CTA of this one has started; pdl_wait() blocks until the
// PREVIOUS kernel has completed and its writes are visible -- it must come before the first read of anything that kernel produced.
// Both are no-ops in a kernel launched the ordinary way.  What this hides is the launch latency and the tail between the ~14
// dependent kernels of one batch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// one tile of a 3-D tensor map (x = byte within the bordered row, y = bordered row, z = frame) -> shared memory; SASS: UTMALDG.3D
__device__ __forceinline__ void tma_tensor3d_g2s(void* dst, const CUtensorMap* tmap, int x, int y, int z, void* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}

#endif

enum Stage { ST_H2D = 0, ST_PYRAMID, ST_FAST, ST_OCTREE, ST_BLUR, ST_ASSEMBLE, ST_ORIENT_DESC, ST_D2H, ST_COUNT };

}  // namespace orbb

struct orbb_extractor {
    orbb_params prm;
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> featPerLevel;
    int umax[16];
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t h2dStream = nullptr, d2hStream = nullptr;   // copy engines for the pipelined host path
    cudaEvent_t evH2D[8]{}, evDone[8]{};
    // execution lanes of the resident batch path: the frames of a batch are split over ORBB_LANES streams so that the
    // latency-bound kernels of one part (quadtree, descriptors) share the SMs with the issue-bound ones of another; inside a
    // lane the blur runs on a side stream beside the quadtree (both only read the pyramid).  Lane 0 uses `stream`.
    struct Lane { cudaStream_t st = nullptr, blurSt = nullptr, blurEdgeSt = nullptr; cudaEvent_t evFork = nullptr, evJoin = nullptr, evJoinEdge = nullptr, evStart = nullptr, evDone = nullptr; };
    Lane lanes[4];
    // one branch per pyramid level in a call with a few frames (run_lane)
    cudaStream_t lvlSt[ORBB_MAX_LEVELS] = {};
    cudaEvent_t evLvl[ORBB_MAX_LEVELS] = {}, evTree[ORBB_MAX_LEVELS] = {};
    int nLanes = 2;                // lanes of a resident batch (ORBB_LANES=1..4)
    // plan for the current image size
    orbb::Plan plan;
    orbb::Plan* dPlan = nullptr;
    orbb::TmapTable tmaps{};       // per-level tensor maps over b.pyr (valid with the plan)
    bool tmapsValid = false;
    bool planValid = false;
    int capacity = 0;              // frames the buffers are sized for
    orbb::Bufs b{};
    std::vector<void*> allocs;
    int lastFrames = 0;
    int pendingFrames = 0, pendingCapacity = 0;   // orbb_extract_batch_host_submit -> _wait
    long long launches = 0;
    bool profiling = false;
    bool capturing = false;        // run_batch is being recorded into the single-frame CUDA graph
    // single-frame calls (the SLAM thread's operator()) replay the launch sequence as a CUDA graph: 14 launches + the blur
    // fork/join cost more host time than the kernels of one frame take
    cudaGraphExec_t g1Exec = nullptr; int g1Lap0 = 0, g1Lap1 = 0; long long g1Launches = 0; bool g1Valid = false;
    cudaEvent_t ev[orbb::ST_COUNT + 1]{};
    bool evValid = false;
    bool stageRan[orbb::ST_COUNT]{};
    // staging: hImg is a DEVICE buffer for frames uploaded from the host; hPyr/hCounts are pinned host memory
    uint8_t* hImg = nullptr; size_t hImgBytes = 0;
    uint8_t* dColor = nullptr; size_t colorBytes = 0;      // device copy of a raw input frame (orbb_extract_color / _rectified / _resized)
    int2* rsTab = nullptr; int rsTabY = 0, rsSw = 0, rsSh = 0, rsDw = 0, rsDh = 0;      // cv::resize tables of orbb_extract_resized
    uint8_t* hPyr = nullptr; size_t hPyrBytes = 0; bool hPyrFresh = false;
    bool apronFull = false;          // the 19-px apron of the last extraction has been built (ensure_full_apron)
    int32_t* hCounts = nullptr; int hCountsCap = 0;
    std::string err;
};

namespace orbb {
int set_err(orbb_extractor* h, int code, const char* fmt, ...);
extern thread_local std::string g_lastError;
}

#define ORBB_CUDA(h, call)                                                                               \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return orbb::set_err(h, ORBB_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

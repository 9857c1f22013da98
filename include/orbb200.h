/*
 * orbb200.h -- C ABI of liborbb200.so: the B200 (sm_100a) ORB-SLAM3 front-end.
 *
 * This is the drop-in boundary for ONE hot path of giltchcity/orb_slam3_ros: ORB extraction
 * (ORB_SLAM3::ORBextractor::operator()), rectified-stereo matching (Frame::ComputeStereoMatches) and the
 * Hamming best / best-2 scans of ORBmatcher.  Plain C types only; no OpenCV, no torch.  The C++ adapter in
 * orb_slam3_ros_b200/host/ (ORBextractor.h / ORBmatcher.h, namespace ORB_SLAM3) sits on top of it and keeps
 * the reference's class interface so Frame.cc / Tracking.cc compile unchanged.
 *
 * Reference interfaces replaced (paths relative to the reference repository root):
 *   orbb_create / orbb_destroy      ORBextractor::ORBextractor            orb_slam3/src/ORBextractor.cc:409-469
 *   orbb_get_tables                 GetScaleFactors() & friends           orb_slam3/include/ORBextractor.h:61-82
 *   orbb_extract                    ORBextractor::operator()              orb_slam3/src/ORBextractor.cc:1086-1168
 *   orbb_pyramid_level              public member mvImagePyramid          orb_slam3/include/ORBextractor.h:84
 *   orbb_extract_batch*             (same operator(), many frames per launch; no reference counterpart)
 *   orbb_stereo_match               Frame::ComputeStereoMatches           orb_slam3/src/Frame.cc:811-981
 *   orbb_hamming_distance           ORBmatcher::DescriptorDistance        orb_slam3/src/ORBmatcher.cc:2058-2074
 *   orbb_best2_csr                  best / second-best candidate scans    orb_slam3/src/ORBmatcher.cc:77-120 (and :273-325,
 *                                                                         :1743-1768 ...: same loop shape)
 *   orbb_rgbd_stereo_batch          Frame::ComputeStereoFromRGBD   orb_slam3/src/Frame.cc:984-1005   ("next" row)
 *   orbb_rotation_check_csr         rotation histogram + ComputeThreeMaxima   orb_slam3/src/ORBmatcher.cc:345-352, :405-423, :2012-2053
 *   orbb_knn2 / orbb_knn2_partial   cv::BFMatcher(NORM_HAMMING).knnMatch(q, t, 2)   orb_slam3/src/Frame.cc:1144
 *   orbb_knn2_merge                 (top-2 merge of database shards after an all-gather; no reference counterpart)
 *   orbb_knn2_sharded               the same knnMatch against a database sharded over an NCCL communicator (BASELINE config 4)
 *   orbb_vocab_create / orbb_bow_transform   DBoW2 TemplatedVocabulary::transform via Frame::ComputeBoW   orb_slam3/src/Frame.cc:738-745,
 *                                   orb_slam3/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1139-1275 ("next" row)
 *   orbb_search_area_best2          Frame::GetFeaturesInArea + SearchByProjection scan   orb_slam3/src/Frame.cc:657-723, ORBmatcher.cc:71-120 ("next" row)
 *   orbb_distinctive_csr            MapPoint::ComputeDistinctiveDescriptors   orb_slam3/src/MapPoint.cc:329-403   ("next" row)
 *   orbb_extract_color / _batch_color   cv::cvtColor(..., COLOR_*2GRAY) + extraction   orb_slam3/src/Tracking.cc:1498-1525, :1605-1618 ("next" row)
 *   orbb_extract_resized / _batch_resized   cv::resize(im, imToFeed, newImSize) before tracking   orb_slam3/src/System.cc:241-244 ("next" row)
 *   orbb_rectifier_* / orbb_remap / orbb_extract_rectified / _batch_rectified   cv::remap(img, M1, M2, INTER_LINEAR) before tracking
 *                                   orb_slam3/src/System.cc:233-240, maps from orb_slam3/src/Settings.cc:506-509   ("next" row)
 *   orbb_undistort_points           Frame::UndistortKeyPoints (cv::undistortPoints)   orb_slam3/src/Frame.cc:747-780   ("next" row)
 *
 * Threading: one thread per handle at a time; distinct handles are independent (own stream, own workspace).
 * Errors: every function returns ORBB_OK (0) or a negative code; orbb_last_error() gives the text.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with ORBB_ERR_CUDA.
 */
#ifndef ORBB200_H
#define ORBB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORBB_OK 0
#define ORBB_ERR_EMPTY (-1)       /* empty image: mirrors "return -1" at ORBextractor.cc:1090-1091 */
#define ORBB_ERR_UNSUPPORTED (-2) /* geometry for which the reference itself has undefined behaviour */
#define ORBB_ERR_CAPACITY (-3)    /* caller buffer too small */
#define ORBB_ERR_ARG (-4)
#define ORBB_ERR_CUDA (-5)
#define ORBB_ERR_INTERNAL (-6)

#define ORBB_MAX_LEVELS 16

typedef struct orbb_extractor orbb_extractor;

/* Same field order as cv::KeyPoint minus class_id (24 bytes). */
typedef struct orbb_keypoint {
    float x, y;     /* pt, in level-0 pixel units (already multiplied by the level's scale factor) */
    float size;     /* (int)(31 * scale[octave])                      ORBextractor.cc:880 */
    float angle;    /* degrees in [0,360), IC_Angle + fastAtan2        ORBextractor.cc:76-103 */
    float response; /* FAST score (max arc minimum - 1)                 cv::FAST */
    int32_t octave;
} orbb_keypoint;

typedef struct orbb_params {
    int32_t nfeatures;    /* ORBextractor.nFeatures   */
    float scale_factor;   /* ORBextractor.scaleFactor */
    int32_t nlevels;      /* ORBextractor.nLevels  (<= ORBB_MAX_LEVELS) */
    int32_t ini_th_fast;  /* ORBextractor.iniThFAST */
    int32_t min_th_fast;  /* ORBextractor.minThFAST */
    int32_t device;       /* CUDA device ordinal */
    int32_t max_batch;    /* frames per batched launch the handle is sized for (>= 1) */
} orbb_params;

/* ---- life cycle ------------------------------------------------------------------------------------ */
int orbb_create(const orbb_params* params, orbb_extractor** out);
void orbb_destroy(orbb_extractor* h);
const char* orbb_last_error(const orbb_extractor* h); /* h may be NULL: last error of a handle-less call */
const char* orbb_version(void);

/* scale / inv_scale / sigma2 / inv_sigma2: nlevels floats each; feats_per_level: nlevels ints; any may be NULL */
int orbb_get_tables(const orbb_extractor* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                    int32_t* feats_per_level);
/* upper bound of keypoints one frame can yield (sum over levels of N_l + slack) */
int orbb_max_keypoints(const orbb_extractor* h);

/* ---- single frame, host buffers (the call ORBextractor::operator() makes) ------------------------------------- */
/* img: 8-bit gray, `stride` bytes per row.  kps/desc: caller buffers of `capacity` entries (desc = capacity*32 bytes).
 * Keypoints with lap0 <= x <= lap1 are written from the back, the others from the front (ORBextractor.cc:1153-1162);
 * *mono_index = number written at the front (the reference's return value).  Synchronous. */
int orbb_extract(orbb_extractor* h, const uint8_t* img, int width, int height, size_t stride, int lap0, int lap1,
                 orbb_keypoint* kps, uint8_t* desc, int capacity, int* n_out, int* mono_index);

/* Pyramid level of the most recent frame (frame 0 of the most recent batch), copied to a host buffer owned by the
 * handle (valid until the next extract call): *ptr points at the level's first pixel; the 19-pixel reflect-101 apron
 * lies around it at negative offsets exactly like the parent buffer of the reference's mvImagePyramid[level]. */
int orbb_pyramid_level(orbb_extractor* h, int level, const uint8_t** ptr, int* width, int* height, size_t* stride);

/* ---- batched, device-resident input (throughput path) ------------------------------------------------------ */
/* dev_imgs: device pointer; frame f, row y starts at dev_imgs + f*frame_stride + y*row_stride.  nframes <= max_batch.
 * Asynchronous on the handle's stream; results stay on the device until fetched. */
int orbb_extract_batch(orbb_extractor* h, const uint8_t* dev_imgs, int nframes, int width, int height, size_t row_stride,
                       size_t frame_stride, int lap0, int lap1);
/* Same, from HOST memory (pinned memory makes the copy asynchronous): H2D + kernels + D2H of the results into the
 * caller's host arrays, synchronous on return.  kps: nframes*capacity entries, desc: nframes*capacity*32 bytes,
 * counts: nframes*2 ints {n, mono_index}. */
int orbb_extract_batch_host(orbb_extractor* h, const uint8_t* host_imgs, int nframes, int width, int height,
                            size_t row_stride, size_t frame_stride, int lap0, int lap1, orbb_keypoint* kps, uint8_t* desc,
                            int capacity, int32_t* counts);
/* The same call split in two so that a caller can overlap batches on two handles (handle A uploads batch i+1 while
 * handle B computes batch i): _submit enqueues H2D + kernels + D2H and returns; the host buffers must stay valid until
 * _wait, which synchronises and fills counts. */
int orbb_extract_batch_host_submit(orbb_extractor* h, const uint8_t* host_imgs, int nframes, int width, int height,
                                   size_t row_stride, size_t frame_stride, int lap0, int lap1, orbb_keypoint* kps,
                                   uint8_t* desc, int capacity);
int orbb_extract_batch_host_wait(orbb_extractor* h, int32_t* counts);
/* Colour input ("next" row of the scope table): the cv::cvtColor(COLOR_{RGB,BGR,RGBA,BGRA}2GRAY) that Tracking applies
 * before building a Frame (Tracking.cc:1498-1525) runs on the device, then the normal extraction.  channels = 3 or 4,
 * rgb_order = 1 for RGB(A), 0 for BGR(A) (Tracking's mbRGB).  Fixed point of OpenCV 4.x: (R*9798 + G*19235 + B*3735 + 2^14) >> 15. */
int orbb_extract_color(orbb_extractor* h, const uint8_t* img, int width, int height, size_t stride, int channels, int rgb_order,
                       int lap0, int lap1, orbb_keypoint* kps, uint8_t* desc, int capacity, int* n_out, int* mono_index);
int orbb_extract_batch_color(orbb_extractor* h, const uint8_t* dev_imgs, int nframes, int width, int height, size_t row_stride,
                             size_t frame_stride, int channels, int rgb_order, int lap0, int lap1);
int orbb_sync(orbb_extractor* h);
/* CUDA stream (cudaStream_t) the handle launches on, for callers that time with their own events */
void* orbb_stream(orbb_extractor* h);
/* counts: nframes*2 ints {n, mono_index}; kps/desc laid out with `capacity` entries per frame.  Synchronises. */
int orbb_batch_fetch(orbb_extractor* h, int nframes, orbb_keypoint* kps, uint8_t* desc, int capacity, int32_t* counts);
/* device views of the last batch's results (per-frame stride = orbb_max_keypoints entries) */
int orbb_batch_device_ptrs(orbb_extractor* h, const orbb_keypoint** kps, const uint8_t** desc, const int32_t** counts);
/* number of kernels this handle launched since creation (bench.py's gpu_launches) */
long long orbb_launch_count(const orbb_extractor* h);
/* device milliseconds spent in each stage of the last batch (CUDA events on the handle's stream); stage names via
 * orbb_stage_name(i); returns the number of stages written (<= cap) */
int orbb_stage_times(orbb_extractor* h, float* ms, int cap);
const char* orbb_stage_name(int i);
int orbb_set_profiling(orbb_extractor* h, int enabled);

/* ---- stage taps for parity tests (frame index within the last batch) ------------------------------------------- */
/* pyramid (blurred=0) or blurred (blurred=1) level pixels -> dst[height*width] tightly packed; bordered=1 returns the
 * (w+38)x(h+38) apron-included buffer of the un-blurred level */
int orbb_debug_level(orbb_extractor* h, int frame, int level, int blurred, int bordered, uint8_t* dst, size_t dst_cap,
                     int* width, int* height);
/* vToDistributeKeys of a level (x, y relative to the 16-px border, response), in the reference's order */
int orbb_debug_raw_keys(orbb_extractor* h, int frame, int level, float* xyr, int cap, int* n);
/* allKeypoints[level] after DistributeOctTree, level coordinates */
int orbb_debug_selected(orbb_extractor* h, int frame, int level, float* xyr, int cap, int* n);

/* ---- stereo ---------------------------------------------------------------------------------------------------- */
/* Frame::ComputeStereoMatches for frame `frame` of the last batches of hL (left) and hR (right), whose pyramids and
 * keypoints/descriptors are still resident.  Outputs (host): u_right[nL], depth[nL] (-1 = no match).  bf = mbf,
 * b = mb (Frame.cc:841-843).  Optional best_r / sad (host, may be NULL) expose the intermediate decisions. */
int orbb_stereo_match(orbb_extractor* hL, orbb_extractor* hR, int frame, float bf, float b, float* u_right, float* depth,
                      int32_t* best_r, int32_t* sad, int capacity, int* n_left);
/* batched: all frames [0,nframes) of the last batch; results stay on the device (see orbb_stereo_fetch) */
int orbb_stereo_match_batch(orbb_extractor* hL, orbb_extractor* hR, int nframes, float bf, float b);
int orbb_stereo_fetch(orbb_extractor* hL, int nframes, float* u_right, float* depth, int capacity);
/* RGB-D counterpart ("next" row): Frame::ComputeStereoFromRGBD (orb_slam3/src/Frame.cc:984-1005) over the keypoints of the
 * extractor's last batch.  dev_depth: device-resident depth maps of the frames (float32, or uint16 scaled by depth_factor like
 * Tracking::GrabImageRGBD's convertTo, Tracking.cc:1576-1577); strides in bytes.  K4 = {fx, fy, cx, cy}, dist = the ndist
 * coefficients of mDistCoef (the undistorted x of Frame::UndistortKeyPoints enters mvuRight); bf = mbf.  Asynchronous on the
 * handle's stream; results via orbb_stereo_fetch. */
int orbb_rgbd_stereo_batch(orbb_extractor* h, const void* dev_depth, int depth_is_u16, float depth_factor, size_t row_stride,
                           size_t frame_stride, int nframes, const float* K4, const float* dist, int ndist, float bf);

/* ---- Hamming matching ----------------------------------------------------------------------------------------- */
/* host-side 256-bit Hamming distance (a single pair never goes to the GPU) */
int orbb_hamming_distance(const uint8_t* a, const uint8_t* b);

typedef struct orbb_matcher orbb_matcher;
int orbb_matcher_create(int device, orbb_matcher** out);
void orbb_matcher_destroy(orbb_matcher* m);
const char* orbb_matcher_last_error(const orbb_matcher* m);
long long orbb_matcher_launch_count(const orbb_matcher* m);
/* CUDA stream (cudaStream_t) the matcher launches on, for callers that time with their own events */
void* orbb_matcher_stream(orbb_matcher* m);

/* Brute-force 2-NN.  q: nq x 32 bytes, db: nd x 32 bytes.  idx2/dist2: nq x 2 (best, second); a missing neighbour is
 * (-1, INT32_MAX).  Ties resolve to the lowest database index (cv::BFMatcher behaviour).  *_dev take device pointers
 * and run asynchronously on the matcher's stream; the host variant copies in/out and synchronises.
 * index_base is added to every returned index (shard offset). */
int orbb_knn2(orbb_matcher* m, const uint8_t* q, int nq, const uint8_t* db, int64_t nd, int32_t* idx2, int32_t* dist2);
int orbb_knn2_dev(orbb_matcher* m, const uint8_t* q_dev, int nq, const uint8_t* db_dev, int64_t nd, int32_t index_base,
                  int32_t* idx2_dev, int32_t* dist2_dev);
/* merge nshards per-shard results (layout [shard][nq][2], e.g. straight out of an all-gather) into the global top-2
 * by lexicographic (distance, index); output nq x 2 */
int orbb_knn2_merge_dev(orbb_matcher* m, const int32_t* idx_sh_dev, const int32_t* dist_sh_dev, int nshards, int nq,
                        int32_t* idx2_dev, int32_t* dist2_dev);
/* The same 2-NN against a database SHARDED over the ranks of an NCCL communicator (BASELINE config 4; no reference call site at
 * that size -- the semantics are those of knnMatch at Frame.cc:1144-1151).  Every rank passes all nq queries (replicated) and ITS
 * contiguous shard of the database (nd_shard rows whose first row has the global index index_base); the per-shard top-2 of all
 * ranks travel in ONE ncclAllGather (idx and dist packed, 16 bytes per query and rank) and are merged on the device by
 * lexicographic (distance, index), which with contiguous shards is BFMatcher's lowest-index tie rule.  idx2_dev / dist2_dev
 * (nq x 2, device) hold the global result on every rank.  nccl_comm is the caller's ncclComm_t; asynchronous on the matcher's
 * stream.  NCCL is bound at run time to the libnccl.so.2 already loaded in the process (else the system one);
 * ORBB_ERR_UNSUPPORTED when there is none. */
int orbb_knn2_sharded(orbb_matcher* m, void* nccl_comm, const uint8_t* q_dev, int nq, const uint8_t* db_shard_dev, int64_t nd_shard,
                      int32_t index_base, int32_t* idx2_dev, int32_t* dist2_dev);
/* For hosts that do not link NCCL themselves: ncclGetUniqueId (id128 = 128 bytes, made on one rank and handed to the others by
 * any means), ncclCommInitRank on `device`, ncclCommDestroy; orbb_nccl_version() = ncclGetVersion() of the bound library, 0 = none. */
int orbb_nccl_version(void);
int orbb_nccl_unique_id(void* id128);
int orbb_nccl_comm_create(int device, int nranks, int rank, const void* id128, void** comm_out);
void orbb_nccl_comm_destroy(void* comm);
/* Lowe ratio test of Frame.cc:1151 on device results: keep[i] = (second exists) && dist0 < dist1 * ratio (double) */
int orbb_ratio_test_dev(orbb_matcher* m, const int32_t* idx2_dev, const int32_t* dist2_dev, int nq, double ratio,
                        uint8_t* keep_dev);

/* best / second-best scan over per-query candidate lists (CSR): query i scans cand[rowptr[i]..rowptr[i+1]) (row
 * indices into train), strict '<' in list order, both distances start at `init` (256 / TH_LOW / INT_MAX ...).
 * out4[i] = {bestDist, bestIdx, secondDist, secondIdx} (idx -1 if none).  Host pointers; synchronous. */
int orbb_best2_csr(orbb_matcher* m, const uint8_t* q, int nq, const uint8_t* train, int ntrain, const int32_t* cand,
                   const int32_t* rowptr, int init, int32_t* out4);

/* Frame::GetFeaturesInArea (Frame.cc:657-723, 64x48 grid of Frame::AssignFeaturesToGrid :385-416) fused with the best /
 * second-best scan of ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&) (ORBmatcher.cc:71-120), all projected map
 * points of a frame in one call.  Host pointers; synchronous.
 *   kps_xy  n x {x, y}   undistorted keypoints (Frame::mvKeysUn);  octaves n;  train n x 32 (Frame::mDescriptors)
 *   grid4   {mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv}
 *   queries nq x {x, y, r, projXR};  qlev nq x {minLevel, maxLevel};  qdesc nq x 32 (MapPoint::GetDescriptor)
 *   skip    optional n bytes: 1 = keypoint already bound to an observed MapPoint (ORBmatcher.cc:88-90)
 *   u_right optional n floats: Frame::mvuRight; when > 0 the candidate is dropped if |projXR - uRight| > r (:92-97)
 *   out4[i] = {bestDist, bestIdx, secondDist, secondIdx}, distances start at `init`, first minimum in the reference's
 *   candidate order (cell column, cell row, keypoint index) wins ties. */
int orbb_search_area_best2(orbb_matcher* m, const float* kps_xy, const int32_t* octaves, const uint8_t* train, int n, const float* grid4,
                           const float* queries, const int32_t* qlev, const uint8_t* qdesc, int nq, const uint8_t* skip,
                           const float* u_right, int init, int32_t* out4);

/* ---- the same scans with the FRAME side resident on the device ----------------------------------------------------------------
 * A tracked frame's key points and descriptors are produced on the GPU by the extractor and read by every Search* call made on
 * that frame; with a frame view they are uploaded at most once (orbb_frame_upload) or not at all (on_device = 1 with the
 * extractor's own buffers from orbb_batch_device_ptrs: x, y at offset 0 and octave at offset 20 of the 24-byte orbb_keypoint
 * records -- valid for frames without lens distortion, where Frame::mvKeysUn == mvKeys, Frame.cc:749-753).  The query side
 * (map-point descriptors and projections: host state of the SLAM threads) goes up per call, the per-query result comes back. */
#define ORBB_FRAME_SLOTS 4
typedef struct orbb_frame_view {
    const void* kps_xy;      /* key point i: float x, y at kps_xy + i * kps_stride      (Frame::mvKeysUn[i].pt) */
    size_t kps_stride;       /* bytes, >= 8 */
    const void* octaves;     /* key point i: int32 octave at octaves + i * oct_stride    (Frame::mvKeysUn[i].octave) */
    size_t oct_stride;       /* bytes, >= 4 */
    const uint8_t* desc;     /* n x 32                                                   (Frame::mDescriptors) */
    const float* u_right;    /* n floats or NULL                                         (Frame::mvuRight) */
    int32_t n;
    int32_t on_device;       /* 1: the pointers above are device pointers; 0: host pointers (uploaded by the call) */
} orbb_frame_view;
/* copies a host-resident frame to the device once (slot 0 .. ORBB_FRAME_SLOTS-1 of the matcher; a slot is overwritten by the next
 * upload into it) and returns the device view to pass to the scans below; asynchronous on the matcher's stream */
int orbb_frame_upload(orbb_matcher* m, int slot, const orbb_frame_view* host, orbb_frame_view* dev);
/* orbb_search_area_best2 generalised: the k (1, 2, 4 or 8) best candidates of every query in the reference's scan order (distance,
 * then cell column, cell row, key point index), out = nq x k x {dist, idx} (missing entries = {init, -1}).  k = 4 lets the host apply
 * the reference's in-order decisions (a key point matched by an earlier map point of the call is skipped by the later ones,
 * ORBmatcher.cc:88-90, :1749-1751) without another scan: dropping taken candidates from the sorted list leaves the same first two. */
int orbb_search_area_topk(orbb_matcher* m, const orbb_frame_view* frame, const float* grid4, const float* queries, const int32_t* qlev,
                          const uint8_t* qdesc, int nq, const uint8_t* skip, int init, int k, int32_t* out);
/* orbb_best2_csr with the train descriptors (Frame::mDescriptors) already on the device */
int orbb_best2_csr_dev(orbb_matcher* m, const uint8_t* q, int nq, const uint8_t* train_dev, int ntrain, const int32_t* cand,
                       const int32_t* rowptr, int init, int32_t* out4);

/* MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:329-403) for many map points at once: group g holds the
 * descriptors desc[rowptr[g] .. rowptr[g+1]) (32 bytes each, host memory); best[g] = index within the group of the
 * descriptor with the least median Hamming distance to the others (first minimum), -1 for an empty group. */
int orbb_distinctive_csr(orbb_matcher* m, const uint8_t* desc, int ntotal, const int32_t* rowptr, int ngroups, int32_t* best);

/* ---- bag of words ("next" row) ------------------------------------------------------------------------------------ */
/* Vocabulary tree as flat arrays (what an adapter reads out of DBoW2's m_nodes): node 0 is the root; the children of node
 * i are child_list[child_begin[i] .. child_begin[i] + child_count[i]) in DBoW2's visiting order; a node without children
 * is a leaf (= word) with word id node_word_id[i] and tf-idf weight node_weight[i]; depth = m_L. */
typedef struct orbb_vocab orbb_vocab;
int orbb_vocab_create(int device, int nnodes, const int32_t* child_begin, const int32_t* child_count, const int32_t* child_list,
                      int nchildren, const uint8_t* node_desc, const double* node_weight, const int32_t* node_word_id, int depth,
                      orbb_vocab** out);
void orbb_vocab_destroy(orbb_vocab* v);
const char* orbb_vocab_last_error(const orbb_vocab* v);
long long orbb_vocab_launch_count(const orbb_vocab* v);
/* TemplatedVocabulary::transform(features, BowVector, FeatureVector, levelsup) for nsets descriptor sets at once (set s =
 * descriptors rowptr[s] .. rowptr[s+1], 32 bytes each, host memory).  norm: 1 = L1 (ORB-SLAM3's vocabulary), 2 = L2,
 * 0 = none.  Outputs are laid out per set at offset rowptr[s] (capacity = the set's feature count):
 *   bow_id / bow_val   the BowVector in ascending word order, counts[3s] entries
 *   fv_node / fv_start the FeatureVector in ascending node order, counts[3s+1] entries; node k owns
 *                      fv_feat[rowptr[s] + fv_start[k] .. next start or counts[3s+2]) (feature indices within the set, ascending)
 *   counts[3s+2]       features with a positive word weight ("not stopped") */
int orbb_bow_transform(orbb_vocab* v, const uint8_t* desc, const int32_t* rowptr, int nsets, int levelsup, int norm, int32_t* bow_id,
                       double* bow_val, int32_t* fv_node, int32_t* fv_start, int32_t* fv_feat, int32_t* counts);

/* ---- input-side resize ("next" row) -------------------------------------------------------------------------------- */
/* cv::resize(im, imToFeed, newImSize) [INTER_LINEAR] that System::TrackStereo / TrackRGBD / TrackMonocular apply when the settings
 * request another image size (orb_slam3/src/System.cc:241-244, :312-318, :383-388), on the device, followed by the extraction
 * of the resized frame(s).  Batch: device-resident raw frames, asynchronous (results via orbb_batch_fetch); single: host frame. */
int orbb_extract_batch_resized(orbb_extractor* h, const uint8_t* dev_imgs, int nframes, int width, int height, size_t row_stride,
                               size_t frame_stride, int new_width, int new_height, int lap0, int lap1);
int orbb_extract_resized(orbb_extractor* h, const uint8_t* img, int width, int height, size_t stride, int new_width, int new_height, int lap0,
                         int lap1, orbb_keypoint* kps, uint8_t* desc, int capacity, int* n_out, int* mono_index);

/* ---- stereo rectification ("next" row) ---------------------------------------------------------------------------- */
/* cv::remap(src, dst, map_x, map_y, cv::INTER_LINEAR) for 8-bit single-channel images with CV_32FC1 maps (what
 * cv::initUndistortRectifyMap(..., CV_32F, M1, M2) returns, Settings.cc:506-509) and the default constant (0) border.  The
 * maps are converted to OpenCV's fixed-point form once, at creation.  map_stride is in floats. */
typedef struct orbb_rectifier orbb_rectifier;
int orbb_rectifier_create(int device, const float* map_x, const float* map_y, size_t map_stride, int dst_width, int dst_height,
                          int src_width, int src_height, orbb_rectifier** out);
void orbb_rectifier_destroy(orbb_rectifier* r);
/* one host frame in, the rectified host frame out (System.cc:239) */
int orbb_remap(orbb_rectifier* r, const uint8_t* src, size_t src_stride, uint8_t* dst, size_t dst_stride);
/* rectification + extraction without the rectified image leaving the device: batch of device-resident raw frames
 * (asynchronous, results via orbb_batch_fetch) / one host frame (synchronous) */
int orbb_extract_batch_rectified(orbb_extractor* h, orbb_rectifier* r, const uint8_t* dev_imgs, int nframes, size_t row_stride,
                                 size_t frame_stride, int lap0, int lap1);
int orbb_extract_rectified(orbb_extractor* h, orbb_rectifier* r, const uint8_t* img, size_t stride, int lap0, int lap1, orbb_keypoint* kps,
                           uint8_t* desc, int capacity, int* n_out, int* mono_index);

/* ---- rotation-consistency filter of the match scans ---------------------------------------------------------------- */
/* ORBmatcher.cc:345-352 (votes), :405-423 (discard), ComputeThreeMaxima :2012-2053, for nsets independent match sets at once
 * (set s = matches rowptr[s] .. rowptr[s+1]): angle_a / angle_b = the keypoint angles of the two sides of every match;
 * keep[i] = 1 when match i falls into one of the three fullest rotation bins of its set; ind3[3s..3s+2] = those bins (-1 = none). */
int orbb_rotation_check_csr(orbb_matcher* m, const float* angle_a, const float* angle_b, int total, const int32_t* rowptr, int nsets,
                            uint8_t* keep, int32_t* ind3);

/* ---- Frame::UndistortKeyPoints ("next" row) ------------------------------------------------------------------------- */
/* cv::undistortPoints(xy, out, K, dist, noArray(), newK) for n points (Frame.cc:766: newK = K).  K4 / newK4 = {fx, fy, cx, cy};
 * dist = ndist (0..12) coefficients k1 k2 p1 p2 [k3 [k4 k5 k6 [s1 s2 s3 s4]]].  dist[0] == 0 copies the input (Frame.cc:749). */
int orbb_undistort_points(orbb_matcher* m, const float* xy, int n, const float* K4, const float* dist, int ndist, const float* newK4,
                          float* out_xy);

/* pinned host memory helpers (so callers without a CUDA runtime can stage asynchronously) */
void* orbb_host_alloc(size_t bytes);
void orbb_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* ORBB200_H */

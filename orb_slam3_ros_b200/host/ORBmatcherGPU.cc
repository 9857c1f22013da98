// GPU-backed replacements for the two ORBmatcher::SearchByProjection overloads that Tracking calls on every frame:
//
//   SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th, bFarPoints, thFarPoints)     reference ORBmatcher.cc:43-213
//       (Tracking::SearchLocalPoints, Tracking.cc:3216 ff.)
//   SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)                           reference ORBmatcher.cc:1676-1887
//       (Tracking::TrackWithMotionModel, Tracking.cc:2925 / :2933)
//
//   SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist)                   reference ORBmatcher.cc:1889-2010
//       (Tracking::Relocalization, Tracking.cc:3765 / :3779)
//   SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches)                                              reference ORBmatcher.cc:223-421
//       (Tracking::TrackReferenceKeyFrame, Tracking.cc:2769; Tracking::Relocalization, :3687)
//
//   SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, vpPoints, vpMatched, th, ratioHamming)         reference ORBmatcher.cc:427-530
//       (LoopClosing.cc:1795 / :1982)
//   SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse)                              reference ORBmatcher.cc:906-1146
//       (LocalMapping::CreateNewMapPoints, LocalMapping.cc:466)
//   Fuse(KeyFrame* pKF, vpMapPoints, th, bRight = false)                                                 reference ORBmatcher.cc:1148-1338
//       (LocalMapping::SearchInNeighbors, LocalMapping.cc:772 / :802)
//   Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, vpPoints, th, vpReplacePoint)                                reference ORBmatcher.cc:1340-1455
//       (LoopClosing::SearchAndFuse, LoopClosing.cc:3464 / :3509)
//   SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12)                                             reference ORBmatcher.cc:765-905
//       (LoopClosing::DetectCommonRegionsFromBoW, LoopClosing.cc:1680 -- the one call here that is not made by the Tracking thread)
//
// and for the matcher of the monocular initialisation, which Tracking calls on every frame until the map exists:
//
//   SearchForInitialization(Frame& F1, Frame& F2, vbPrevMatched, vnMatches12, windowSize)                reference ORBmatcher.cc:648-766
//       (Tracking::MonocularInitialization, Tracking.cc:2527)
//
// for monocular, rectified-stereo and RGB-D frames (Frame::Nleft == -1).  What stays on the host is what is host state in the
// reference as well: the MapPoint / Frame objects, the projection of every point, the accept / reject decisions.  What moves to the
// GPU is the part that costs: Frame::GetFeaturesInArea (Frame.cc:657-723) + the DescriptorDistance scan of every candidate
// (orbb_search_area_topk: all points of the call in one launch, the frame's key points and descriptors uploaded once per frame).
//
// The reference walks its points in order and updates F.mvpMapPoints as it goes; a key point that has just received a map point with
// observations is skipped by every later point (:88-90, :1749-1751).  The batched scan sees the frame as it was BEFORE the call, so
// it returns the FOUR best candidates of every point in the reference's scan order and the decision loop below drops the ones that
// were taken meanwhile: what remains at the head of the list is what the reference's scan would have found.  Only when all four are
// gone does a point go back to the device, alone, with the current mask.
//
// Integration: at the top of the two reference methods
//     if (F.Nleft == -1) return ORBmatcherGPU::Instance().SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints, mfNNratio);
// Compiled here against tests/host/slam_stub (this image has no Eigen / Sophus / OpenCV headers) and compared on the GPU with the
// reference's own bodies (oracle/_ref, cut out of ORBmatcher.cc at build time): tests/test_gpu_matcher_host.py.
#include "ORBmatcherGPU.h"

#include <cmath>
#include <cstring>
#include <set>

#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"

namespace ORB_SLAM3 {

namespace {
const int kTopK = 4;

struct Cand { int dist, idx; };

// the frame side of the scans, uploaded once per (frame id, slot) and reused by later calls on the same frame
struct FrameOnDevice {
    unsigned long id = ~0ul;
    int n = -1;
    const void* descData = nullptr;
    orbb_frame_view view;
};
}  // namespace

struct ORBmatcherGPU::Impl {
    FrameOnDevice slot[2];
    std::vector<float> xy, q, ur;
    std::vector<int32_t> oct, qlev, out;
    std::vector<unsigned char> qdesc, skip;
    std::vector<int> src;
    std::vector<int32_t> cand, rowptr;
};

ORBmatcherGPU::Impl& ORBmatcherGPU::Scratch() {
    if (!mpImpl) {
        mpImpl = new Impl();
        mpImplFree = [](Impl* p) { delete p; };
    }
    return *mpImpl;
}

ORBmatcherGPU& ORBmatcherGPU::Instance(int device) {
    static thread_local ORBmatcherGPU inst(device);      // one matcher (stream + scratch) per SLAM thread, like the stack objects of the reference
    return inst;
}

// the frame's undistorted key points, octaves, descriptors and mvuRight on the device (slot 0: current frame, 1: any other)
static const orbb_frame_view* FrameView(orbb_matcher* m, ORBmatcherGPU::Impl& s, const Frame& F, int slotIdx) {
    FrameOnDevice& c = s.slot[slotIdx];
    if (c.id == F.mnId && c.n == F.N && c.descData == (const void*)F.mDescriptors.data) return &c.view;
    s.xy.resize((size_t)F.N * 2);
    s.oct.resize(F.N);
    for (int i = 0; i < F.N; i++) { s.xy[2 * i] = F.mvKeysUn[i].pt.x; s.xy[2 * i + 1] = F.mvKeysUn[i].pt.y; s.oct[i] = F.mvKeysUn[i].octave; }
    orbb_frame_view h;
    h.kps_xy = s.xy.data(); h.kps_stride = 8; h.octaves = s.oct.data(); h.oct_stride = 4;
    h.desc = F.mDescriptors.ptr<uchar>(); h.u_right = (int)F.mvuRight.size() == F.N ? F.mvuRight.data() : nullptr;
    h.n = F.N; h.on_device = 0;
    if (!F.mDescriptors.isContinuous()) throw std::runtime_error("Frame::mDescriptors must be continuous");
    if (orbb_frame_upload(m, slotIdx, &h, &c.view) != ORBB_OK)
        throw std::runtime_error(std::string("orbb_frame_upload failed: ") + orbb_matcher_last_error(m));
    c.id = F.mnId; c.n = F.N; c.descData = F.mDescriptors.data;
    return &c.view;
}

// which key points a scan leaves out, evaluated on the live objects: the two per-frame searches skip key points whose map point has
// observations (ORBmatcher.cc:88-90 / :1749-1751), the relocalisation search every key point that holds a map point (:1952)
static inline bool Taken(const Frame& F, int idx, bool any = false) {
    MapPoint* p = F.mvpMapPoints[idx];
    return p && (any || p->Observations() > 0);
}

// first `want` candidates of query j that are still free; false when the list is exhausted although it was full (more candidates
// may exist beyond the k returned: the caller scans that one query again)
static bool LiveHead(const Frame& F, const int32_t* list, int k, int want, Cand* out, int& nout, bool any = false) {
    nout = 0;
    int valid = 0;
    for (int t = 0; t < k; t++) {
        const int idx = list[2 * t + 1];
        if (idx < 0) break;
        valid++;
        if (Taken(F, idx, any)) continue;
        if (nout < want) { out[nout].dist = list[2 * t]; out[nout].idx = idx; nout++; }
    }
    return nout >= want || valid < k;
}

void ORBmatcherGPU::RunScan(const Frame& F, int nq, int k, std::vector<int32_t>& out, bool maskTaken, bool rightCheck, int init, bool takenAny) {
    Impl& s = Scratch();
    orbb_frame_view fv = *FrameView(mpMatcher, s, F, 0);
    if (!rightCheck) fv.u_right = nullptr;                 // (a scan without the mvuRight test of ORBmatcher.cc:94-100 / :1755-1761)
    s.skip.assign(F.N, 0);
    if (maskTaken) for (int i = 0; i < F.N; i++) s.skip[i] = Taken(F, i, takenAny);
    const float grid4[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
    out.assign((size_t)nq * k * 2, -1);
    if (nq == 0) return;
    if (orbb_search_area_topk(mpMatcher, &fv, grid4, s.q.data(), s.qlev.data(), s.qdesc.data(), nq, s.skip.data(), init, k, out.data()) != ORBB_OK)
        throw std::runtime_error(std::string("orbb_search_area_topk failed: ") + orbb_matcher_last_error(mpMatcher));
}

// one query again, with the frame as it is now
void ORBmatcherGPU::Rescan(const Frame& F, int j, int want, void* candOut, int& nout, bool rightCheck, bool takenAny) {
    Impl& s = Scratch();
    orbb_frame_view view = *FrameView(mpMatcher, s, F, 0);
    if (!rightCheck) view.u_right = nullptr;
    const orbb_frame_view* fv = &view;
    for (int i = 0; i < F.N; i++) s.skip[i] = Taken(F, i, takenAny);
    const float grid4[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
    int32_t o[4] = {256, -1, 256, -1};
    if (orbb_search_area_topk(mpMatcher, fv, grid4, &s.q[4 * (size_t)j], &s.qlev[2 * (size_t)j], &s.qdesc[32 * (size_t)j], 1, s.skip.data(), 256, 2, o) != ORBB_OK)
        throw std::runtime_error(std::string("orbb_search_area_topk failed: ") + orbb_matcher_last_error(mpMatcher));
    Cand* c = (Cand*)candOut;
    nout = 0;
    for (int t = 0; t < 2 && t < want; t++)
        if (o[2 * t + 1] >= 0) { c[nout].dist = o[2 * t]; c[nout].idx = o[2 * t + 1]; nout++; }
    mnRescans++;
}

static inline float RadiusByViewingCos(float viewCos) { return viewCos > 0.998f ? 2.5f : 4.0f; }      // ORBmatcher.cc:215-221

// The same search on a fisheye-stereo frame (Nleft != -1; KannalaBrandt8 rigs): two key point sets with their own grids -- mvKeys[0 .. Nleft) and
// mvKeysRight, descriptors stacked left then right -- and a map point can be visible in either eye (mbTrackInView / mbTrackInViewR) with its own
// projection, level and window.  Two batched scans (left queries against the left set, right queries against the right set, no mvuRight test:
// :94), then the reference's loop in order: the left decision, which also claims the stereo partner in the right set (:127-131), then the right
// decision, which also claims the partner in the left set (:196-200); a map point whose left ratio test fails skips its right half too (the
// `continue` at :122 leaves the outer loop body).
int ORBmatcherGPU::SearchByProjectionFisheye(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th, const bool bFarPoints,
                                             const float thFarPoints, const float nnratio) {
    Impl& s = Scratch();
    const bool bFactor = th != 1.0;
    const int nL = F.Nleft, nR = (int)F.mvKeysRight.size();
    if (!F.mDescriptors.isContinuous()) throw std::runtime_error("Frame::mDescriptors must be continuous");
    const float grid4[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
    // queries of the two eyes; qOf*[iMP] = query index or -1
    std::vector<float> qL, qR;
    std::vector<int32_t> levL, levR, outL, outR;
    std::vector<unsigned char> dL, dR, skipL, skipR;
    std::vector<int> qOfL(vpMapPoints.size(), -1), qOfR(vpMapPoints.size(), -1);
    for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
        MapPoint* pMP = vpMapPoints[iMP];
        if (!pMP->mbTrackInView && !pMP->mbTrackInViewR) continue;
        if (bFarPoints && pMP->mTrackDepth > thFarPoints) continue;
        if (pMP->isBad()) continue;
        const cv::Mat d = pMP->GetDescriptor();
        if (pMP->mbTrackInView) {
            const int nPredictedLevel = pMP->mnTrackScaleLevel;
            float r = RadiusByViewingCos(pMP->mTrackViewCos);
            if (bFactor) r *= th;
            const float q4[4] = {pMP->mTrackProjX, pMP->mTrackProjY, r * F.mvScaleFactors[nPredictedLevel], -1.0f};
            qOfL[iMP] = (int)levL.size() / 2;
            qL.insert(qL.end(), q4, q4 + 4);
            levL.push_back(nPredictedLevel - 1); levL.push_back(nPredictedLevel);
            dL.insert(dL.end(), d.ptr<uchar>(), d.ptr<uchar>() + 32);
        }
        if (pMP->mbTrackInViewR && pMP->mnTrackScaleLevelR != -1) {
            const int nPredictedLevel = pMP->mnTrackScaleLevelR;
            float r = RadiusByViewingCos(pMP->mTrackViewCosR);                    // (:144: no th factor on the right window)
            const float q4[4] = {pMP->mTrackProjXR, pMP->mTrackProjYR, r * F.mvScaleFactors[nPredictedLevel], -1.0f};
            qOfR[iMP] = (int)levR.size() / 2;
            qR.insert(qR.end(), q4, q4 + 4);
            levR.push_back(nPredictedLevel - 1); levR.push_back(nPredictedLevel);
            dR.insert(dR.end(), d.ptr<uchar>(), d.ptr<uchar>() + 32);
        }
    }
    std::vector<float> xyL((size_t)nL * 2), xyR((size_t)nR * 2);
    std::vector<int32_t> octL(nL), octR(nR);
    for (int i = 0; i < nL; i++) { xyL[2 * i] = F.mvKeys[i].pt.x; xyL[2 * i + 1] = F.mvKeys[i].pt.y; octL[i] = F.mvKeys[i].octave; }
    for (int i = 0; i < nR; i++) { xyR[2 * i] = F.mvKeysRight[i].pt.x; xyR[2 * i + 1] = F.mvKeysRight[i].pt.y; octR[i] = F.mvKeysRight[i].octave; }
    orbb_frame_view vL, vR;
    vL.kps_xy = xyL.data(); vL.kps_stride = 8; vL.octaves = octL.data(); vL.oct_stride = 4; vL.desc = F.mDescriptors.ptr<uchar>(); vL.u_right = nullptr;
    vL.n = nL; vL.on_device = 0;
    vR = vL;
    vR.kps_xy = xyR.data(); vR.octaves = octR.data(); vR.desc = F.mDescriptors.ptr<uchar>() + (size_t)32 * nL; vR.n = nR;
    auto scan = [&](bool right, int first, int count, int k, int32_t* out) {
        const int n = right ? nR : nL, base = right ? nL : 0;
        std::vector<unsigned char>& skip = right ? skipR : skipL;
        skip.assign(n, 0);
        for (int i = 0; i < n; i++) skip[i] = Taken(F, i + base);
        const std::vector<float>& q = right ? qR : qL;
        const std::vector<int32_t>& lev = right ? levR : levL;
        const std::vector<unsigned char>& d = right ? dR : dL;
        if (orbb_search_area_topk(mpMatcher, right ? &vR : &vL, grid4, &q[4 * (size_t)first], &lev[2 * (size_t)first], &d[32 * (size_t)first], count,
                                  skip.data(), 256, k, out) != ORBB_OK)
            throw std::runtime_error(std::string("orbb_search_area_topk failed: ") + orbb_matcher_last_error(mpMatcher));
    };
    const int nqL = (int)levL.size() / 2, nqR = (int)levR.size() / 2;
    outL.assign((size_t)nqL * kTopK * 2, -1);
    outR.assign((size_t)nqR * kTopK * 2, -1);
    if (nqL) scan(false, 0, nqL, kTopK, outL.data());
    if (nqR) scan(true, 0, nqR, kTopK, outR.data());
    // the first two candidates of a list that are still free (base = offset of the eye's key points in mvpMapPoints); one query again when the
    // list is used up
    auto head = [&](bool right, int j, Cand* c, int& nc) {
        const int base = right ? nL : 0;
        const int32_t* list = right ? &outR[(size_t)j * kTopK * 2] : &outL[(size_t)j * kTopK * 2];
        nc = 0;
        int valid = 0;
        for (int t = 0; t < kTopK; t++) {
            const int idx = list[2 * t + 1];
            if (idx < 0) break;
            valid++;
            if (Taken(F, idx + base)) continue;
            if (nc < 2) { c[nc].dist = list[2 * t]; c[nc].idx = idx; nc++; }
        }
        if (nc < 2 && valid == kTopK) {
            int32_t o[4] = {256, -1, 256, -1};
            scan(right, j, 1, 2, o);
            nc = 0;
            for (int t = 0; t < 2; t++) if (o[2 * t + 1] >= 0) { c[nc].dist = o[2 * t]; c[nc].idx = o[2 * t + 1]; nc++; }
            mnRescans++;
        }
    };
    int nmatches = 0;
    for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {
        MapPoint* pMP = vpMapPoints[iMP];
        if (qOfL[iMP] >= 0) {                                                     // :60-137
            Cand c[2];
            int nc = 0;
            head(false, qOfL[iMP], c, nc);
            if (nc > 0) {
                const int bestDist = c[0].dist, bestIdx = c[0].idx, bestDist2 = nc > 1 ? c[1].dist : 256;
                const int bestLevel = F.mvKeys[bestIdx].octave, bestLevel2 = nc > 1 ? F.mvKeys[c[1].idx].octave : -1;
                if (bestDist <= TH_HIGH) {
                    if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;      // (skips the right half of this point too)
                    if (bestLevel != bestLevel2 || bestDist <= nnratio * bestDist2) {
                        F.mvpMapPoints[bestIdx] = pMP;
                        if (F.mvLeftToRightMatch[bestIdx] != -1) {
                            F.mvpMapPoints[F.mvLeftToRightMatch[bestIdx] + nL] = pMP;
                            nmatches++;
                        }
                        nmatches++;
                    }
                }
            }
        }
        if (qOfR[iMP] >= 0) {                                                     // :139-208
            Cand c[2];
            int nc = 0;
            head(true, qOfR[iMP], c, nc);
            if (nc == 0) continue;
            const int bestDist = c[0].dist, bestIdx = c[0].idx, bestDist2 = nc > 1 ? c[1].dist : 256;
            const int bestLevel = F.mvKeysRight[bestIdx].octave, bestLevel2 = nc > 1 ? F.mvKeysRight[c[1].idx].octave : -1;
            if (bestDist <= TH_HIGH) {
                if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
                if (F.mvRightToLeftMatch[bestIdx] != -1) {
                    F.mvpMapPoints[F.mvRightToLeftMatch[bestIdx]] = pMP;
                    nmatches++;
                }
                F.mvpMapPoints[bestIdx + nL] = pMP;
                nmatches++;
            }
        }
    }
    return nmatches;
}

int ORBmatcherGPU::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th, const bool bFarPoints,
                                      const float thFarPoints, const float nnratio) {
    if (F.Nleft != -1) return SearchByProjectionFisheye(F, vpMapPoints, th, bFarPoints, thFarPoints, nnratio);
    Impl& s = Scratch();
    const bool bFactor = th != 1.0;
    s.q.clear(); s.qlev.clear(); s.qdesc.clear(); s.src.clear();
    for (size_t iMP = 0; iMP < vpMapPoints.size(); iMP++) {                       // :49-75
        MapPoint* pMP = vpMapPoints[iMP];
        if (!pMP->mbTrackInView && !pMP->mbTrackInViewR) continue;
        if (bFarPoints && pMP->mTrackDepth > thFarPoints) continue;
        if (pMP->isBad()) continue;
        if (!pMP->mbTrackInView) continue;
        const int nPredictedLevel = pMP->mnTrackScaleLevel;
        float r = RadiusByViewingCos(pMP->mTrackViewCos);
        if (bFactor) r *= th;
        const float q4[4] = {pMP->mTrackProjX, pMP->mTrackProjY, r * F.mvScaleFactors[nPredictedLevel], pMP->mTrackProjXR};
        s.q.insert(s.q.end(), q4, q4 + 4);
        s.qlev.push_back(nPredictedLevel - 1); s.qlev.push_back(nPredictedLevel);
        const cv::Mat d = pMP->GetDescriptor();
        s.qdesc.insert(s.qdesc.end(), d.ptr<uchar>(), d.ptr<uchar>() + 32);
        s.src.push_back((int)iMP);
    }
    const int nq = (int)s.src.size();
    RunScan(F, nq, kTopK, s.out);
    int nmatches = 0;
    for (int j = 0; j < nq; j++) {                                                // :77-141, in list order
        Cand c[2];
        int nc = 0;
        if (!LiveHead(F, &s.out[(size_t)j * kTopK * 2], kTopK, 2, c, nc)) Rescan(F, j, 2, c, nc);
        if (nc == 0) continue;
        const int bestDist = c[0].dist, bestIdx = c[0].idx, bestDist2 = nc > 1 ? c[1].dist : 256;
        const int bestLevel = F.mvKeysUn[bestIdx].octave, bestLevel2 = nc > 1 ? F.mvKeysUn[c[1].idx].octave : -1;
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            if (bestLevel != bestLevel2 || bestDist <= nnratio * bestDist2) {
                F.mvpMapPoints[bestIdx] = vpMapPoints[s.src[j]];
                nmatches++;
            }
        }
    }
    return nmatches;
}

// The motion-model search on fisheye-stereo frames (Nleft != -1, ORBmatcher.cc:1676-1887 with the right-eye half :1798-1860): every map point
// of the last frame is projected into the left eye and -- through GetRelativePoseTrl(), with the LEFT camera model as the reference does --
// into the right eye.  Three batched scans: the left set with the taken key points masked (four candidates), the left set unmasked with
// k = 1 (the reference leaves a point -- right half included -- when its left window holds no key point at all, :1736, which is not the same
// as "every candidate is taken"), and the right set masked.  Then the reference's loop in order.
int ORBmatcherGPU::SearchByProjectionFisheye(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono, const bool checkOrientation) {
    Impl& s = Scratch();
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
    const float factor = 1.0f / HISTO_LENGTH;
    const Sophus::SE3f Tcw = CurrentFrame.GetPose();
    const Eigen::Vector3f twc = Tcw.inverse().translation();
    const Sophus::SE3f Tlw = LastFrame.GetPose();
    const Eigen::Vector3f tlc = Tlw * twc;
    const bool bForward = tlc(2) > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc(2) > CurrentFrame.mb && !bMono;
    const int nL = CurrentFrame.Nleft, nR = (int)CurrentFrame.mvKeysRight.size();
    if (!CurrentFrame.mDescriptors.isContinuous()) throw std::runtime_error("Frame::mDescriptors must be continuous");
    const float grid4[4] = {Frame::mnMinX, Frame::mnMinY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv};
    std::vector<float> qL, qR;
    std::vector<int32_t> lev, outL, outAny, outR;
    std::vector<unsigned char> dq, skipL, skipR;
    s.src.clear();
    for (int i = 0; i < LastFrame.N; i++) {                                       // :1695-1727 and :1799-1813
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        Eigen::Vector3f x3Dw = pMP->GetWorldPos();
        Eigen::Vector3f x3Dc = Tcw * x3Dw;
        const float invzc = 1.0 / x3Dc(2);
        if (invzc < 0) continue;
        Eigen::Vector2f uv = CurrentFrame.mpCamera->project(x3Dc);
        if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
        if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
        const int nLastOctave = (LastFrame.Nleft == -1 || i < LastFrame.Nleft) ? LastFrame.mvKeys[i].octave : LastFrame.mvKeysRight[i - LastFrame.Nleft].octave;
        const float radius = th * CurrentFrame.mvScaleFactors[nLastOctave];
        Eigen::Vector3f x3Dr = CurrentFrame.GetRelativePoseTrl() * x3Dc;
        Eigen::Vector2f uvr = CurrentFrame.mpCamera->project(x3Dr);
        const float a4[4] = {uv(0), uv(1), radius, -1.0f}, b4[4] = {uvr(0), uvr(1), radius, -1.0f};
        qL.insert(qL.end(), a4, a4 + 4);
        qR.insert(qR.end(), b4, b4 + 4);
        if (bForward) { lev.push_back(nLastOctave); lev.push_back(-1); }
        else if (bBackward) { lev.push_back(0); lev.push_back(nLastOctave); }
        else { lev.push_back(nLastOctave - 1); lev.push_back(nLastOctave + 1); }
        const cv::Mat d = pMP->GetDescriptor();
        dq.insert(dq.end(), d.ptr<uchar>(), d.ptr<uchar>() + 32);
        s.src.push_back(i);
    }
    const int nq = (int)s.src.size();
    std::vector<float> xyL((size_t)nL * 2), xyR((size_t)nR * 2);
    std::vector<int32_t> octL(nL), octR(nR);
    for (int i = 0; i < nL; i++) { xyL[2 * i] = CurrentFrame.mvKeys[i].pt.x; xyL[2 * i + 1] = CurrentFrame.mvKeys[i].pt.y; octL[i] = CurrentFrame.mvKeys[i].octave; }
    for (int i = 0; i < nR; i++) {
        xyR[2 * i] = CurrentFrame.mvKeysRight[i].pt.x; xyR[2 * i + 1] = CurrentFrame.mvKeysRight[i].pt.y; octR[i] = CurrentFrame.mvKeysRight[i].octave;
    }
    orbb_frame_view vL, vR;
    vL.kps_xy = xyL.data(); vL.kps_stride = 8; vL.octaves = octL.data(); vL.oct_stride = 4; vL.desc = CurrentFrame.mDescriptors.ptr<uchar>();
    vL.u_right = nullptr; vL.n = nL; vL.on_device = 0;
    vR = vL;
    vR.kps_xy = xyR.data(); vR.octaves = octR.data(); vR.desc = CurrentFrame.mDescriptors.ptr<uchar>() + (size_t)32 * nL; vR.n = nR;
    auto scan = [&](bool right, bool masked, int first, int count, int k, int init, int32_t* out) {
        const int n = right ? nR : nL, base = right ? nL : 0;
        std::vector<unsigned char>& skip = right ? skipR : skipL;
        skip.assign(n, 0);
        if (masked) for (int i = 0; i < n; i++) skip[i] = Taken(CurrentFrame, i + base);
        const std::vector<float>& q = right ? qR : qL;
        if (orbb_search_area_topk(mpMatcher, right ? &vR : &vL, grid4, &q[4 * (size_t)first], &lev[2 * (size_t)first], &dq[32 * (size_t)first], count,
                                  skip.data(), init, k, out) != ORBB_OK)
            throw std::runtime_error(std::string("orbb_search_area_topk failed: ") + orbb_matcher_last_error(mpMatcher));
    };
    outL.assign((size_t)nq * kTopK * 2, -1);
    outAny.assign((size_t)nq * 2, -1);
    outR.assign((size_t)nq * kTopK * 2, -1);
    if (nq) {
        scan(false, true, 0, nq, kTopK, 256, outL.data());
        scan(false, false, 0, nq, 1, 257, outAny.data());
        scan(true, true, 0, nq, kTopK, 256, outR.data());
    }
    auto head = [&](bool right, int j, Cand& c) {                                // best candidate that is still free; false: none
        const int base = right ? nL : 0;
        const int32_t* list = right ? &outR[(size_t)j * kTopK * 2] : &outL[(size_t)j * kTopK * 2];
        int valid = 0;
        for (int t = 0; t < kTopK; t++) {
            const int idx = list[2 * t + 1];
            if (idx < 0) break;
            valid++;
            if (Taken(CurrentFrame, idx + base)) continue;
            c.dist = list[2 * t]; c.idx = idx;
            return true;
        }
        if (valid < kTopK) return false;
        int32_t o[2] = {256, -1};
        scan(right, true, j, 1, 1, 256, o);
        mnRescans++;
        if (o[1] < 0) return false;
        c.dist = o[0]; c.idx = o[1];
        return true;
    };
    int nmatches = 0;
    for (int j = 0; j < nq; j++) {
        const int i = s.src[j];
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (outAny[2 * (size_t)j + 1] < 0) continue;                              // :1735-1736 vIndices2.empty(): the right half is skipped too
        const float lastAngle = (LastFrame.Nleft == -1 || i < LastFrame.Nleft) ? LastFrame.mvKeys[i].angle : LastFrame.mvKeysRight[i - LastFrame.Nleft].angle;
        Cand c;
        if (head(false, j, c) && c.dist <= TH_HIGH) {                             // :1738-1796
            CurrentFrame.mvpMapPoints[c.idx] = pMP;
            nmatches++;
            if (checkOrientation) {
                float rot = lastAngle - CurrentFrame.mvKeys[c.idx].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(c.idx);
            }
        }
        if (head(true, j, c) && c.dist <= TH_HIGH) {                              // :1815-1859
            CurrentFrame.mvpMapPoints[c.idx + nL] = pMP;
            nmatches++;
            if (checkOrientation) {
                float rot = lastAngle - CurrentFrame.mvKeysRight[c.idx].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(c.idx + nL);
            }
        }
    }
    if (checkOrientation) {                                                       // :1866-1884
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i != ind1 && i != ind2 && i != ind3) {
                for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) {
                    CurrentFrame.mvpMapPoints[rotHist[i][j]] = static_cast<MapPoint*>(NULL);
                    nmatches--;
                }
            }
        }
    }
    return nmatches;
}

int ORBmatcherGPU::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono, const bool checkOrientation) {
    if (CurrentFrame.Nleft != -1) return SearchByProjectionFisheye(CurrentFrame, LastFrame, th, bMono, checkOrientation);
    if (LastFrame.Nleft != -1)
        throw std::logic_error("ORBmatcherGPU::SearchByProjection: a fisheye-stereo last frame with a monocular current frame is not a rig the reference builds");
    Impl& s = Scratch();
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    const Sophus::SE3f Tcw = CurrentFrame.GetPose();                              // :1686-1693
    const Eigen::Vector3f twc = Tcw.inverse().translation();
    const Sophus::SE3f Tlw = LastFrame.GetPose();
    const Eigen::Vector3f tlc = Tlw * twc;
    const bool bForward = tlc(2) > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc(2) > CurrentFrame.mb && !bMono;
    s.q.clear(); s.qlev.clear(); s.qdesc.clear(); s.src.clear();
    for (int i = 0; i < LastFrame.N; i++) {                                       // :1695-1735: project, window, level range
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        Eigen::Vector3f x3Dw = pMP->GetWorldPos();
        Eigen::Vector3f x3Dc = Tcw * x3Dw;
        const float invzc = 1.0 / x3Dc(2);
        if (invzc < 0) continue;
        Eigen::Vector2f uv = CurrentFrame.mpCamera->project(x3Dc);
        if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
        if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
        const int nLastOctave = LastFrame.mvKeys[i].octave;
        const float radius = th * CurrentFrame.mvScaleFactors[nLastOctave];
        const float q4[4] = {uv(0), uv(1), radius, uv(0) - CurrentFrame.mbf * invzc};      // (:1757: ur of the stereo check)
        s.q.insert(s.q.end(), q4, q4 + 4);
        if (bForward) { s.qlev.push_back(nLastOctave); s.qlev.push_back(-1); }
        else if (bBackward) { s.qlev.push_back(0); s.qlev.push_back(nLastOctave); }
        else { s.qlev.push_back(nLastOctave - 1); s.qlev.push_back(nLastOctave + 1); }
        const cv::Mat d = pMP->GetDescriptor();
        s.qdesc.insert(s.qdesc.end(), d.ptr<uchar>(), d.ptr<uchar>() + 32);
        s.src.push_back(i);
    }
    const int nq = (int)s.src.size();
    RunScan(CurrentFrame, nq, kTopK, s.out);
    int nmatches = 0;
    for (int j = 0; j < nq; j++) {                                                // :1737-1790: best candidate, TH_HIGH, rotation vote
        Cand c[1];
        int nc = 0;
        if (!LiveHead(CurrentFrame, &s.out[(size_t)j * kTopK * 2], kTopK, 1, c, nc)) Rescan(CurrentFrame, j, 1, c, nc);
        if (nc == 0) continue;
        const int bestDist = c[0].dist, bestIdx2 = c[0].idx, i = s.src[j];
        if (bestDist <= TH_HIGH) {
            CurrentFrame.mvpMapPoints[bestIdx2] = LastFrame.mvpMapPoints[i];
            nmatches++;
            if (checkOrientation) {
                float rot = LastFrame.mvKeysUn[i].angle - CurrentFrame.mvKeysUn[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOrientation) {                                                       // :1866-1884
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i != ind1 && i != ind2 && i != ind3) {
                for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) {
                    CurrentFrame.mvpMapPoints[rotHist[i][j]] = static_cast<MapPoint*>(NULL);
                    nmatches--;
                }
            }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12) (ORBmatcher.cc:765-905; LoopClosing's candidate check,
// LoopClosing.cc:1680): as above between two key frames -- a candidate must hold a good map point itself and must not have been matched
// earlier in the call (vbMatched2, :826), the accept test is bestDist1 < TH_LOW (strict).  The candidate lists are collected with the
// static part of that filter applied; one orbb_best2_csr launch against the second key frame's descriptors; lists whose head was matched
// meanwhile are walked again on the host.
int ORBmatcherGPU::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float nnratio, const bool checkOrientation) {
    if (pKF1->NLeft != -1 || pKF2->NLeft != -1)
        throw std::logic_error("ORBmatcherGPU::SearchByBoW: fisheye-stereo key frames keep the reference's host path");
    Impl& s = Scratch();
    const std::vector<cv::KeyPoint>& vKeysUn1 = pKF1->mvKeysUn;
    const std::vector<cv::KeyPoint>& vKeysUn2 = pKF2->mvKeysUn;
    const DBoW2::FeatureVector& vFeatVec1 = pKF1->mFeatVec;
    const DBoW2::FeatureVector& vFeatVec2 = pKF2->mFeatVec;
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    const cv::Mat& Descriptors1 = pKF1->mDescriptors;
    const cv::Mat& Descriptors2 = pKF2->mDescriptors;
    vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    std::vector<bool> vbMatched2(vpMapPoints2.size(), false);
    std::vector<unsigned char> good2(vpMapPoints2.size(), 0);                     // :824-830, the part that does not change during the call
    for (size_t i = 0; i < vpMapPoints2.size(); i++) good2[i] = vpMapPoints2[i] && !vpMapPoints2[i]->isBad();
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
    const float factor = 1.0f / HISTO_LENGTH;
    s.qdesc.clear(); s.src.clear(); s.cand.clear(); s.rowptr.assign(1, 0);
    DBoW2::FeatureVector::const_iterator f1it = vFeatVec1.begin(), f2it = vFeatVec2.begin();
    const DBoW2::FeatureVector::const_iterator f1end = vFeatVec1.end(), f2end = vFeatVec2.end();
    while (f1it != f1end && f2it != f2end) {                                      // :793-882
        if (f1it->first == f2it->first) {
            for (size_t i1 = 0, iend1 = f1it->second.size(); i1 < iend1; i1++) {
                const size_t idx1 = f1it->second[i1];
                MapPoint* pMP1 = vpMapPoints1[idx1];
                if (!pMP1 || pMP1->isBad()) continue;
                const uchar* d = Descriptors1.ptr<uchar>((int)idx1);
                s.qdesc.insert(s.qdesc.end(), d, d + 32);
                s.src.push_back((int)idx1);
                for (size_t i2 = 0, iend2 = f2it->second.size(); i2 < iend2; i2++)
                    if (good2[f2it->second[i2]]) s.cand.push_back((int32_t)f2it->second[i2]);
                s.rowptr.push_back((int32_t)s.cand.size());
            }
            f1it++;
            f2it++;
        } else if (f1it->first < f2it->first) {
            f1it = vFeatVec1.lower_bound(f2it->first);
        } else {
            f2it = vFeatVec2.lower_bound(f1it->first);
        }
    }
    const int nq = (int)s.src.size();
    s.out.assign((size_t)nq * 4, -1);
    if (nq > 0 && !s.cand.empty()) {
        if (!Descriptors2.isContinuous()) throw std::runtime_error("KeyFrame::mDescriptors must be continuous");
        if (orbb_best2_csr(mpMatcher, s.qdesc.data(), nq, Descriptors2.ptr<uchar>(), Descriptors2.rows, s.cand.data(), s.rowptr.data(), 256, s.out.data()) != ORBB_OK)
            throw std::runtime_error(std::string("orbb_best2_csr failed: ") + orbb_matcher_last_error(mpMatcher));
    } else {
        for (int j = 0; j < nq; j++) { s.out[4 * (size_t)j] = 256; s.out[4 * (size_t)j + 2] = 256; }
    }
    int nmatches = 0;
    for (int j = 0; j < nq; j++) {
        const int idx1 = s.src[j];
        int bestDist1 = s.out[4 * (size_t)j], bestIdx2 = s.out[4 * (size_t)j + 1], bestDist2 = s.out[4 * (size_t)j + 2];
        const int secondIdx = s.out[4 * (size_t)j + 3];
        if ((bestIdx2 >= 0 && vbMatched2[bestIdx2]) || (secondIdx >= 0 && vbMatched2[secondIdx])) {
            bestDist1 = 256; bestIdx2 = -1; bestDist2 = 256;                      // :816-846 on the live vbMatched2
            const uchar* d1 = s.qdesc.data() + (size_t)32 * j;
            for (int c = s.rowptr[j]; c < s.rowptr[j + 1]; c++) {
                const int idx2 = s.cand[c];
                if (vbMatched2[idx2]) continue;
                const int dist = orbb_hamming_distance(d1, Descriptors2.ptr<uchar>(idx2));
                if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                else if (dist < bestDist2) bestDist2 = dist;
            }
            mnRescans++;
        }
        if (bestDist1 < TH_LOW) {                                                 // :848-868
            if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                vpMatches12[idx1] = vpMapPoints2[bestIdx2];
                vbMatched2[bestIdx2] = true;
                if (checkOrientation) {
                    float rot = vKeysUn1[idx1].angle - vKeysUn2[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(idx1);
                }
                nmatches++;
            }
        }
    }
    if (checkOrientation) {                                                       // :884-902
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) {
                vpMatches12[rotHist[i][j]] = static_cast<MapPoint*>(NULL);
                nmatches--;
            }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchForInitialization (ORBmatcher.cc:648-766; Tracking::MonocularInitialization, Tracking.cc:2527, windowSize 100): every
// level-0 key point of F1 looks for its best / second-best descriptor among the level-0 key points of F2 inside a square window around
// vbPrevMatched[i1].  Unlike the projection searches the scan of the reference is NOT masked by what is taken: a candidate i2 is left
// out only when an earlier key point holds it with a distance <= this one (vMatchedDistance, :687), and a later, closer key point takes
// it over (:706-710).  The batched scan returns the four best candidates of every key point in scan order; the loop below applies the
// distance test on the live vMatchedDistance, and only a key point that loses more than two of its four candidates that way walks its
// window on the host, exactly as the reference does.
int ORBmatcherGPU::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                           int windowSize, const float nnratio, const bool checkOrientation) {
    if (F1.Nleft != -1 || F2.Nleft != -1)
        throw std::logic_error("ORBmatcherGPU::SearchForInitialization: fisheye-stereo frames (Nleft != -1) keep the reference's host path");
    Impl& s = Scratch();
    const int n1 = (int)F1.mvKeysUn.size(), n2 = (int)F2.mvKeysUn.size();
    const int INF = 0x7fffffff;
    int nmatches = 0;
    vnMatches12 = std::vector<int>(n1, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<int> vMatchedDistance(n2, INF), vnMatches21(n2, -1);
    s.q.clear(); s.qlev.clear(); s.qdesc.clear(); s.src.clear();
    for (int i1 = 0; i1 < n1; i1++) {                                             // :661-668
        if (F1.mvKeysUn[i1].octave > 0) continue;
        const float q4[4] = {vbPrevMatched[i1].x, vbPrevMatched[i1].y, (float)windowSize, -1.0f};
        s.q.insert(s.q.end(), q4, q4 + 4);
        s.qlev.push_back(0); s.qlev.push_back(0);                                 // GetFeaturesInArea(.., level1, level1) with level1 == 0
        const uchar* d = F1.mDescriptors.ptr<uchar>(i1);
        s.qdesc.insert(s.qdesc.end(), d, d + 32);
        s.src.push_back(i1);
    }
    const int nq = (int)s.src.size();
    RunScan(F2, nq, kTopK, s.out, false, false, 257);                             // (bestDist starts at INT_MAX: every distance 0..256 counts)
    for (int j = 0; j < nq; j++) {
        const int i1 = s.src[j];
        const int32_t* list = &s.out[(size_t)j * kTopK * 2];
        int bestDist = INF, bestDist2 = INF, bestIdx2 = -1, valid = 0, got = 0;
        for (int t = 0; t < kTopK && got < 2; t++) {                              // :679-700 on the head of the list
            const int idx = list[2 * t + 1];
            if (idx < 0) break;
            valid++;
            const int dist = list[2 * t];
            if (vMatchedDistance[idx] <= dist) continue;
            if (got == 0) { bestDist = dist; bestIdx2 = idx; } else bestDist2 = dist;
            got++;
        }
        if (got < 2 && valid == kTopK) {                                          // more candidates may follow the four: the reference's loop for this one
            const std::vector<size_t> vIndices2 = F2.GetFeaturesInArea(vbPrevMatched[i1].x, vbPrevMatched[i1].y, windowSize, 0, 0);
            bestDist = INF; bestDist2 = INF; bestIdx2 = -1;
            const uchar* d1 = F1.mDescriptors.ptr<uchar>(i1);
            for (size_t v = 0; v < vIndices2.size(); v++) {
                const size_t i2 = vIndices2[v];
                const int dist = orbb_hamming_distance(d1, F2.mDescriptors.ptr<uchar>((int)i2));
                if (vMatchedDistance[i2] <= dist) continue;
                if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = (int)i2; }
                else if (dist < bestDist2) bestDist2 = dist;
            }
            mnRescans++;
        }
        if (bestDist <= TH_LOW) {                                                 // :702-728
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) {
                    vnMatches12[vnMatches21[bestIdx2]] = -1;
                    nmatches--;
                }
                vnMatches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOrientation) {
                    float rot = F1.mvKeysUn[i1].angle - F2.mvKeysUn[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(i1);
                }
            }
        }
    }
    if (checkOrientation) {                                                       // :732-755
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) {
                const int idx1 = rotHist[i][j];
                if (vnMatches12[idx1] >= 0) {
                    vnMatches12[idx1] = -1;
                    nmatches--;
                }
            }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)                                               // :757-760
        if (vnMatches12[i1] >= 0) vbPrevMatched[i1] = F2.mvKeysUn[vnMatches12[i1]].pt;
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1889-2010;
// Tracking::Relocalization, Tracking.cc:3765 / :3779): the key frame's map points that were not found yet are projected with the
// frame's pose, the scale level predicted from the distance, and the best descriptor inside the window wins if it is within ORBdist.
// Same shape as the motion-model search; the scan leaves out every key point that holds a map point (:1952) and has no mvuRight test.
int ORBmatcherGPU::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist,
                                      const bool checkOrientation) {
    if (CurrentFrame.Nleft != -1)
        throw std::logic_error("ORBmatcherGPU::SearchByProjection: fisheye-stereo frames (Nleft != -1) keep the reference's host path");
    Impl& s = Scratch();
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
    const float factor = 1.0f / HISTO_LENGTH;
    const Sophus::SE3f Tcw = CurrentFrame.GetPose();                              // :1893-1894
    Eigen::Vector3f Ow = Tcw.inverse().translation();
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    s.q.clear(); s.qlev.clear(); s.qdesc.clear(); s.src.clear();
    for (size_t i = 0, iend = vpMPs.size(); i < iend; i++) {                      // :1904-1944: project, distance range, predicted level, window
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        Eigen::Vector3f x3Dw = pMP->GetWorldPos();
        Eigen::Vector3f x3Dc = Tcw * x3Dw;
        const Eigen::Vector2f uv = CurrentFrame.mpCamera->project(x3Dc);
        if (uv(0) < CurrentFrame.mnMinX || uv(0) > CurrentFrame.mnMaxX) continue;
        if (uv(1) < CurrentFrame.mnMinY || uv(1) > CurrentFrame.mnMaxY) continue;
        Eigen::Vector3f PO = x3Dw - Ow;
        float dist3D = PO.norm();
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        int nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame);
        const float radius = th * CurrentFrame.mvScaleFactors[nPredictedLevel];
        const float q4[4] = {uv(0), uv(1), radius, -1.0f};
        s.q.insert(s.q.end(), q4, q4 + 4);
        s.qlev.push_back(nPredictedLevel - 1); s.qlev.push_back(nPredictedLevel + 1);
        const cv::Mat d = pMP->GetDescriptor();
        s.qdesc.insert(s.qdesc.end(), d.ptr<uchar>(), d.ptr<uchar>() + 32);
        s.src.push_back((int)i);
    }
    const int nq = (int)s.src.size();
    RunScan(CurrentFrame, nq, kTopK, s.out, true, false, 256, true);
    int nmatches = 0;
    for (int j = 0; j < nq; j++) {                                                // :1946-1982
        Cand c[1];
        int nc = 0;
        if (!LiveHead(CurrentFrame, &s.out[(size_t)j * kTopK * 2], kTopK, 1, c, nc, true)) Rescan(CurrentFrame, j, 1, c, nc, false, true);
        if (nc == 0) continue;
        const int bestDist = c[0].dist, bestIdx2 = c[0].idx, i = s.src[j];
        if (bestDist <= ORBdist) {
            CurrentFrame.mvpMapPoints[bestIdx2] = vpMPs[i];
            nmatches++;
            if (checkOrientation) {
                float rot = pKF->mvKeysUn[i].angle - CurrentFrame.mvKeysUn[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (checkOrientation) {                                                       // :1988-2007
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i != ind1 && i != ind2 && i != ind3) {
                for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) {
                    CurrentFrame.mvpMapPoints[rotHist[i][j]] = NULL;
                    nmatches--;
                }
            }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, vpPoints, vpMatched, th, ratioHamming) (ORBmatcher.cc:427-530;
// LoopClosing.cc:1795 / :1982) and its sibling with vpPointsKFs / vpMatchedKF (:532-646, LoopClosing.cc:1773): map points of a loop / merge candidate projected into a key frame with a Sim3.  The scan side is a key
// frame here: its undistorted key points, octaves and descriptors go up with the call (key frames are not cached: a loop candidate is
// scanned once or twice); KeyFrame::GetFeaturesInArea (KeyFrame.cc:707-751) has the grid walk of Frame::GetFeaturesInArea without the level
// test, which the matcher applies itself (:508-511) -- the same filter as a level window [n - 1, n] of orbb_search_area_topk.  Key points
// that are matched already (vpMatched non-null, :505) are masked, on the live vector.
int ORBmatcherGPU::SearchByProjectionSim3(KeyFrame* pKF, const float* R9, const float* t3, float scale, const std::vector<MapPoint*>& vpPoints,
                                          std::vector<MapPoint*>& vpMatched, int th, float ratioHamming, const std::vector<KeyFrame*>* vpPointsKFs,
                                          std::vector<KeyFrame*>* vpMatchedKF) {
    if (pKF->NLeft != -1) throw std::logic_error("ORBmatcherGPU::SearchByProjection: fisheye-stereo key frames keep the reference's host path");
    Impl& s = Scratch();
    Eigen::Matrix3f Rm;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rm(i, j) = R9[3 * i + j];
    Eigen::Vector3f tv(t3[0], t3[1], t3[2]);
    Sophus::SE3f Tcw = Sophus::SE3f(Rm, tv / scale);                              // :435-436: SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale())
    Eigen::Vector3f Ow = Tcw.inverse().translation();
    std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
    s.q.clear(); s.qlev.clear(); s.qdesc.clear(); s.src.clear();
    for (int iMP = 0, iendMP = (int)vpPoints.size(); iMP < iendMP; iMP++) {       // :445-490: project, range, viewing angle, predicted level, window
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        Eigen::Vector3f p3Dw = pMP->GetWorldPos();
        Eigen::Vector3f p3Dc = Tcw * p3Dw;
        if (p3Dc(2) < 0.0) continue;
        float u, v;
        if (!vpPointsKFs) {                                                       // :463 the camera model's projection
            const Eigen::Vector2f uv = pKF->mpCamera->project(p3Dc);
            u = uv(0); v = uv(1);
        } else {                                                                  // :573-578 the second overload projects by hand
            const float invz = 1 / p3Dc(2);
            const float x = p3Dc(0) * invz;
            const float y = p3Dc(1) * invz;
            u = pKF->fx * x + pKF->cx;
            v = pKF->fy * y + pKF->cy;
        }
        if (!pKF->IsInImage(u, v)) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        Eigen::Vector3f PO = p3Dw - Ow;
        const float dist = PO.norm();
        if (dist < minDistance || dist > maxDistance) continue;
        Eigen::Vector3f Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist) continue;
        int nPredictedLevel = pMP->PredictScale(dist, pKF);
        const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
        const float q4[4] = {u, v, radius, -1.0f};
        s.q.insert(s.q.end(), q4, q4 + 4);
        s.qlev.push_back(nPredictedLevel - 1); s.qlev.push_back(nPredictedLevel);
        const cv::Mat d = pMP->GetDescriptor();
        s.qdesc.insert(s.qdesc.end(), d.ptr<uchar>(), d.ptr<uchar>() + 32);
        s.src.push_back(iMP);
    }
    const int nq = (int)s.src.size(), n = (int)pKF->mvKeysUn.size();
    if (nq == 0) return 0;
    if (!pKF->mDescriptors.isContinuous()) throw std::runtime_error("KeyFrame::mDescriptors must be continuous");
    s.xy.resize((size_t)n * 2);
    s.oct.resize(n);
    for (int i = 0; i < n; i++) { s.xy[2 * i] = pKF->mvKeysUn[i].pt.x; s.xy[2 * i + 1] = pKF->mvKeysUn[i].pt.y; s.oct[i] = pKF->mvKeysUn[i].octave; }
    orbb_frame_view view;
    view.kps_xy = s.xy.data(); view.kps_stride = 8; view.octaves = s.oct.data(); view.oct_stride = 4;
    view.desc = pKF->mDescriptors.ptr<uchar>(); view.u_right = nullptr; view.n = n; view.on_device = 0;
    const float grid4[4] = {(float)pKF->mnMinX, (float)pKF->mnMinY, pKF->mfGridElementWidthInv, pKF->mfGridElementHeightInv};
    auto scan = [&](int first, int count, int k, int32_t* out) {
        s.skip.assign(n, 0);
        for (int i = 0; i < n; i++) s.skip[i] = vpMatched[i] != NULL;
        if (orbb_search_area_topk(mpMatcher, &view, grid4, &s.q[4 * (size_t)first], &s.qlev[2 * (size_t)first], &s.qdesc[32 * (size_t)first], count,
                                  s.skip.data(), 256, k, out) != ORBB_OK)
            throw std::runtime_error(std::string("orbb_search_area_topk failed: ") + orbb_matcher_last_error(mpMatcher));
    };
    s.out.assign((size_t)nq * kTopK * 2, -1);
    scan(0, nq, kTopK, s.out.data());
    int nmatches = 0;
    for (int j = 0; j < nq; j++) {                                                // :499-526, in the order of vpPoints
        const int32_t* list = &s.out[(size_t)j * kTopK * 2];
        int bestDist = 256, bestIdx = -1, valid = 0;
        for (int t = 0; t < kTopK; t++) {
            const int idx = list[2 * t + 1];
            if (idx < 0) break;
            valid++;
            if (vpMatched[idx]) continue;
            bestDist = list[2 * t]; bestIdx = idx;
            break;
        }
        if (bestIdx < 0 && valid == kTopK) {                                      // all four were matched meanwhile: this point again, with the live mask
            int32_t o[2] = {256, -1};
            scan(j, 1, 1, o);
            bestDist = o[0]; bestIdx = o[1];
            mnRescans++;
        }
        if (bestIdx >= 0 && bestDist <= TH_LOW * ratioHamming) {
            vpMatched[bestIdx] = vpPoints[s.src[j]];
            if (vpMatchedKF) (*vpMatchedKF)[bestIdx] = (*vpPointsKFs)[s.src[j]];
            nmatches++;
        }
    }
    return nmatches;
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (ORBmatcher.cc:906-1146; LocalMapping::CreateNewMapPoints,
// LocalMapping.cc:466) for two monocular / rectified-stereo key frames.  Key points without map points that share a vocabulary node are
// compared; a candidate counts only if it is not too close to the epipole (:1019-1027) and satisfies the epipolar constraint of the
// camera model (:1067), and the reference's update rule `dist > bestDist -> continue` (:1010) lets a LATER candidate with the same distance
// replace the current best.  vbMatched2 is never written in the reference, so the scans of all key points are independent: the geometric
// tests -- pure functions of the two key points -- run on the host while the candidate lists are collected, in REVERSED order, and one
// orbb_best2_csr launch with the bound TH_LOW + 1 returns the first minimum of every reversed list = the last minimum of the original one.
int ORBmatcherGPU::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo,
                                          const bool bCoarse, const bool checkOrientation) {
    if (pKF1->NLeft != -1 || pKF2->NLeft != -1 || pKF1->mpCamera2 || pKF2->mpCamera2)
        throw std::logic_error("ORBmatcherGPU::SearchForTriangulation: fisheye-stereo key frames keep the reference's host path");
    Impl& s = Scratch();
    const DBoW2::FeatureVector& vFeatVec1 = pKF1->mFeatVec;
    const DBoW2::FeatureVector& vFeatVec2 = pKF2->mFeatVec;
    Sophus::SE3f T1w = pKF1->GetPose();                                           // :913-930
    Sophus::SE3f T2w = pKF2->GetPose();
    Sophus::SE3f Tw2 = pKF2->GetPoseInverse();
    Eigen::Vector3f Cw = pKF1->GetCameraCenter();
    Eigen::Vector3f C2 = T2w * Cw;
    Eigen::Vector2f ep = pKF2->mpCamera->project(C2);
    Sophus::SE3f T12 = T1w * Tw2;
    Eigen::Matrix3f R12 = T12.rotationMatrix();
    Eigen::Vector3f t12 = T12.translation();
    GeometricCamera* pCamera1 = pKF1->mpCamera, *pCamera2 = pKF2->mpCamera;
    int nmatches = 0;
    std::vector<int> vMatches12(pKF1->N, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
    const float factor = 1.0f / HISTO_LENGTH;
    std::vector<unsigned char> free2(pKF2->N, 0);                                 // :1000-1008, the part that does not depend on the first key point
    for (int i = 0; i < pKF2->N; i++) free2[i] = !pKF2->GetMapPoint(i) && (!bOnlyStereo || pKF2->mvuRight[i] >= 0);
    s.qdesc.clear(); s.src.clear(); s.cand.clear(); s.rowptr.assign(1, 0);
    DBoW2::FeatureVector::const_iterator f1it = vFeatVec1.begin(), f2it = vFeatVec2.begin();
    const DBoW2::FeatureVector::const_iterator f1end = vFeatVec1.end(), f2end = vFeatVec2.end();
    while (f1it != f1end && f2it != f2end) {
        if (f1it->first == f2it->first) {
            for (size_t i1 = 0, iend1 = f1it->second.size(); i1 < iend1; i1++) {
                const size_t idx1 = f1it->second[i1];
                if (pKF1->GetMapPoint(idx1)) continue;
                const bool bStereo1 = pKF1->mvuRight[idx1] >= 0;
                if (bOnlyStereo && !bStereo1) continue;
                const cv::KeyPoint& kp1 = pKF1->mvKeysUn[idx1];
                const size_t first = s.cand.size();
                for (size_t i2 = f2it->second.size(); i2-- > 0;) {                 // reversed (see above)
                    const size_t idx2 = f2it->second[i2];
                    if (!free2[idx2]) continue;
                    const bool bStereo2 = pKF2->mvuRight[idx2] >= 0;
                    const cv::KeyPoint& kp2 = pKF2->mvKeysUn[idx2];
                    if (!bStereo1 && !bStereo2) {                                 // :1019-1027
                        const float distex = ep(0) - kp2.pt.x;
                        const float distey = ep(1) - kp2.pt.y;
                        if (distex * distex + distey * distey < 100 * pKF2->mvScaleFactors[kp2.octave]) continue;
                    }
                    if (bCoarse || pCamera1->epipolarConstrain(pCamera2, kp1, kp2, R12, t12, pKF1->mvLevelSigma2[kp1.octave], pKF2->mvLevelSigma2[kp2.octave]))
                        s.cand.push_back((int32_t)idx2);
                }
                if (s.cand.size() == first) continue;                             // (no candidate can match: no query)
                const uchar* d = pKF1->mDescriptors.ptr<uchar>((int)idx1);
                s.qdesc.insert(s.qdesc.end(), d, d + 32);
                s.src.push_back((int)idx1);
                s.rowptr.push_back((int32_t)s.cand.size());
            }
            f1it++;
            f2it++;
        } else if (f1it->first < f2it->first) {
            f1it = vFeatVec1.lower_bound(f2it->first);
        } else {
            f2it = vFeatVec2.lower_bound(f1it->first);
        }
    }
    const int nq = (int)s.src.size();
    s.out.assign((size_t)nq * 4, -1);
    if (nq > 0) {
        if (!pKF2->mDescriptors.isContinuous()) throw std::runtime_error("KeyFrame::mDescriptors must be continuous");
        if (orbb_best2_csr(mpMatcher, s.qdesc.data(), nq, pKF2->mDescriptors.ptr<uchar>(), pKF2->mDescriptors.rows, s.cand.data(), s.rowptr.data(), TH_LOW + 1,
                           s.out.data()) != ORBB_OK)
            throw std::runtime_error(std::string("orbb_best2_csr failed: ") + orbb_matcher_last_error(mpMatcher));
    }
    for (int j = 0; j < nq; j++) {                                                // :1077-1096
        const int bestIdx2 = s.out[4 * (size_t)j + 1], idx1 = s.src[j];
        if (bestIdx2 < 0) continue;
        vMatches12[idx1] = bestIdx2;
        nmatches++;
        if (checkOrientation) {
            float rot = pKF1->mvKeysUn[idx1].angle - pKF2->mvKeysUn[bestIdx2].angle;
            if (rot < 0.0) rot += 360.0f;
            int bin = round(rot * factor);
            if (bin == HISTO_LENGTH) bin = 0;
            rotHist[bin].push_back(idx1);
        }
    }
    if (checkOrientation) {                                                       // :1111-1130
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) {
                vMatches12[rotHist[i][j]] = -1;
                nmatches--;
            }
        }
    }
    vMatchedPairs.clear();
    vMatchedPairs.reserve(nmatches);
    for (size_t i = 0, iend = vMatches12.size(); i < iend; i++) {
        if (vMatches12[i] < 0) continue;
        vMatchedPairs.push_back(std::make_pair(i, (size_t)vMatches12[i]));
    }
    return nmatches;
}

// ORBmatcher::Fuse(KeyFrame* pKF, vpMapPoints, th, bRight = false) (ORBmatcher.cc:1148-1338; LocalMapping::SearchInNeighbors, LocalMapping.cc:772 /
// :802) for monocular / rectified-stereo key frames.  Every candidate inside the window and the level band must also pass a chi-square test
// of its reprojection error (:1272-1296: 5.99 on (ex, ey), 7.8 with the right coordinate where the key point has one) -- a test on the key
// point's position, not on its descriptor.  The scan is unmasked, so ONE launch returns the eight nearest candidates of every point and
// the loop below takes the first of them that passes the test: what the reference's scan keeps.  A point whose eight all fail walks its
// window on the host.  Which points are bad or already observed by the key frame changes while the loop replaces and adds (:1187-1196 are
// evaluated per iteration in the reference): the same tests run again at decision time on the live objects.
int ORBmatcherGPU::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th) {
    if (pKF->NLeft != -1) throw std::logic_error("ORBmatcherGPU::Fuse: fisheye-stereo key frames keep the reference's host path");
    Impl& s = Scratch();
    const int kFuseK = 8;
    GeometricCamera* pCamera = pKF->mpCamera;
    Sophus::SE3f Tcw = pKF->GetPose();
    Eigen::Vector3f Ow = pKF->GetCameraCenter();
    const float& bf = pKF->mbf;
    const int nMPs = (int)vpMapPoints.size();
    std::vector<float> urs;                                                       // per query: the projected right coordinate (:1219)
    s.q.clear(); s.qlev.clear(); s.qdesc.clear(); s.src.clear();
    for (int i = 0; i < nMPs; i++) {                                              // :1177-1252
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP) continue;
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;                     // (both can only stay true or turn true during the call)
        Eigen::Vector3f p3Dw = pMP->GetWorldPos();
        Eigen::Vector3f p3Dc = Tcw * p3Dw;
        if (p3Dc(2) < 0.0f) continue;
        const float invz = 1 / p3Dc(2);
        const Eigen::Vector2f uv = pCamera->project(p3Dc);
        if (!pKF->IsInImage(uv(0), uv(1))) continue;
        const float ur = uv(0) - bf * invz;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        Eigen::Vector3f PO = p3Dw - Ow;
        const float dist3D = PO.norm();
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        Eigen::Vector3f Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist3D) continue;
        int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
        const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
        const float q4[4] = {uv(0), uv(1), radius, -1.0f};
        s.q.insert(s.q.end(), q4, q4 + 4);
        s.qlev.push_back(nPredictedLevel - 1); s.qlev.push_back(nPredictedLevel);
        const cv::Mat d = pMP->GetDescriptor();
        s.qdesc.insert(s.qdesc.end(), d.ptr<uchar>(), d.ptr<uchar>() + 32);
        s.src.push_back(i);
        urs.push_back(ur);
    }
    const int nq = (int)s.src.size(), n = (int)pKF->mvKeysUn.size();
    if (nq == 0) return 0;
    if (!pKF->mDescriptors.isContinuous()) throw std::runtime_error("KeyFrame::mDescriptors must be continuous");
    s.xy.resize((size_t)n * 2);
    s.oct.resize(n);
    for (int i = 0; i < n; i++) { s.xy[2 * i] = pKF->mvKeysUn[i].pt.x; s.xy[2 * i + 1] = pKF->mvKeysUn[i].pt.y; s.oct[i] = pKF->mvKeysUn[i].octave; }
    orbb_frame_view view;
    view.kps_xy = s.xy.data(); view.kps_stride = 8; view.octaves = s.oct.data(); view.oct_stride = 4;
    view.desc = pKF->mDescriptors.ptr<uchar>(); view.u_right = nullptr; view.n = n; view.on_device = 0;
    const float grid4[4] = {(float)pKF->mnMinX, (float)pKF->mnMinY, pKF->mfGridElementWidthInv, pKF->mfGridElementHeightInv};
    s.out.assign((size_t)nq * kFuseK * 2, -1);
    if (orbb_search_area_topk(mpMatcher, &view, grid4, s.q.data(), s.qlev.data(), s.qdesc.data(), nq, nullptr, 256, kFuseK, s.out.data()) != ORBB_OK)
        throw std::runtime_error(std::string("orbb_search_area_topk failed: ") + orbb_matcher_last_error(mpMatcher));
    auto passes = [&](int idx, float u, float v, float ur) {                     // :1272-1296
        const cv::KeyPoint& kp = pKF->mvKeysUn[idx];
        const int& kpLevel = kp.octave;
        if (pKF->mvuRight[idx] >= 0) {
            const float& kpx = kp.pt.x;
            const float& kpy = kp.pt.y;
            const float& kpr = pKF->mvuRight[idx];
            const float ex = u - kpx;
            const float ey = v - kpy;
            const float er = ur - kpr;
            const float e2 = ex * ex + ey * ey + er * er;
            return !(e2 * pKF->mvInvLevelSigma2[kpLevel] > 7.8);
        }
        const float& kpx = kp.pt.x;
        const float& kpy = kp.pt.y;
        const float ex = u - kpx;
        const float ey = v - kpy;
        const float e2 = ex * ex + ey * ey;
        return !(e2 * pKF->mvInvLevelSigma2[kpLevel] > 5.99);
    };
    int nFused = 0;
    for (int j = 0; j < nq; j++) {
        MapPoint* pMP = vpMapPoints[s.src[j]];
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;                     // :1187-1196 on the live objects
        const float u = s.q[4 * (size_t)j], v = s.q[4 * (size_t)j + 1], radius = s.q[4 * (size_t)j + 2], ur = urs[j];
        const int32_t* list = &s.out[(size_t)j * kFuseK * 2];
        int bestDist = 256, bestIdx = -1, valid = 0;
        for (int t = 0; t < kFuseK; t++) {
            const int idx = list[2 * t + 1];
            if (idx < 0) break;
            valid++;
            if (!passes(idx, u, v, ur)) continue;
            bestDist = list[2 * t]; bestIdx = idx;
            break;
        }
        if (bestIdx < 0 && valid == kFuseK) {                                     // the reference's own scan for this point (:1246-1309)
            const int nPredictedLevel = s.qlev[2 * (size_t)j + 1];
            const std::vector<size_t> vIndices = pKF->GetFeaturesInArea(u, v, radius, false);
            const uchar* dMP = s.qdesc.data() + (size_t)32 * j;
            for (std::vector<size_t>::const_iterator vit = vIndices.begin(), vend = vIndices.end(); vit != vend; vit++) {
                const size_t idx = *vit;
                const int& kpLevel = pKF->mvKeysUn[idx].octave;
                if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
                if (!passes((int)idx, u, v, ur)) continue;
                const int dist = orbb_hamming_distance(dMP, pKF->mDescriptors.ptr<uchar>((int)idx));
                if (dist < bestDist) { bestDist = dist; bestIdx = (int)idx; }
            }
            mnRescans++;
        }
        if (bestIdx >= 0 && bestDist <= TH_LOW) {                                 // :1312-1331
            MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) {
                    if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                    else pMPinKF->Replace(pMP);
                }
            } else {
                pMP->AddObservation(pKF, bestIdx);
                pKF->AddMapPoint(pMP, bestIdx);
            }
            nFused++;
        }
    }
    return nFused;
}

// ORBmatcher::Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, vpPoints, th, vpReplacePoint) (ORBmatcher.cc:1340-1455; LoopClosing::SearchAndFuse,
// LoopClosing.cc:3464 / :3509): the projection of the Sim3 search above, but the scan is unmasked (every key point inside the window and
// the level band competes, bestDist starts at INT_MAX) -- so the scans of all points are independent of what the loop decides, and one
// launch serves the whole call.  What the loop then does with the best key point -- replace the point it holds, or add the observation
// -- is map bookkeeping on the reference's own objects, in the reference's order.
int ORBmatcherGPU::FuseSim3(KeyFrame* pKF, const float* R9, const float* t3, float scale, const std::vector<MapPoint*>& vpPoints, float th,
                            std::vector<MapPoint*>& vpReplacePoint) {
    if (pKF->NLeft != -1) throw std::logic_error("ORBmatcherGPU::Fuse: fisheye-stereo key frames keep the reference's host path");
    Impl& s = Scratch();
    Eigen::Matrix3f Rm;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rm(i, j) = R9[3 * i + j];
    Eigen::Vector3f tv(t3[0], t3[1], t3[2]);
    Sophus::SE3f Tcw = Sophus::SE3f(Rm, tv / scale);                              // :1349
    Eigen::Vector3f Ow = Tcw.inverse().translation();
    const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
    const int nPoints = (int)vpPoints.size();
    s.q.clear(); s.qlev.clear(); s.qdesc.clear(); s.src.clear();
    for (int iMP = 0; iMP < nPoints; iMP++) {                                     // :1360-1406
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        Eigen::Vector3f p3Dw = pMP->GetWorldPos();
        Eigen::Vector3f p3Dc = Tcw * p3Dw;
        if (p3Dc(2) < 0.0f) continue;
        const Eigen::Vector2f uv = pKF->mpCamera->project(p3Dc);
        if (!pKF->IsInImage(uv(0), uv(1))) continue;
        const float maxDistance = pMP->GetMaxDistanceInvariance();
        const float minDistance = pMP->GetMinDistanceInvariance();
        Eigen::Vector3f PO = p3Dw - Ow;
        const float dist3D = PO.norm();
        if (dist3D < minDistance || dist3D > maxDistance) continue;
        Eigen::Vector3f Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist3D) continue;
        const int nPredictedLevel = pMP->PredictScale(dist3D, pKF);
        const float radius = th * pKF->mvScaleFactors[nPredictedLevel];
        const float q4[4] = {uv(0), uv(1), radius, -1.0f};
        s.q.insert(s.q.end(), q4, q4 + 4);
        s.qlev.push_back(nPredictedLevel - 1); s.qlev.push_back(nPredictedLevel);
        const cv::Mat d = pMP->GetDescriptor();
        s.qdesc.insert(s.qdesc.end(), d.ptr<uchar>(), d.ptr<uchar>() + 32);
        s.src.push_back(iMP);
    }
    const int nq = (int)s.src.size(), n = (int)pKF->mvKeysUn.size();
    if (nq == 0) return 0;
    if (!pKF->mDescriptors.isContinuous()) throw std::runtime_error("KeyFrame::mDescriptors must be continuous");
    s.xy.resize((size_t)n * 2);
    s.oct.resize(n);
    for (int i = 0; i < n; i++) { s.xy[2 * i] = pKF->mvKeysUn[i].pt.x; s.xy[2 * i + 1] = pKF->mvKeysUn[i].pt.y; s.oct[i] = pKF->mvKeysUn[i].octave; }
    orbb_frame_view view;
    view.kps_xy = s.xy.data(); view.kps_stride = 8; view.octaves = s.oct.data(); view.oct_stride = 4;
    view.desc = pKF->mDescriptors.ptr<uchar>(); view.u_right = nullptr; view.n = n; view.on_device = 0;
    const float grid4[4] = {(float)pKF->mnMinX, (float)pKF->mnMinY, pKF->mfGridElementWidthInv, pKF->mfGridElementHeightInv};
    s.out.assign((size_t)nq * 2, -1);
    if (orbb_search_area_topk(mpMatcher, &view, grid4, s.q.data(), s.qlev.data(), s.qdesc.data(), nq, nullptr, 257, 1, s.out.data()) != ORBB_OK)
        throw std::runtime_error(std::string("orbb_search_area_topk failed: ") + orbb_matcher_last_error(mpMatcher));
    int nFused = 0;
    for (int j = 0; j < nq; j++) {                                                // :1436-1451
        const int bestDist = s.out[2 * (size_t)j], bestIdx = s.out[2 * (size_t)j + 1], iMP = s.src[j];
        if (bestIdx < 0 || bestDist > TH_LOW) continue;
        MapPoint* pMP = vpPoints[iMP];
        MapPoint* pMPinKF = pKF->GetMapPoint(bestIdx);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) vpReplacePoint[iMP] = pMPinKF;
        } else {
            pMP->AddObservation(pKF, bestIdx);
            pKF->AddMapPoint(pMP, bestIdx);
        }
        nFused++;
    }
    return nFused;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (ORBmatcher.cc:223-421; Tracking::TrackReferenceKeyFrame, Tracking.cc:2769, and
// Tracking::Relocalization): every key-frame feature that holds a good map point is compared with the frame's features of the SAME vocabulary
// node.  The merge loop over the two feature vectors only collects (key-frame feature, candidate list) pairs here, in the reference's order;
// the DescriptorDistance scans of all of them are one launch (orbb_best2_csr_dev against the frame's descriptors, uploaded once per frame
// id).  The reference skips frame features that were matched earlier in the call (:278): when the best or the second candidate of a
// feature has been taken meanwhile, its (short) list is walked again on the host with the live vpMapPointMatches.
int ORBmatcherGPU::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches, const float nnratio, const bool checkOrientation) {
    if (F.Nleft != -1 || pKF->mpCamera2)
        throw std::logic_error("ORBmatcherGPU::SearchByBoW: fisheye-stereo frames / key frames keep the reference's host path");
    Impl& s = Scratch();
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    const DBoW2::FeatureVector& vFeatVecKF = pKF->mFeatVec;
    std::vector<int> rotHist[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) rotHist[i].reserve(500);
    const float factor = 1.0f / HISTO_LENGTH;
    s.qdesc.clear(); s.src.clear(); s.cand.clear(); s.rowptr.assign(1, 0);
    DBoW2::FeatureVector::const_iterator KFit = vFeatVecKF.begin(), Fit = F.mFeatVec.begin();
    const DBoW2::FeatureVector::const_iterator KFend = vFeatVecKF.end(), Fend = F.mFeatVec.end();
    while (KFit != KFend && Fit != Fend) {                                        // :244-394
        if (KFit->first == Fit->first) {
            const std::vector<unsigned int>& vIndicesKF = KFit->second;
            const std::vector<unsigned int>& vIndicesF = Fit->second;
            for (size_t iKF = 0; iKF < vIndicesKF.size(); iKF++) {
                const unsigned int realIdxKF = vIndicesKF[iKF];
                MapPoint* pMP = vpMapPointsKF[realIdxKF];
                if (!pMP || pMP->isBad()) continue;
                const uchar* d = pKF->mDescriptors.ptr<uchar>((int)realIdxKF);
                s.qdesc.insert(s.qdesc.end(), d, d + 32);
                s.src.push_back((int)realIdxKF);
                s.cand.insert(s.cand.end(), vIndicesF.begin(), vIndicesF.end());
                s.rowptr.push_back((int32_t)s.cand.size());
            }
            KFit++;
            Fit++;
        } else if (KFit->first < Fit->first) {
            KFit = vFeatVecKF.lower_bound(Fit->first);
        } else {
            Fit = F.mFeatVec.lower_bound(KFit->first);
        }
    }
    const int nq = (int)s.src.size();
    s.out.assign((size_t)nq * 4, -1);
    if (nq > 0) {
        const orbb_frame_view* fv = FrameView(mpMatcher, s, F, 0);
        if (orbb_best2_csr_dev(mpMatcher, s.qdesc.data(), nq, fv->desc, F.N, s.cand.data(), s.rowptr.data(), 256, s.out.data()) != ORBB_OK)
            throw std::runtime_error(std::string("orbb_best2_csr_dev failed: ") + orbb_matcher_last_error(mpMatcher));
    }
    int nmatches = 0;
    for (int j = 0; j < nq; j++) {                                                // :265-356 per key-frame feature, in the order of the merge loop
        const int realIdxKF = s.src[j];
        int bestDist1 = s.out[4 * (size_t)j], bestIdxF = s.out[4 * (size_t)j + 1], bestDist2 = s.out[4 * (size_t)j + 2];
        const int secondIdx = s.out[4 * (size_t)j + 3];
        if ((bestIdxF >= 0 && vpMapPointMatches[bestIdxF]) || (secondIdx >= 0 && vpMapPointMatches[secondIdx])) {
            bestDist1 = 256; bestIdxF = -1; bestDist2 = 256;                      // :273-294 on the live matches
            const uchar* dKF = s.qdesc.data() + (size_t)32 * j;
            for (int c = s.rowptr[j]; c < s.rowptr[j + 1]; c++) {
                const int realIdxF = s.cand[c];
                if (vpMapPointMatches[realIdxF]) continue;
                const int dist = orbb_hamming_distance(dKF, F.mDescriptors.ptr<uchar>(realIdxF));
                if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                else if (dist < bestDist2) bestDist2 = dist;
            }
            mnRescans++;
        }
        if (bestDist1 <= TH_LOW) {                                                // :327-356
            if (static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                vpMapPointMatches[bestIdxF] = vpMapPointsKF[realIdxKF];
                if (checkOrientation) {
                    float rot = pKF->mvKeysUn[realIdxKF].angle - F.mvKeys[bestIdxF].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = round(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotHist[bin].push_back(bestIdxF);
                }
                nmatches++;
            }
        }
    }
    if (checkOrientation) {                                                       // :396-415
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) {
                vpMapPointMatches[rotHist[i][j]] = static_cast<MapPoint*>(NULL);
                nmatches--;
            }
        }
    }
    return nmatches;
}

// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2012-2053): the three fullest bins; the second / third are dropped below 10 % of the first
void ORBmatcherGPU::ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
}

}  // namespace ORB_SLAM3

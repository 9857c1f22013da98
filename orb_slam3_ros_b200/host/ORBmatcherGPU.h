// GPU-backed building blocks for ORB_SLAM3::ORBmatcher (reference orb_slam3/include/ORBmatcher.h:38-94,
// orb_slam3/src/ORBmatcher.cc).  The twelve Search*/Fuse methods keep their host geometry (projection, grid lookup,
// MapPoint bookkeeping -- out of scope); what they all share is
//      "scan a candidate list with DescriptorDistance keeping best / second best with strict '<'"
// and that scan is what this header provides, plus the brute-force kNN-2 that Frame::ComputeStereoFishEyeMatches
// obtains from cv::BFMatcher (Frame.cc:1144).  INTEGRATION.md shows the call-site changes.
#ifndef ORBMATCHER_GPU_H
#define ORBMATCHER_GPU_H

#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include <opencv2/opencv.hpp>

#include "orbb200.h"

namespace ORB_SLAM3 {

class Frame;
class KeyFrame;
class MapPoint;

class ORBmatcherGPU {
public:
    static const int TH_LOW = 50;        // ORBmatcher.cc:36
    static const int TH_HIGH = 100;      // ORBmatcher.cc:35
    static const int HISTO_LENGTH = 30;  // ORBmatcher.cc:37

    struct Impl;                          // scratch of the compiled SearchByProjection replacements (ORBmatcherGPU.cc)

    explicit ORBmatcherGPU(int device = 0) : mpMatcher(nullptr), mpImpl(nullptr), mpImplFree(nullptr), mnRescans(0) {
        if (orbb_matcher_create(device, &mpMatcher) != ORBB_OK)
            throw std::runtime_error(std::string("orbb_matcher_create failed: ") + orbb_matcher_last_error(nullptr));
    }
    ~ORBmatcherGPU() {
        if (mpImpl && mpImplFree) mpImplFree(mpImpl);
        orbb_matcher_destroy(mpMatcher);
    }

    // ---- compiled replacements of the two per-frame ORBmatcher::SearchByProjection overloads (ORBmatcherGPU.cc) ----------------
    // one instance per calling thread (the reference constructs an ORBmatcher on the stack per call; this one owns a stream)
    static ORBmatcherGPU& Instance(int device = 0);
    // ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)   ORBmatcher.cc:43-213; nnratio = mfNNratio
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th, const bool bFarPoints, const float thFarPoints,
                           const float nnratio);
    // ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)           ORBmatcher.cc:1676-1887; checkOrientation = mbCheckOrientation
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono, const bool checkOrientation);
    // ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vbPrevMatched, vnMatches12, windowSize)  ORBmatcher.cc:648-766; nnratio = mfNNratio
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize,
                                const float nnratio, const bool checkOrientation);
    // ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)   ORBmatcher.cc:1889-2010
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist,
                           const bool checkOrientation);
    // ORBmatcher::SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, th, ratioHamming)
    //                                                                                                      ORBmatcher.cc:427-530
    // (a template on the Sim3 type so that this header needs no Sophus declaration: the transform crosses as nine + three + one floats)
    template <class Sim3T>
    int SearchByProjection(KeyFrame* pKF, Sim3T& Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th, float ratioHamming) {
        float R[9], t[3];
        const auto Rm = Scw.rotationMatrix();
        const auto tv = Scw.translation();
        for (int i = 0; i < 3; i++) {
            t[i] = tv(i);
            for (int j = 0; j < 3; j++) R[3 * i + j] = Rm(i, j);
        }
        return SearchByProjectionSim3(pKF, R, t, Scw.scale(), vpPoints, vpMatched, th, ratioHamming, nullptr, nullptr);
    }
    // ... and the overload that also reports the key frame every matched point comes from              ORBmatcher.cc:532-646
    template <class Sim3T>
    int SearchByProjection(KeyFrame* pKF, Sim3T& Scw, const std::vector<MapPoint*>& vpPoints, const std::vector<KeyFrame*>& vpPointsKFs,
                           std::vector<MapPoint*>& vpMatched, std::vector<KeyFrame*>& vpMatchedKF, int th, float ratioHamming) {
        float R[9], t[3];
        const auto Rm = Scw.rotationMatrix();
        const auto tv = Scw.translation();
        for (int i = 0; i < 3; i++) {
            t[i] = tv(i);
            for (int j = 0; j < 3; j++) R[3 * i + j] = Rm(i, j);
        }
        return SearchByProjectionSim3(pKF, R, t, Scw.scale(), vpPoints, vpMatched, th, ratioHamming, &vpPointsKFs, &vpMatchedKF);
    }
    int SearchByProjectionSim3(KeyFrame* pKF, const float* R9, const float* t3, float scale, const std::vector<MapPoint*>& vpPoints,
                               std::vector<MapPoint*>& vpMatched, int th, float ratioHamming, const std::vector<KeyFrame*>* vpPointsKFs,
                               std::vector<KeyFrame*>* vpMatchedKF);
    // ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse)                   ORBmatcher.cc:906-1146
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo,
                               const bool bCoarse, const bool checkOrientation);
    // ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th, bRight)                   ORBmatcher.cc:1148-1338 (bRight == false)
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th);
    // ORBmatcher::Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints, th, vector<MapPoint*>& vpReplacePoint)   ORBmatcher.cc:1340-1455
    template <class Sim3T>
    int Fuse(KeyFrame* pKF, Sim3T& Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
        float R[9], t[3];
        const auto Rm = Scw.rotationMatrix();
        const auto tv = Scw.translation();
        for (int i = 0; i < 3; i++) {
            t[i] = tv(i);
            for (int j = 0; j < 3; j++) R[3 * i + j] = Rm(i, j);
        }
        return FuseSim3(pKF, R, t, Scw.scale(), vpPoints, th, vpReplacePoint);
    }
    int FuseSim3(KeyFrame* pKF, const float* R9, const float* t3, float scale, const std::vector<MapPoint*>& vpPoints, float th,
                 std::vector<MapPoint*>& vpReplacePoint);
    // ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)             ORBmatcher.cc:223-421
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches, const float nnratio, const bool checkOrientation);
    // ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)              ORBmatcher.cc:765-905
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float nnratio, const bool checkOrientation);
    static void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);      // ORBmatcher.cc:2012-2053
    long Rescans() const { return mnRescans; }      // points that had to be scanned a second time (their 4 candidates did not survive the in-order decisions)
    ORBmatcherGPU(const ORBmatcherGPU&) = delete;
    ORBmatcherGPU& operator=(const ORBmatcherGPU&) = delete;

    // ORBmatcher::DescriptorDistance (ORBmatcher.cc:2058-2074): one pair stays on the host.
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return orbb_hamming_distance(a.ptr<uchar>(), b.ptr<uchar>()); }

    struct BestTwo { int bestDist, bestIdx, secondDist, secondIdx; };

    // queries: nq x 32 CV_8U; train: nt x 32 CV_8U (e.g. Frame::mDescriptors); candidates of query i are
    // cand[rowPtr[i] .. rowPtr[i+1]); both distances start at `init` (256 at ORBmatcher.cc:77, TH_LOW/TH_HIGH/INT_MAX elsewhere)
    std::vector<BestTwo> BestTwoCSR(const cv::Mat& queries, const cv::Mat& train, const std::vector<int>& cand,
                                    const std::vector<int>& rowPtr, int init = 256) {
        std::vector<BestTwo> out(queries.rows);
        if (queries.rows == 0) return out;
        if (!queries.isContinuous() || !train.isContinuous()) throw std::runtime_error("descriptor matrices must be continuous");
        const int rc = orbb_best2_csr(mpMatcher, queries.ptr<uchar>(), queries.rows, train.ptr<uchar>(), train.rows, cand.data(),
                                      rowPtr.data(), init, reinterpret_cast<int32_t*>(out.data()));
        if (rc != ORBB_OK) throw std::runtime_error(std::string("orbb_best2_csr failed: ") + orbb_matcher_last_error(mpMatcher));
        return out;
    }

    // The decision loop of ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, ..) (ORBmatcher.cc:77-141) over the
    // results of ONE batched scan of all map points of the call, in list order.  The reference updates F.mvpMapPoints while it
    // walks the list and later map points skip a key point that has just been matched (:88-90); the batched scan saw the key
    // points as they were before the call, so a query whose best or second candidate has been taken meanwhile is scanned again
    // -- `rescan(j, taken)` must return the best two of query j among its candidates with taken[idx] == 0 (excluding any other
    // candidate cannot change the best two).  `octave[idx]` = octave of key point idx (F.mvKeysUn[idx].octave); `taken` (one byte
    // per key point) starts as "holds a map point with observations" and is updated here.  Returns the number of matches;
    // matchOf[idx] = query assigned to key point idx by this call, -1 otherwise.  Pure host code.
    template <class Rescan>
    static int ResolveInOrder(std::vector<BestTwo> best, const std::vector<int>& octave, std::vector<unsigned char>& taken, float nnratio,
                              Rescan rescan, std::vector<int>& matchOf) {
        matchOf.assign(octave.size(), -1);
        int nmatches = 0;
        for (size_t j = 0; j < best.size(); j++) {
            BestTwo b = best[j];
            if ((b.bestIdx >= 0 && taken[b.bestIdx]) || (b.secondIdx >= 0 && taken[b.secondIdx])) b = rescan((int)j, taken);
            if (b.bestIdx < 0 || b.bestDist > TH_HIGH) continue;                                        // :122
            const int bestLevel = octave[b.bestIdx], bestLevel2 = b.secondIdx >= 0 ? octave[b.secondIdx] : -1;
            if (bestLevel == bestLevel2 && b.bestDist > nnratio * b.secondDist) continue;               // :124-125
            matchOf[b.bestIdx] = (int)j;                                                                // :127-128
            taken[b.bestIdx] = 1;
            nmatches++;
        }
        return nmatches;
    }

    // cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, matches, 2): idx/dist are nq x 2, missing = (-1, INT_MAX)
    void KnnMatch2(const cv::Mat& query, const cv::Mat& train, std::vector<int>& idx, std::vector<int>& dist) {
        idx.assign((size_t)query.rows * 2, -1);
        dist.assign((size_t)query.rows * 2, 0x7fffffff);
        if (query.rows == 0) return;
        const int rc = orbb_knn2(mpMatcher, query.ptr<uchar>(), query.rows, train.ptr<uchar>(), train.rows, idx.data(), dist.data());
        if (rc != ORBB_OK) throw std::runtime_error(std::string("orbb_knn2 failed: ") + orbb_matcher_last_error(mpMatcher));
    }

    // Frame::ComputeStereoFishEyeMatches (Frame.cc:1126-1166), the descriptor part: the key points inside the lapping area sit at the
    // tail of both descriptor matrices (rows monoLeft.. / monoRight.., ORBextractor.cc:1153-1162), knnMatch(.., 2) between the two
    // tails, Lowe's test (*it)[0].distance < (*it)[1].distance * 0.7 (:1151).  matches: (index in the left tail, index in the right
    // tail, distance) of the pairs that pass, in left order -- what the reference triangulates next (:1152 ff.).
    struct StereoTailMatch { int queryIdx, trainIdx, distance; };
    std::vector<StereoTailMatch> StereoFishEyeMatches(const cv::Mat& descLeft, int monoLeft, const cv::Mat& descRight, int monoRight) {
        std::vector<StereoTailMatch> out;
        if (descLeft.rows <= monoLeft || descRight.rows <= monoRight) return out;
        const cv::Mat l = descLeft.rowRange(monoLeft, descLeft.rows), r = descRight.rowRange(monoRight, descRight.rows);
        std::vector<int> idx, dist;
        KnnMatch2(l, r, idx, dist);
        for (int i = 0; i < l.rows; i++)
            if (idx[2 * i + 1] >= 0 && (double)(float)dist[2 * i] < (double)(float)dist[2 * i + 1] * 0.7)      // size() >= 2 && ratio
                out.push_back(StereoTailMatch{i, idx[2 * i], dist[2 * i]});
        return out;
    }

    // Frame::UndistortKeyPoints (Frame.cc:747-780): cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK) over the keypoints.
    // K = (fx, fy, cx, cy); dist = mDistCoef.ptr<float>(), ndist = mDistCoef.rows (4 or 5); dist[0] == 0 copies the keys (:749).
    void UndistortKeyPoints(const std::vector<cv::KeyPoint>& keys, float fx, float fy, float cx, float cy, const float* dist, int ndist,
                            std::vector<cv::KeyPoint>& keysUn) {
        keysUn = keys;
        if (keys.empty()) return;
        std::vector<float> xy(keys.size() * 2), out(keys.size() * 2);
        for (size_t i = 0; i < keys.size(); i++) { xy[2 * i] = keys[i].pt.x; xy[2 * i + 1] = keys[i].pt.y; }
        const float K4[4] = {fx, fy, cx, cy};
        const int rc = orbb_undistort_points(mpMatcher, xy.data(), (int)keys.size(), K4, dist, ndist, K4, out.data());
        if (rc != ORBB_OK) throw std::runtime_error(std::string("orbb_undistort_points failed: ") + orbb_matcher_last_error(mpMatcher));
        for (size_t i = 0; i < keys.size(); i++) { keysUn[i].pt.x = out[2 * i]; keysUn[i].pt.y = out[2 * i + 1]; }
    }

private:
    int SearchByProjectionFisheye(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th, const bool bFarPoints, const float thFarPoints,
                                  const float nnratio);
    int SearchByProjectionFisheye(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono, const bool checkOrientation);
    void RunScan(const Frame& F, int nq, int k, std::vector<int32_t>& out, bool maskTaken = true, bool rightCheck = true, int init = 256,
                 bool takenAny = false);
    void Rescan(const Frame& F, int j, int want, void* candOut, int& nout, bool rightCheck = true, bool takenAny = false);
    Impl& Scratch();
    orbb_matcher* mpMatcher;
    Impl* mpImpl;
    void (*mpImplFree)(Impl*);
    long mnRescans;
};

}  // namespace ORB_SLAM3

#endif

// ORACLE (test infrastructure, NOT product code): the reference's OWN text of the matcher / stereo / grid functions,
// compiled inside minimal stand-in classes.
//
// Frame.cc and ORBmatcher.cc cannot be compiled as a whole in this image (their headers pull in Eigen, Sophus, boost,
// Pangolin, g2o).  `make ref` therefore cuts the DEFINITIONS of the functions below out of those files at build time
// (oracle/cut_reference.py -> oracle/_ref/cut/*.inc, a git-ignored build directory removed again after linking; no reference text
// lives in the repository)
// and this file #includes them between declarations of `class Frame` / `class ORBmatcher` that carry exactly the members
// those bodies touch, with the reference's names and types (Frame.h:44-45, :214-360; ORBmatcher.h:38-106).  The bodies are
// compiled unmodified; OpenCV comes from cvshim/ as for the extractor.
//
//   ORBmatcher::TH_HIGH / TH_LOW / HISTO_LENGTH   ORBmatcher.cc:35-37
//   ORBmatcher::DescriptorDistance                ORBmatcher.cc:2058-2074
//   ORBmatcher::ComputeThreeMaxima                ORBmatcher.cc:2012-2053
//   Frame::ComputeStereoMatches                   Frame.cc:811-981
//   Frame::ComputeStereoFromRGBD                  Frame.cc:984-1005
//   Frame::AssignFeaturesToGrid / PosInGrid       Frame.cc:385-416, :725-735
//   Frame::GetFeaturesInArea                      Frame.cc:657-723
//   MapPoint::ComputeDistinctiveDescriptors       MapPoint.cc:329-403
//   ORBmatcher::ORBmatcher, RadiusByViewingCos    ORBmatcher.cc:39-41, :215-221
//   ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints)   ORBmatcher.cc:43-213
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)   ORBmatcher.cc:223-421 (over the vendored DBoW2::FeatureVector)
//   ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)   ORBmatcher.cc:1676-1887 (the
//       motion-model call of Tracking::TrackWithMotionModel, Tracking.cc:2925/:2933; Eigen / Sophus from cvshim/mini_geom.hpp)
//   ORBmatcher::SearchForInitialization           ORBmatcher.cc:648-766 (Tracking::MonocularInitialization, Tracking.cc:2527)
//   ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, th, ORBdist)   ORBmatcher.cc:1889-2010
//       (Tracking::Relocalization, Tracking.cc:3765/:3779) with MapPoint::PredictScale(float, Frame*) and GetMin/MaxDistanceInvariance,
//       MapPoint.cc:502-546
//   ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>&)   ORBmatcher.cc:765-905 (LoopClosing.cc:1680)
//   ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse)   ORBmatcher.cc:906-1146 (LocalMapping::CreateNewMapPoints,
//       LocalMapping.cc:466; the camera's epipolarConstrain is the pinhole stand-in of cvshim/mini_geom.hpp on both sides)
//   ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, th, bRight)   ORBmatcher.cc:1148-1338 (LocalMapping::SearchInNeighbors,
//       LocalMapping.cc:772/:802; MapPoint::Replace / IsInKeyFrame and KeyFrame::ReplaceMapPointMatch / EraseMapPointMatch / GetPose /
//       GetCameraCenter are stand-ins with the reference's effect on the one key frame of a test)
//   ORBmatcher::Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints, th, vpReplacePoint)   ORBmatcher.cc:1340-1455
//       (LoopClosing::SearchAndFuse, LoopClosing.cc:3464/:3509; KeyFrame::GetMapPoints / GetMapPoint / AddMapPoint and MapPoint::AddObservation
//       are plain stand-ins: map bookkeeping, not matching)
//   ORBmatcher::SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, th,
//       ratioHamming) and the overload with vpPointsKFs / vpMatchedKF   ORBmatcher.cc:427-646 (LoopClosing.cc:1773/:1795/:1982) with KeyFrame::GetFeaturesInArea / IsInImage (KeyFrame.cc:707-756)
//       and MapPoint::PredictScale(float, KeyFrame*) (MapPoint.cc:514-529)
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <vector>

#include "DBoW2/BowVector.h"
#include "DBoW2/FeatureVector.h"
#include "ORBextractor.h"
#include "mini_geom.hpp"

#define FRAME_GRID_ROWS 48      // Frame.h:44
#define FRAME_GRID_COLS 64      // Frame.h:45

using namespace std;            // as Frame.cc / ORBmatcher.cc do

namespace ORB_SLAM3 {

class Frame;
class KeyFrame;
class MapPoint;

class ORBmatcher {              // ORBmatcher.h:38-106
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                           const float thFarPoints = 50.0f);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);
    int Fuse(KeyFrame* pKF, Sophus::Sim3f& Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint);
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0, const bool bRight = false);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo,
                               const bool bCoarse = false);
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);
    int SearchByProjection(KeyFrame* pKF, Sophus::Sim3f& Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th,
                           float ratioHamming = 1.0);
    int SearchByProjection(KeyFrame* pKF, Sophus::Sim3<float>& Scw, const std::vector<MapPoint*>& vpPoints, const std::vector<KeyFrame*>& vpPointsKFs,
                           std::vector<MapPoint*>& vpMatched, std::vector<KeyFrame*>& vpMatchedKF, int th, float ratioHamming = 1.0);
    static const int TH_LOW;
    static const int TH_HIGH;
    static const int HISTO_LENGTH;
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};

class KeyFrame {                // KeyFrame.h:256, :324-334, :380-522: the members the cut functions read (the const members as plain ones)
public:
    bool isBad() { return mbBad; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPoint*> GetMapPoints() {                    // KeyFrame.cc:404-418 (the good ones)
        std::set<MapPoint*> s;
        for (size_t i = 0; i < mvpMapPoints.size(); i++) if (mvpMapPoints[i] && !mapPointIsBad(mvpMapPoints[i])) s.insert(mvpMapPoints[i]);
        return s;
    }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }                       // KeyFrame.cc:471-475
    void AddMapPoint(MapPoint* pMP, const size_t& idx) { mvpMapPoints[idx] = pMP; }              // KeyFrame.cc:350-354
    void ReplaceMapPointMatch(const int& idx, MapPoint* pMP) { mvpMapPoints[idx] = pMP; }        // KeyFrame.cc:386-389
    void EraseMapPointMatch(const int& idx) { mvpMapPoints[idx] = static_cast<MapPoint*>(NULL); }   // KeyFrame.cc:356-360
    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTcw.inverse(); }
    Sophus::SE3f GetRightPoseInverse() { return mTcw.inverse(); }
    std::vector<float> mvLevelSigma2;
    Eigen::Vector3f GetCameraCenter() { return mTcw.inverse().translation(); }                  // KeyFrame.cc:143-146 (mTwc.translation())
    Sophus::SE3f GetRightPose() { return mTcw; }                  // (declared for the body's bRight branch; the tests run NLeft == -1, bRight == false)
    Eigen::Vector3f GetRightCameraCenter() { return GetCameraCenter(); }
    Sophus::SE3f mTcw;
    float mbf = 0;
    std::vector<float> mvuRight, mvInvLevelSigma2;
    static bool mapPointIsBad(MapPoint* p);
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;
    bool IsInImage(const float& x, const float& y) const;
    int N = 0;
    int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    std::vector<std::vector<std::vector<size_t> > > mGrid, mGridRight;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    std::vector<float> mvScaleFactors;
    int mnScaleLevels = 0;
    float mfLogScaleFactor = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    std::vector<MapPoint*> mvpMapPoints;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    int NLeft = -1, NRight = -1;
    bool mbBad = false;
};

class MapPoint {                // MapPoint.h:114-207: the members the cut functions touch; the three accessors are plain stand-ins
public:
    void ComputeDistinctiveDescriptors();
    int PredictScale(const float& currentDist, Frame* pF);
    int PredictScale(const float& currentDist, KeyFrame* pKF);
    void AddObservation(KeyFrame* pKF, int idx) { mObservations[pKF] = std::tuple<int, int>(idx, -1); nObs++; }      // (stand-in for MapPoint.cc:137-166)
    bool IsInKeyFrame(KeyFrame* pKF) { return mObservations.count(pKF); }                                            // MapPoint.cc:420-424
    void Replace(MapPoint* pMP);                                    // stand-in for MapPoint.cc:248-300 (below, once KeyFrame is complete)
    MapPoint* mpReplaced = nullptr;
    Eigen::Vector3f GetNormal() { return mNormalVector; }
    Eigen::Vector3f mNormalVector;
    float GetMinDistanceInvariance();
    float GetMaxDistanceInvariance();
    float mfMinDistance = 0, mfMaxDistance = 0;
    std::mutex mMutexPos;
    int Observations() { return nObs; }
    bool isBad() { return mbBad; }
    cv::Mat GetDescriptor() { return mDescriptor.clone(); }
    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    Eigen::Vector3f mWorldPos;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = -1;
    float mTrackViewCos = 1, mTrackViewCosR = 1;
    std::map<KeyFrame*, std::tuple<int, int>> mObservations;
    cv::Mat mDescriptor;
    int nObs = 0;
    bool mbBad = false;
    std::mutex mMutexFeatures;
};

class Frame {                   // Frame.h:214-360
public:
    void ComputeStereoMatches();
    void ComputeStereoFromRGBD(const cv::Mat& imDepth);
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1,
                                     const bool bRight = false) const;

    Sophus::SE3<float> GetPose() const { return mTcw; }            // Frame.h:144-147
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }              // Frame.cc:1054
    Sophus::SE3<float> mTcw, mTrl;
    std::vector<bool> mvbOutlier;

    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    float mbf = 0, mb = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<float> mvuRight, mvDepth;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors, mDescriptorsRight;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    int mnScaleLevels = 0;
    float mfLogScaleFactor = 0;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    int Nleft = -1, Nright = -1;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
};
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
bool KeyFrame::mapPointIsBad(MapPoint* p) { return p->isBad(); }
// MapPoint::Replace (MapPoint.cc:248-300) without the found / visible counters and the descriptor update: this point goes bad, its observations
// move to pMP (or are erased where pMP is observed already)
void MapPoint::Replace(MapPoint* pMP) {
    if (pMP == this) return;
    std::map<KeyFrame*, std::tuple<int, int>> obs = mObservations;
    mObservations.clear();
    mbBad = true;
    mpReplaced = pMP;
    for (std::map<KeyFrame*, std::tuple<int, int>>::iterator mit = obs.begin(); mit != obs.end(); mit++) {
        KeyFrame* pKF = mit->first;
        const int leftIndex = std::get<0>(mit->second);
        if (!pMP->IsInKeyFrame(pKF)) {
            if (leftIndex != -1) { pKF->ReplaceMapPointMatch(leftIndex, pMP); pMP->AddObservation(pKF, leftIndex); }
        } else if (leftIndex != -1) pKF->EraseMapPointMatch(leftIndex);
    }
}

// ---- the reference's own definitions (build-time cuts) ----
#include "cut/ORBmatcher_TH_HIGH.inc"
#include "cut/ORBmatcher_TH_LOW.inc"
#include "cut/ORBmatcher_HISTO_LENGTH.inc"
#include "cut/ORBmatcher_ctor.inc"
#include "cut/ORBmatcher_SearchByProjection_local.inc"
#include "cut/ORBmatcher_RadiusByViewingCos.inc"
#include "cut/ORBmatcher_SearchByBoW_KF_F.inc"
#include "cut/ORBmatcher_SearchByBoW_KF_KF.inc"
#include "cut/ORBmatcher_SearchByProjection_motion.inc"
#include "cut/ORBmatcher_SearchForInitialization.inc"
#include "cut/ORBmatcher_SearchByProjection_reloc.inc"
#include "cut/ORBmatcher_SearchByProjection_sim3.inc"
#include "cut/ORBmatcher_SearchByProjection_sim3_kfs.inc"
#include "cut/ORBmatcher_Fuse_sim3.inc"
#include "cut/ORBmatcher_Fuse_kf.inc"
#include "cut/ORBmatcher_SearchForTriangulation.inc"
#include "cut/KeyFrame_GetFeaturesInArea.inc"
#include "cut/KeyFrame_IsInImage.inc"
#include "cut/MapPoint_PredictScale_KeyFrame.inc"
#include "cut/ORBmatcher_ComputeThreeMaxima.inc"
#include "cut/ORBmatcher_DescriptorDistance.inc"
#include "cut/Frame_AssignFeaturesToGrid.inc"
#include "cut/Frame_PosInGrid.inc"
#include "cut/Frame_GetFeaturesInArea.inc"
#include "cut/Frame_ComputeStereoMatches.inc"
#include "cut/Frame_ComputeStereoFromRGBD.inc"
#include "cut/MapPoint_ComputeDistinctiveDescriptors.inc"
#include "cut/MapPoint_PredictScale_Frame.inc"
#include "cut/MapPoint_GetMinDistanceInvariance.inc"
#include "cut/MapPoint_GetMaxDistanceInvariance.inc"

}  // namespace ORB_SLAM3

namespace {
struct PortKP { float x, y, size, angle, response; int octave; };

std::vector<cv::KeyPoint> to_keypoints(const PortKP* k, int n) {
    std::vector<cv::KeyPoint> v(n);
    for (int i = 0; i < n; i++) v[i] = cv::KeyPoint(k[i].x, k[i].y, k[i].size, k[i].angle, k[i].response, k[i].octave);
    return v;
}
cv::Mat to_descriptors(const uint8_t* d, int n) {
    cv::Mat m(std::max(n, 1), 32, CV_8U);
    if (n) memcpy(m.data, d, (size_t)n * 32);
    return m;
}
}  // namespace

extern "C" {

int refcut_constants(int* th_low, int* th_high, int* histo_length) {
    *th_low = ORB_SLAM3::ORBmatcher::TH_LOW; *th_high = ORB_SLAM3::ORBmatcher::TH_HIGH; *histo_length = ORB_SLAM3::ORBmatcher::HISTO_LENGTH;
    return 0;
}

int refcut_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    return ORB_SLAM3::ORBmatcher::DescriptorDistance(to_descriptors(a, 1), to_descriptors(b, 1));
}

// histo: counts per bin -> lists of that length (the function only looks at histo[i].size())
void refcut_three_maxima(const int32_t* counts, int L, int32_t* ind3) {
    std::vector<std::vector<int>> histo(L);
    for (int i = 0; i < L; i++) histo[i].resize(counts[i]);
    int i1 = -1, i2 = -1, i3 = -1;                                         // ORBmatcher.cc:347-349
    ORB_SLAM3::ORBmatcher().ComputeThreeMaxima(histo.data(), L, i1, i2, i3);
    ind3[0] = i1; ind3[1] = i2; ind3[2] = i3;
}

// Frame::ComputeStereoMatches on the key points / descriptors of the two eyes; extL / extR = ref_create handles that have just
// extracted the two images (their mvImagePyramid is what the function reads).  Outputs mvuRight / mvDepth (nL floats each).
int refcut_stereo(void* extL, void* extR, const PortKP* kL, const uint8_t* dL, int nL, const PortKP* kR, const uint8_t* dR, int nR, float bf,
                  float b, float* uRight, float* depth) {
    ORB_SLAM3::Frame F;
    F.mpORBextractorLeft = (ORB_SLAM3::ORBextractor*)extL;
    F.mpORBextractorRight = (ORB_SLAM3::ORBextractor*)extR;
    F.mvScaleFactors = F.mpORBextractorLeft->GetScaleFactors();            // Frame.cc:110-116
    F.mvInvScaleFactors = F.mpORBextractorLeft->GetInverseScaleFactors();
    F.mvKeys = to_keypoints(kL, nL);
    F.mvKeysRight = to_keypoints(kR, nR);
    F.mDescriptors = to_descriptors(dL, nL);
    F.mDescriptorsRight = to_descriptors(dR, nR);
    F.N = nL; F.mbf = bf; F.mb = b;
    F.ComputeStereoMatches();
    for (int i = 0; i < nL; i++) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
    return 0;
}

// Frame::ComputeStereoFromRGBD: keys (x, y), undistorted x, float depth map
int refcut_rgbd(const float* xy, const float* xUn, int n, const float* depthMap, int w, int h, size_t strideFloats, float bf, float* uRight,
                float* depth) {
    ORB_SLAM3::Frame F;
    F.N = n; F.mbf = bf;
    F.mvKeys.resize(n); F.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) { F.mvKeys[i].pt.x = xy[2 * i]; F.mvKeys[i].pt.y = xy[2 * i + 1]; F.mvKeysUn[i].pt.x = xUn[i]; F.mvKeysUn[i].pt.y = xy[2 * i + 1]; }
    cv::Mat im(h, w, CV_32F, const_cast<float*>(depthMap), strideFloats * sizeof(float));
    F.ComputeStereoFromRGBD(im);
    for (int i = 0; i < n; i++) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
    return 0;
}

// AssignFeaturesToGrid + GetFeaturesInArea for every query, followed by the best / second scan of
// ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&) (ORBmatcher.cc:77-120, restated here: that function needs MapPoint)
// with the reference's DescriptorDistance.  Same arguments and output as port_search_area_best2.
void refcut_search_area_best2(const float* kps, const int32_t* oct, const uint8_t* train, int n, const float* grid4, const float* queries,
                              const int32_t* qlev, const uint8_t* qdesc, int nq, const uint8_t* skip, const float* uRight, int init,
                              int32_t* out4) {
    using ORB_SLAM3::Frame;
    Frame* F = new Frame();                                                // (the two grids make the object large)
    Frame::mnMinX = grid4[0]; Frame::mnMinY = grid4[1]; Frame::mfGridElementWidthInv = grid4[2]; Frame::mfGridElementHeightInv = grid4[3];
    F->N = n; F->Nleft = -1;
    F->mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) { F->mvKeysUn[i].pt.x = kps[2 * i]; F->mvKeysUn[i].pt.y = kps[2 * i + 1]; F->mvKeysUn[i].octave = oct[i]; }
    F->AssignFeaturesToGrid();
    const cv::Mat tr = to_descriptors(train, n);
    for (int q = 0; q < nq; q++) {
        const float x = queries[4 * q], y = queries[4 * q + 1], r = queries[4 * q + 2], xr = queries[4 * q + 3];
        const std::vector<size_t> idxs = F->GetFeaturesInArea(x, y, r, qlev[2 * q], qlev[2 * q + 1]);
        const cv::Mat dq = to_descriptors(qdesc + (size_t)q * 32, 1);
        int bestDist = init, bestDist2 = init, bestIdx = -1, bestIdx2 = -1;
        for (size_t idx : idxs) {
            if (skip && skip[idx]) continue;
            if (uRight && uRight[idx] > 0) {
                const float er = std::fabs(xr - uRight[idx]);
                if (er > r) continue;
            }
            const int dist = ORB_SLAM3::ORBmatcher::DescriptorDistance(dq, tr.row((int)idx));
            if (dist < bestDist) { bestDist2 = bestDist; bestIdx2 = bestIdx; bestDist = dist; bestIdx = (int)idx; }
            else if (dist < bestDist2) { bestDist2 = dist; bestIdx2 = (int)idx; }
        }
        out4[4 * q] = bestDist; out4[4 * q + 1] = bestIdx; out4[4 * q + 2] = bestDist2; out4[4 * q + 3] = bestIdx2;
    }
    delete F;
}

// MapPoint::ComputeDistinctiveDescriptors for one map point observed in nkf key frames (left and right index each, -1 = none;
// desc = the key frames' descriptor matrices, 2 rows per key frame).  The observations map is ordered by KeyFrame address: the
// stand-in key frames live in one array, so that order is the index order.  -> 1 and out[32] = mDescriptor, or 0 when the
// function returned early (no usable observation).
int refcut_distinctive(const uint8_t* desc, const int32_t* leftRight, const uint8_t* bad, int nkf, uint8_t* out) {
    std::vector<ORB_SLAM3::KeyFrame> kfs(nkf);
    ORB_SLAM3::MapPoint mp;
    for (int i = 0; i < nkf; i++) {
        kfs[i].mDescriptors = to_descriptors(desc + (size_t)i * 64, 2);
        kfs[i].mbBad = bad && bad[i];
        mp.mObservations[&kfs[i]] = std::make_tuple((int)leftRight[2 * i], (int)leftRight[2 * i + 1]);
    }
    mp.ComputeDistinctiveDescriptors();
    if (mp.mDescriptor.empty()) return 0;
    memcpy(out, mp.mDescriptor.data, 32);
    return 1;
}

// Tracking::SearchLocalPoints' call: ORBmatcher(nnratio).SearchByProjection(F, vpMapPoints, th) on a monocular / rectified-stereo
// frame (Nleft == -1).  Key points (undistorted x, y, octave), descriptors, mvuRight (or null) and the frame's current matches
// (hasPoint[i] != 0: key point i already holds a map point with observations).  Map points: proj = {x, y, xR, viewCos} floats,
// level, descriptor, inView flag.  -> matchOf[n] = index of the map point assigned to key point i by this call (-1 none);
// returns nmatches.
int refcut_search_by_projection(const float* kps, const int32_t* oct, const uint8_t* train, int n, const float* grid4, const float* uRight,
                                const uint8_t* hasPoint, const float* scaleFactors, int nlevels, const float* proj, const int32_t* level,
                                const uint8_t* mpDesc, const uint8_t* inView, int nmp, float nnratio, float th, int32_t* matchOf) {
    using namespace ORB_SLAM3;
    Frame* F = new Frame();
    Frame::mnMinX = grid4[0]; Frame::mnMinY = grid4[1]; Frame::mfGridElementWidthInv = grid4[2]; Frame::mfGridElementHeightInv = grid4[3];
    F->N = n; F->Nleft = -1;
    F->mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) { F->mvKeysUn[i].pt.x = kps[2 * i]; F->mvKeysUn[i].pt.y = kps[2 * i + 1]; F->mvKeysUn[i].octave = oct[i]; }
    F->AssignFeaturesToGrid();
    F->mDescriptors = to_descriptors(train, n);
    F->mvuRight.assign(n, -1.0f);
    if (uRight) for (int i = 0; i < n; i++) F->mvuRight[i] = uRight[i];
    F->mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    MapPoint old;                                                      // what the key points matched earlier hold
    old.nObs = 1;
    F->mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++) if (hasPoint && hasPoint[i]) F->mvpMapPoints[i] = &old;
    std::vector<MapPoint> mps(nmp);
    std::vector<MapPoint*> vp(nmp);
    for (int j = 0; j < nmp; j++) {
        MapPoint& m = mps[j];
        m.mTrackProjX = proj[4 * j]; m.mTrackProjY = proj[4 * j + 1]; m.mTrackProjXR = proj[4 * j + 2]; m.mTrackViewCos = proj[4 * j + 3];
        m.mnTrackScaleLevel = level[j];
        m.mbTrackInView = inView[j] != 0;
        m.mDescriptor = to_descriptors(mpDesc + (size_t)32 * j, 1);
        m.nObs = 1;                                                    // local map points have observations
        vp[j] = &m;
    }
    ORBmatcher matcher(nnratio);
    const int nmatches = matcher.SearchByProjection(*F, vp, th);
    for (int i = 0; i < n; i++) {
        MapPoint* p = F->mvpMapPoints[i];
        matchOf[i] = (p && p != &old) ? (int)(p - mps.data()) : -1;
    }
    delete F;
    return nmatches;
}

// The same call on a fisheye-stereo frame (Nleft != -1, the KannalaBrandt8 rigs of config/Stereo/TUM-VI.yaml): left key points mvKeys[0 .. nL) and
// right key points mvKeysRight[0 .. nR) with their own grids, descriptors stacked left then right, mvLeftToRightMatch / mvRightToLeftMatch of the
// stereo pairs; hasPoint over all nL + nR entries.  Map points: projL = {x, y, viewCos}, levelL, inViewL, projR = {xR, yR, viewCosR}, levelR (-1: none),
// inViewR.  -> matchOf[nL + nR]; returns nmatches.
int refcut_search_by_projection_fisheye(const float* kpsL, const int32_t* octL, int nL, const float* kpsR, const int32_t* octR, int nR, const uint8_t* desc,
                                        const float* fp, const int32_t* l2r, const int32_t* r2l, const uint8_t* hasPoint, const float* scaleFactors,
                                        int nlevels, const float* projL, const int32_t* levelL, const uint8_t* inViewL, const float* projR,
                                        const int32_t* levelR, const uint8_t* inViewR, const uint8_t* mpDesc, int nmp, float nnratio, float th,
                                        int32_t* matchOf) {
    using namespace ORB_SLAM3;
    Frame* F = new Frame();
    Frame::mnMinX = fp[0]; Frame::mnMaxX = fp[1]; Frame::mnMinY = fp[2]; Frame::mnMaxY = fp[3];
    Frame::mfGridElementWidthInv = fp[4]; Frame::mfGridElementHeightInv = fp[5];
    const int n = nL + nR;
    F->N = n; F->Nleft = nL; F->Nright = nR;
    F->mvKeys.resize(nL); F->mvKeysRight.resize(nR);
    for (int i = 0; i < nL; i++) { F->mvKeys[i].pt.x = kpsL[2 * i]; F->mvKeys[i].pt.y = kpsL[2 * i + 1]; F->mvKeys[i].octave = octL[i]; }
    for (int i = 0; i < nR; i++) { F->mvKeysRight[i].pt.x = kpsR[2 * i]; F->mvKeysRight[i].pt.y = kpsR[2 * i + 1]; F->mvKeysRight[i].octave = octR[i]; }
    F->AssignFeaturesToGrid();
    F->mDescriptors = to_descriptors(desc, n);
    F->mvuRight.assign(n, -1.0f);
    F->mvLeftToRightMatch.assign(l2r, l2r + nL);
    F->mvRightToLeftMatch.assign(r2l, r2l + nR);
    F->mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    MapPoint old;
    old.nObs = 1;
    F->mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++) if (hasPoint && hasPoint[i]) F->mvpMapPoints[i] = &old;
    std::vector<MapPoint> mps(nmp);
    std::vector<MapPoint*> vp(nmp);
    for (int j = 0; j < nmp; j++) {
        MapPoint& m = mps[j];
        m.mTrackProjX = projL[3 * j]; m.mTrackProjY = projL[3 * j + 1]; m.mTrackViewCos = projL[3 * j + 2];
        m.mnTrackScaleLevel = levelL[j]; m.mbTrackInView = inViewL[j] != 0;
        m.mTrackProjXR = projR[3 * j]; m.mTrackProjYR = projR[3 * j + 1]; m.mTrackViewCosR = projR[3 * j + 2];
        m.mnTrackScaleLevelR = levelR[j]; m.mbTrackInViewR = inViewR[j] != 0;
        m.mDescriptor = to_descriptors(mpDesc + (size_t)32 * j, 1);
        m.nObs = 1;
        vp[j] = &m;
    }
    ORBmatcher matcher(nnratio);
    const int nmatches = matcher.SearchByProjection(*F, vp, th);
    for (int i = 0; i < n; i++) {
        MapPoint* p = F->mvpMapPoints[i];
        matchOf[i] = (p && p != &old) ? (int)(p - mps.data()) : -1;
    }
    delete F;
    return nmatches;
}

// Tracking::TrackWithMotionModel's call: ORBmatcher(nnratio, checkOri).SearchByProjection(CurrentFrame, LastFrame, th, bMono) on
// monocular / rectified-stereo / RGB-D frames (Nleft == -1).
//   current frame: undistorted key points (x, y), octaves, angles, descriptors, mvuRight (or null), curState[i] = 0 no map point /
//     1 a map point with observations / 2 a map point without (e.g. a temporal stereo point); fp = {mnMinX, mnMaxX, mnMinY, mnMaxY,
//     mfGridElementWidthInv, mfGridElementHeightInv, mbf, mb}; pose Tcw = {R (9, row major), t (3)}; pinhole cam4 = {fx, fy, cx, cy}
//   last frame: per feature its octave, angle, whether it holds a map point (lastState: 0 none / 1 with / 2 without observations),
//     the outlier flag, the map point's world position and descriptor; pose Tlw
// -> matchOf[i] = last-frame feature whose map point key point i holds after the call and did not hold before (-1 otherwise);
//    returns nmatches.
int refcut_search_by_projection_motion(const float* kps, const int32_t* oct, const float* angle, const uint8_t* desc, int n, const float* fp,
                                       const float* uRight, const uint8_t* curState, const float* scaleFactors, int nlevels, const float* Tcw,
                                       const float* cam4, int nLast, const int32_t* lastOct, const float* lastAngle, const uint8_t* lastState,
                                       const uint8_t* lastOutlier, const float* lastPos, const uint8_t* lastDesc, const float* Tlw, float th,
                                       int bMono, float nnratio, int checkOri, int32_t* matchOf) {
    using namespace ORB_SLAM3;
    Frame* C = new Frame();
    Frame* L = new Frame();
    Frame::mnMinX = fp[0]; Frame::mnMaxX = fp[1]; Frame::mnMinY = fp[2]; Frame::mnMaxY = fp[3];
    Frame::mfGridElementWidthInv = fp[4]; Frame::mfGridElementHeightInv = fp[5];
    GeometricCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    C->N = n; C->Nleft = -1; C->mbf = fp[6]; C->mb = fp[7]; C->mpCamera = &cam;
    C->mTcw = Sophus::SE3f(Tcw, Tcw + 9);
    C->mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) {
        C->mvKeysUn[i].pt.x = kps[2 * i]; C->mvKeysUn[i].pt.y = kps[2 * i + 1]; C->mvKeysUn[i].octave = oct[i]; C->mvKeysUn[i].angle = angle[i];
    }
    C->mvKeys = C->mvKeysUn;
    C->AssignFeaturesToGrid();
    C->mDescriptors = to_descriptors(desc, n);
    C->mvuRight.assign(n, -1.0f);
    if (uRight) for (int i = 0; i < n; i++) C->mvuRight[i] = uRight[i];
    C->mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    MapPoint oldObs, oldNoObs;                                         // what the current frame's key points hold before the call
    oldObs.nObs = 1; oldNoObs.nObs = 0;
    C->mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++) if (curState && curState[i]) C->mvpMapPoints[i] = curState[i] == 1 ? &oldObs : &oldNoObs;
    L->N = nLast; L->Nleft = -1; L->mTcw = Sophus::SE3f(Tlw, Tlw + 9);
    L->mvKeysUn.resize(nLast);
    std::vector<MapPoint> mps(nLast);
    L->mvpMapPoints.assign(nLast, nullptr);
    L->mvbOutlier.assign(nLast, false);
    for (int j = 0; j < nLast; j++) {
        L->mvKeysUn[j].octave = lastOct[j]; L->mvKeysUn[j].angle = lastAngle[j];
        L->mvbOutlier[j] = lastOutlier[j] != 0;
        if (lastState[j]) {
            mps[j].nObs = lastState[j] == 1 ? 1 : 0;
            mps[j].mWorldPos = Eigen::Vector3f(lastPos[3 * j], lastPos[3 * j + 1], lastPos[3 * j + 2]);
            mps[j].mDescriptor = to_descriptors(lastDesc + (size_t)32 * j, 1);
            L->mvpMapPoints[j] = &mps[j];
        }
    }
    L->mvKeys = L->mvKeysUn;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int nmatches = matcher.SearchByProjection(*C, *L, th, bMono != 0);
    for (int i = 0; i < n; i++) {
        MapPoint* p = C->mvpMapPoints[i];
        matchOf[i] = (p && p != &oldObs && p != &oldNoObs) ? (int)(p - mps.data()) : -1;
    }
    delete C;
    delete L;
    return nmatches;
}

// Tracking::Relocalization's refinement call (Tracking.cc:3765, :3779): ORBmatcher(0.9, true).SearchByProjection(CurrentFrame, pKF, sFound, th, ORBdist)
// (ORBmatcher.cc:1889-2010).  Current frame as in refcut_search_by_projection_motion, curHolds[i] != 0: key point i already holds a map point
// (any: the scan skips non-null entries, :1952); fp additionally carries {.., mnScaleLevels, mfLogScaleFactor} at [8], [9].  Key frame: per
// feature kfState (0 no map point / 1 good / 2 bad / 3 in sAlreadyFound), angle, world position, descriptor, mfMinDistance, mfMaxDistance.
// -> matchOf[i] = key-frame feature whose map point key point i received in this call (-1 otherwise); returns nmatches.
int refcut_search_by_projection_reloc(const float* kps, const int32_t* oct, const float* angle, const uint8_t* desc, int n, const float* fp,
                                      const uint8_t* curHolds, const float* scaleFactors, int nlevels, const float* Tcw, const float* cam4, int nK,
                                      const float* kfAngle, const uint8_t* kfState, const float* kfPos, const uint8_t* kfDesc, const float* kfMinDist,
                                      const float* kfMaxDist, float th, int ORBdist, float nnratio, int checkOri, int32_t* matchOf) {
    using namespace ORB_SLAM3;
    Frame* Cf = new Frame();
    Frame::mnMinX = fp[0]; Frame::mnMaxX = fp[1]; Frame::mnMinY = fp[2]; Frame::mnMaxY = fp[3];
    Frame::mfGridElementWidthInv = fp[4]; Frame::mfGridElementHeightInv = fp[5];
    GeometricCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    Cf->N = n; Cf->Nleft = -1; Cf->mpCamera = &cam;
    Cf->mnScaleLevels = (int)fp[8]; Cf->mfLogScaleFactor = fp[9];
    Cf->mTcw = Sophus::SE3f(Tcw, Tcw + 9);
    Cf->mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) {
        Cf->mvKeysUn[i].pt.x = kps[2 * i]; Cf->mvKeysUn[i].pt.y = kps[2 * i + 1]; Cf->mvKeysUn[i].octave = oct[i]; Cf->mvKeysUn[i].angle = angle[i];
    }
    Cf->mvKeys = Cf->mvKeysUn;
    Cf->AssignFeaturesToGrid();
    Cf->mDescriptors = to_descriptors(desc, n);
    Cf->mvuRight.assign(n, -1.0f);
    Cf->mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    MapPoint held;
    Cf->mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++) if (curHolds && curHolds[i]) Cf->mvpMapPoints[i] = &held;
    KeyFrame kf;
    std::vector<MapPoint> mps(nK);
    std::set<MapPoint*> sFound;
    kf.mvKeysUn.resize(nK); kf.mvpMapPoints.assign(nK, nullptr);
    for (int j = 0; j < nK; j++) {
        kf.mvKeysUn[j].angle = kfAngle[j];
        if (!kfState[j]) continue;
        mps[j].mbBad = kfState[j] == 2;
        mps[j].mWorldPos = Eigen::Vector3f(kfPos[3 * j], kfPos[3 * j + 1], kfPos[3 * j + 2]);
        mps[j].mDescriptor = to_descriptors(kfDesc + (size_t)32 * j, 1);
        mps[j].mfMinDistance = kfMinDist[j]; mps[j].mfMaxDistance = kfMaxDist[j];
        kf.mvpMapPoints[j] = &mps[j];
        if (kfState[j] == 3) sFound.insert(&mps[j]);
    }
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int nmatches = matcher.SearchByProjection(*Cf, &kf, sFound, th, ORBdist);
    for (int i = 0; i < n; i++) {
        MapPoint* p = Cf->mvpMapPoints[i];
        matchOf[i] = (p && p != &held) ? (int)(p - mps.data()) : -1;
    }
    delete Cf;
    return nmatches;
}

// LoopClosing's Sim3 projection search (LoopClosing.cc:1795 / :1982): ORBmatcher(0.9, true).SearchByProjection(pKF, Scw, vpPoints, vpMatched, th,
// ratioHamming) (ORBmatcher.cc:427-530).  Key frame: undistorted key points (x, y), octaves, descriptors, bounds / grid fp = {mnMinX, mnMaxX,
// mnMinY, mnMaxY (integers in KeyFrame.h), gridWInv, gridHInv, 0, 0, mnScaleLevels, mfLogScaleFactor}, scale factors, pinhole cam4; held[i] != 0:
// key point i is matched already (vpMatched[i] non-null, a map point that is not among vpPoints).  Sim3 = {R (9), t (3), s}.  Map points: state
// (1 good / 2 bad), world position, normal, descriptor, mfMinDistance, mfMaxDistance.  -> matchOf[i] = map point that key point i received in
// this call (-1 otherwise); returns nmatches.
static int sim3_search(int withKFs, const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held,
                       const float* scaleFactors, int nlevels, const float* sim3, const float* cam4, int nP, const uint8_t* pState, const float* pPos,
                       const float* pNormal, const uint8_t* pDesc, const float* pMinDist, const float* pMaxDist, int th, float ratioHamming,
                       int32_t* matchOf, int32_t* matchKF);
int refcut_search_by_projection_sim3(const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held,
                                     const float* scaleFactors, int nlevels, const float* sim3, const float* cam4, int nP, const uint8_t* pState,
                                     const float* pPos, const float* pNormal, const uint8_t* pDesc, const float* pMinDist, const float* pMaxDist, int th,
                                     float ratioHamming, int32_t* matchOf) {
    return sim3_search(0, kps, oct, desc, n, fp, held, scaleFactors, nlevels, sim3, cam4, nP, pState, pPos, pNormal, pDesc, pMinDist, pMaxDist, th,
                       ratioHamming, matchOf, nullptr);
}
// the overload that also reports the key frame each matched point came from (ORBmatcher.cc:532-646, LoopClosing.cc:1773): point j comes from
// "key frame" j % 7 of a small pool; matchKF[i] = that index for the key points matched in this call (-1 otherwise)
int refcut_search_by_projection_sim3_kfs(const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held,
                                         const float* scaleFactors, int nlevels, const float* sim3, const float* cam4, int nP, const uint8_t* pState,
                                         const float* pPos, const float* pNormal, const uint8_t* pDesc, const float* pMinDist, const float* pMaxDist,
                                         int th, float ratioHamming, int32_t* matchOf, int32_t* matchKF) {
    return sim3_search(1, kps, oct, desc, n, fp, held, scaleFactors, nlevels, sim3, cam4, nP, pState, pPos, pNormal, pDesc, pMinDist, pMaxDist, th,
                       ratioHamming, matchOf, matchKF);
}
static int sim3_search(int withKFs, const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held,
                       const float* scaleFactors, int nlevels, const float* sim3, const float* cam4, int nP, const uint8_t* pState, const float* pPos,
                       const float* pNormal, const uint8_t* pDesc, const float* pMinDist, const float* pMaxDist, int th, float ratioHamming,
                       int32_t* matchOf, int32_t* matchKF) {
    using namespace ORB_SLAM3;
    KeyFrame kf;
    GeometricCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    kf.fx = cam4[0]; kf.fy = cam4[1]; kf.cx = cam4[2]; kf.cy = cam4[3];
    kf.mpCamera = &cam;
    kf.N = n; kf.NLeft = -1;
    kf.mnMinX = (int)fp[0]; kf.mnMaxX = (int)fp[1]; kf.mnMinY = (int)fp[2]; kf.mnMaxY = (int)fp[3];
    kf.mfGridElementWidthInv = fp[4]; kf.mfGridElementHeightInv = fp[5];
    kf.mnScaleLevels = (int)fp[8]; kf.mfLogScaleFactor = fp[9];
    kf.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    kf.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) { kf.mvKeysUn[i].pt.x = kps[2 * i]; kf.mvKeysUn[i].pt.y = kps[2 * i + 1]; kf.mvKeysUn[i].octave = oct[i]; }
    kf.mDescriptors = to_descriptors(desc, n);
    {   // the grid a KeyFrame copies from its Frame (KeyFrame.cc:60-71 <- Frame::AssignFeaturesToGrid): the reference's own assignment
        Frame* F = new Frame();
        Frame::mnMinX = fp[0]; Frame::mnMaxX = fp[1]; Frame::mnMinY = fp[2]; Frame::mnMaxY = fp[3];
        Frame::mfGridElementWidthInv = fp[4]; Frame::mfGridElementHeightInv = fp[5];
        F->N = n; F->Nleft = -1; F->mvKeysUn = kf.mvKeysUn;
        F->AssignFeaturesToGrid();
        kf.mGrid.resize(kf.mnGridCols);
        for (int i = 0; i < kf.mnGridCols; i++) {
            kf.mGrid[i].resize(kf.mnGridRows);
            for (int j = 0; j < kf.mnGridRows; j++) kf.mGrid[i][j] = F->mGrid[i][j];
        }
        delete F;
    }
    MapPoint other;
    std::vector<MapPoint*> vpMatched(n, nullptr);
    for (int i = 0; i < n; i++) if (held && held[i]) vpMatched[i] = &other;
    std::vector<MapPoint> mps(nP);
    std::vector<MapPoint*> vpPoints(nP);
    for (int j = 0; j < nP; j++) {
        mps[j].mbBad = pState[j] == 2;
        mps[j].mWorldPos = Eigen::Vector3f(pPos[3 * j], pPos[3 * j + 1], pPos[3 * j + 2]);
        mps[j].mNormalVector = Eigen::Vector3f(pNormal[3 * j], pNormal[3 * j + 1], pNormal[3 * j + 2]);
        mps[j].mDescriptor = to_descriptors(pDesc + (size_t)32 * j, 1);
        mps[j].mfMinDistance = pMinDist[j]; mps[j].mfMaxDistance = pMaxDist[j];
        vpPoints[j] = &mps[j];
    }
    Sophus::Sim3f Scw;
    for (int i = 0; i < 9; i++) Scw.R.m[i] = sim3[i];
    Scw.t = Eigen::Vector3f(sim3[9], sim3[10], sim3[11]);
    Scw.s = sim3[12];
    ORBmatcher matcher(0.9f, true);
    int nmatches;
    if (withKFs) {
        KeyFrame pool[7];
        std::vector<KeyFrame*> vpPointsKFs(nP), vpMatchedKF(n, nullptr);
        for (int j = 0; j < nP; j++) vpPointsKFs[j] = &pool[j % 7];
        nmatches = matcher.SearchByProjection(&kf, Scw, vpPoints, vpPointsKFs, vpMatched, vpMatchedKF, th, ratioHamming);
        for (int i = 0; i < n; i++) matchKF[i] = vpMatchedKF[i] ? (int)(vpMatchedKF[i] - pool) : -1;
    } else {
        nmatches = matcher.SearchByProjection(&kf, Scw, vpPoints, vpMatched, th, ratioHamming);
    }
    for (int i = 0; i < n; i++) matchOf[i] = (vpMatched[i] && vpMatched[i] != &other) ? (int)(vpMatched[i] - mps.data()) : -1;
    return nmatches;
}

// LoopClosing::SearchAndFuse's call (LoopClosing.cc:3464 / :3509): ORBmatcher(0.8).Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (ORBmatcher.cc:1340-1455).
// Key frame and map points as in refcut_search_by_projection_sim3; held[i]: 0 key point i holds no map point / 1 a good one / 2 a bad one (those are
// points that are NOT among vpPoints).  -> replaceOf[j] = key point whose map point replaces point j (-1 none), addedAt[j] = key point that
// received point j as a new observation (-1 none); returns nFused.
int refcut_fuse_sim3(const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held, const float* scaleFactors,
                     int nlevels, const float* sim3, const float* cam4, int nP, const uint8_t* pState, const float* pPos, const float* pNormal,
                     const uint8_t* pDesc, const float* pMinDist, const float* pMaxDist, float th, int32_t* replaceOf, int32_t* addedAt) {
    using namespace ORB_SLAM3;
    KeyFrame kf;
    GeometricCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    kf.fx = cam4[0]; kf.fy = cam4[1]; kf.cx = cam4[2]; kf.cy = cam4[3];
    kf.mpCamera = &cam;
    kf.N = n; kf.NLeft = -1;
    kf.mnMinX = (int)fp[0]; kf.mnMaxX = (int)fp[1]; kf.mnMinY = (int)fp[2]; kf.mnMaxY = (int)fp[3];
    kf.mfGridElementWidthInv = fp[4]; kf.mfGridElementHeightInv = fp[5];
    kf.mnScaleLevels = (int)fp[8]; kf.mfLogScaleFactor = fp[9];
    kf.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    kf.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) { kf.mvKeysUn[i].pt.x = kps[2 * i]; kf.mvKeysUn[i].pt.y = kps[2 * i + 1]; kf.mvKeysUn[i].octave = oct[i]; }
    kf.mDescriptors = to_descriptors(desc, n);
    {
        Frame* F = new Frame();
        Frame::mnMinX = fp[0]; Frame::mnMaxX = fp[1]; Frame::mnMinY = fp[2]; Frame::mnMaxY = fp[3];
        Frame::mfGridElementWidthInv = fp[4]; Frame::mfGridElementHeightInv = fp[5];
        F->N = n; F->Nleft = -1; F->mvKeysUn = kf.mvKeysUn;
        F->AssignFeaturesToGrid();
        kf.mGrid.resize(kf.mnGridCols);
        for (int i = 0; i < kf.mnGridCols; i++) {
            kf.mGrid[i].resize(kf.mnGridRows);
            for (int j = 0; j < kf.mnGridRows; j++) kf.mGrid[i][j] = F->mGrid[i][j];
        }
        delete F;
    }
    std::vector<MapPoint> own(n);                                      // what the key points hold before the call
    kf.mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++) if (held && held[i]) { own[i].mbBad = held[i] == 2; kf.mvpMapPoints[i] = &own[i]; }
    std::vector<MapPoint> mps(nP);
    std::vector<MapPoint*> vpPoints(nP), vpReplace(nP, nullptr);
    for (int j = 0; j < nP; j++) {
        mps[j].mbBad = pState[j] == 2;
        mps[j].mWorldPos = Eigen::Vector3f(pPos[3 * j], pPos[3 * j + 1], pPos[3 * j + 2]);
        mps[j].mNormalVector = Eigen::Vector3f(pNormal[3 * j], pNormal[3 * j + 1], pNormal[3 * j + 2]);
        mps[j].mDescriptor = to_descriptors(pDesc + (size_t)32 * j, 1);
        mps[j].mfMinDistance = pMinDist[j]; mps[j].mfMaxDistance = pMaxDist[j];
        vpPoints[j] = &mps[j];
    }
    Sophus::Sim3f Scw;
    for (int i = 0; i < 9; i++) Scw.R.m[i] = sim3[i];
    Scw.t = Eigen::Vector3f(sim3[9], sim3[10], sim3[11]);
    Scw.s = sim3[12];
    ORBmatcher matcher(0.8f, true);
    const int nFused = matcher.Fuse(&kf, Scw, vpPoints, th, vpReplace);
    for (int j = 0; j < nP; j++) {
        // (a point can be replaced by one that an earlier iteration of this call added to the key frame: reported as 1000000 + its index)
        replaceOf[j] = !vpReplace[j] ? -1 : (vpReplace[j] >= own.data() && vpReplace[j] < own.data() + n) ? (int)(vpReplace[j] - own.data())
                                                                                                    : 1000000 + (int)(vpReplace[j] - mps.data());
        addedAt[j] = mps[j].mObservations.count(&kf) ? std::get<0>(mps[j].mObservations[&kf]) : -1;
    }
    return nFused;
}

// LocalMapping::SearchInNeighbors' calls (LocalMapping.cc:772 / :802): ORBmatcher().Fuse(pKF, vpMapPoints, th) (ORBmatcher.cc:1148-1338) on a monocular
// / rectified-stereo key frame (NLeft == -1, bRight == false).  Key frame as in refcut_fuse_sim3 plus its pose Tcw (12), mbf at fp[6], mvuRight (or null)
// and mvInvLevelSigma2; held[i]: 0 none / 1 a good map point / 2 a bad one, heldObs[i] = Observations() of that point.  Map points: state 0 = null
// entry of the list / 1 good / 2 bad, their Observations() in pObs.  -> kpHolds[i] = what key point i holds after the call (-1 nothing, i' < 1000000: its
// own point i', 1000000 + j: list point j), ownBad[i] / ptBad[j] = bad flags after the call; returns nFused.
int refcut_fuse_kf(const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held, const int32_t* heldObs,
                   const float* uRight, const float* invSigma2, const float* scaleFactors, int nlevels, const float* Tcw, const float* cam4, int nP,
                   const uint8_t* pState, const int32_t* pObs, const float* pPos, const float* pNormal, const uint8_t* pDesc, const float* pMinDist,
                   const float* pMaxDist, float th, int32_t* kpHolds, uint8_t* ownBad, uint8_t* ptBad) {
    using namespace ORB_SLAM3;
    KeyFrame kf;
    GeometricCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    kf.fx = cam4[0]; kf.fy = cam4[1]; kf.cx = cam4[2]; kf.cy = cam4[3];
    kf.mpCamera = &cam;
    kf.N = n; kf.NLeft = -1;
    kf.mnMinX = (int)fp[0]; kf.mnMaxX = (int)fp[1]; kf.mnMinY = (int)fp[2]; kf.mnMaxY = (int)fp[3];
    kf.mfGridElementWidthInv = fp[4]; kf.mfGridElementHeightInv = fp[5];
    kf.mbf = fp[6];
    kf.mnScaleLevels = (int)fp[8]; kf.mfLogScaleFactor = fp[9];
    kf.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    kf.mvInvLevelSigma2.assign(invSigma2, invSigma2 + nlevels);
    kf.mTcw = Sophus::SE3f(Tcw, Tcw + 9);
    kf.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) { kf.mvKeysUn[i].pt.x = kps[2 * i]; kf.mvKeysUn[i].pt.y = kps[2 * i + 1]; kf.mvKeysUn[i].octave = oct[i]; }
    kf.mvuRight.assign(n, -1.0f);
    if (uRight) kf.mvuRight.assign(uRight, uRight + n);
    kf.mDescriptors = to_descriptors(desc, n);
    {
        Frame* F = new Frame();
        Frame::mnMinX = fp[0]; Frame::mnMaxX = fp[1]; Frame::mnMinY = fp[2]; Frame::mnMaxY = fp[3];
        Frame::mfGridElementWidthInv = fp[4]; Frame::mfGridElementHeightInv = fp[5];
        F->N = n; F->Nleft = -1; F->mvKeysUn = kf.mvKeysUn;
        F->AssignFeaturesToGrid();
        kf.mGrid.resize(kf.mnGridCols);
        for (int i = 0; i < kf.mnGridCols; i++) {
            kf.mGrid[i].resize(kf.mnGridRows);
            for (int j = 0; j < kf.mnGridRows; j++) kf.mGrid[i][j] = F->mGrid[i][j];
        }
        delete F;
    }
    std::vector<MapPoint> own(n);
    kf.mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++)
        if (held && held[i]) {
            own[i].mbBad = held[i] == 2;
            own[i].nObs = heldObs[i];
            own[i].mObservations[&kf] = std::tuple<int, int>(i, -1);
            kf.mvpMapPoints[i] = &own[i];
        }
    std::vector<MapPoint> mps(nP);
    std::vector<MapPoint*> vpPoints(nP, nullptr);
    for (int j = 0; j < nP; j++) {
        if (!pState[j]) continue;
        mps[j].mbBad = pState[j] == 2;
        mps[j].nObs = pObs[j];
        mps[j].mWorldPos = Eigen::Vector3f(pPos[3 * j], pPos[3 * j + 1], pPos[3 * j + 2]);
        mps[j].mNormalVector = Eigen::Vector3f(pNormal[3 * j], pNormal[3 * j + 1], pNormal[3 * j + 2]);
        mps[j].mDescriptor = to_descriptors(pDesc + (size_t)32 * j, 1);
        mps[j].mfMinDistance = pMinDist[j]; mps[j].mfMaxDistance = pMaxDist[j];
        vpPoints[j] = &mps[j];
    }
    ORBmatcher matcher(0.6f, true);
    const int nFused = matcher.Fuse(&kf, vpPoints, th);
    for (int i = 0; i < n; i++) {
        MapPoint* p = kf.mvpMapPoints[i];
        kpHolds[i] = !p ? -1 : (p >= own.data() && p < own.data() + n) ? (int)(p - own.data()) : 1000000 + (int)(p - mps.data());
        ownBad[i] = own[i].mbBad;
    }
    for (int j = 0; j < nP; j++) ptBad[j] = mps[j].mbBad;
    return nFused;
}

// LocalMapping::CreateNewMapPoints' call (LocalMapping.cc:466): ORBmatcher(0.6, false).SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo,
// bCoarse) (ORBmatcher.cc:906-1146) on two monocular / rectified-stereo key frames.  Per key frame: undistorted key points (x, y), octaves, angles,
// descriptors, hasPoint (the key point holds a map point), mvuRight (or null), feature vector, pose Tcw (12), level sigma^2, scale factors, pinhole cam4.
// -> pairs[2 * k], pairs[2 * k + 1] = the k-th matched pair (ascending idx1); returns nmatches.
int refcut_search_for_triangulation(const float* kps1, const int32_t* oct1, const float* angle1, const uint8_t* desc1, const uint8_t* has1, const float* ur1,
                                    int n1, const int32_t* node1, const int32_t* start1, const int32_t* feat1, int nodes1, int feats1, const float* T1,
                                    const float* kps2, const int32_t* oct2, const float* angle2, const uint8_t* desc2, const uint8_t* has2, const float* ur2,
                                    int n2, const int32_t* node2, const int32_t* start2, const int32_t* feat2, int nodes2, int feats2, const float* T2,
                                    const float* sigma2, const float* scaleFactors, int nlevels, const float* cam4, int onlyStereo, int coarse, int checkOri,
                                    int32_t* pairs) {
    using namespace ORB_SLAM3;
    GeometricCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    KeyFrame k1, k2;
    MapPoint some;
    auto fill = [&](KeyFrame& k, const float* kps, const int32_t* oct, const float* angle, const uint8_t* desc, const uint8_t* has, const float* ur, int n,
                    const int32_t* node, const int32_t* start, const int32_t* feat, int nodes, int feats, const float* T) {
        k.mpCamera = &cam; k.N = n; k.NLeft = -1;
        k.mTcw = Sophus::SE3f(T, T + 9);
        k.mvKeysUn.resize(n); k.mvpMapPoints.assign(n, nullptr);
        for (int i = 0; i < n; i++) {
            k.mvKeysUn[i].pt.x = kps[2 * i]; k.mvKeysUn[i].pt.y = kps[2 * i + 1]; k.mvKeysUn[i].octave = oct[i]; k.mvKeysUn[i].angle = angle[i];
            if (has[i]) k.mvpMapPoints[i] = &some;
        }
        k.mvuRight.assign(n, -1.0f);
        if (ur) k.mvuRight.assign(ur, ur + n);
        k.mDescriptors = to_descriptors(desc, n);
        k.mvLevelSigma2.assign(sigma2, sigma2 + nlevels);
        k.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
        for (int g = 0; g < nodes; g++)
            for (int f = start[g]; f < (g + 1 < nodes ? start[g + 1] : feats); f++) k.mFeatVec.addFeature(node[g], feat[f]);
    };
    fill(k1, kps1, oct1, angle1, desc1, has1, ur1, n1, node1, start1, feat1, nodes1, feats1, T1);
    fill(k2, kps2, oct2, angle2, desc2, has2, ur2, n2, node2, start2, feat2, nodes2, feats2, T2);
    std::vector<std::pair<size_t, size_t> > vMatchedPairs;
    ORBmatcher matcher(0.6f, checkOri != 0);
    const int nm = matcher.SearchForTriangulation(&k1, &k2, vMatchedPairs, onlyStereo != 0, coarse != 0);
    for (size_t k = 0; k < vMatchedPairs.size(); k++) { pairs[2 * k] = (int)vMatchedPairs[k].first; pairs[2 * k + 1] = (int)vMatchedPairs[k].second; }
    return nm;
}

// Tracking::MonocularInitialization's call (Tracking.cc:2396 ff.): ORBmatcher(0.9, true).SearchForInitialization(mInitialFrame, mCurrentFrame,
// mvbPrevMatched, mvIniMatches, 100) (ORBmatcher.cc:648-766).  Frame 1: undistorted key points' octaves and angles + descriptors; frame 2: undistorted
// key points (x, y), octaves, angles, descriptors; fp = {mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv}; prev = n1 x 2
// search centres, updated in place as the reference does (:757-760).  -> matches12[n1]; returns nmatches.
int refcut_search_for_initialization(const int32_t* oct1, const float* angle1, const uint8_t* desc1, int n1, const float* kps2, const int32_t* oct2,
                                     const float* angle2, const uint8_t* desc2, int n2, const float* fp, float* prev, int windowSize, float nnratio,
                                     int checkOri, int32_t* matches12) {
    using namespace ORB_SLAM3;
    Frame* F1 = new Frame();
    Frame* F2 = new Frame();
    Frame::mnMinX = fp[0]; Frame::mnMaxX = fp[1]; Frame::mnMinY = fp[2]; Frame::mnMaxY = fp[3];
    Frame::mfGridElementWidthInv = fp[4]; Frame::mfGridElementHeightInv = fp[5];
    F1->N = n1; F1->Nleft = -1;
    F1->mvKeysUn.resize(n1);
    for (int i = 0; i < n1; i++) { F1->mvKeysUn[i].octave = oct1[i]; F1->mvKeysUn[i].angle = angle1[i]; }
    F1->mvKeys = F1->mvKeysUn;
    F1->mDescriptors = to_descriptors(desc1, n1);
    F2->N = n2; F2->Nleft = -1;
    F2->mvKeysUn.resize(n2);
    for (int i = 0; i < n2; i++) {
        F2->mvKeysUn[i].pt.x = kps2[2 * i]; F2->mvKeysUn[i].pt.y = kps2[2 * i + 1]; F2->mvKeysUn[i].octave = oct2[i]; F2->mvKeysUn[i].angle = angle2[i];
    }
    F2->mvKeys = F2->mvKeysUn;
    F2->AssignFeaturesToGrid();
    F2->mDescriptors = to_descriptors(desc2, n2);
    std::vector<cv::Point2f> vbPrevMatched(n1);
    for (int i = 0; i < n1; i++) { vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }
    std::vector<int> vnMatches12;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int nmatches = matcher.SearchForInitialization(*F1, *F2, vbPrevMatched, vnMatches12, windowSize);
    for (int i = 0; i < n1; i++) { matches12[i] = vnMatches12[i]; prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
    delete F1;
    delete F2;
    return nmatches;
}

// The motion-model call on fisheye-stereo frames (Nleft != -1 in both): the current frame's two key point sets (x, y, octave, angle), descriptors
// stacked, curState over nL + nR entries, Trl (12) = GetRelativePoseTrl(); the last frame's features are its left set followed by its right set
// (lastNL of them left): octave, angle, state, outlier flag, world position, descriptor.  -> matchOf[nL + nR]; returns nmatches.
int refcut_search_by_projection_motion_fisheye(const float* kpsL, const int32_t* octL, const float* angL, int nL, const float* kpsR, const int32_t* octR,
                                               const float* angR, int nR, const uint8_t* desc, const float* fp, const uint8_t* curState,
                                               const float* scaleFactors, int nlevels, const float* Tcw, const float* Trl, const float* cam4, int nLast,
                                               int lastNL, const int32_t* lastOct, const float* lastAngle, const uint8_t* lastState,
                                               const uint8_t* lastOutlier, const float* lastPos, const uint8_t* lastDesc, const float* Tlw, float th,
                                               int bMono, float nnratio, int checkOri, int32_t* matchOf) {
    using namespace ORB_SLAM3;
    Frame* Cf = new Frame();
    Frame* Lf = new Frame();
    Frame::mnMinX = fp[0]; Frame::mnMaxX = fp[1]; Frame::mnMinY = fp[2]; Frame::mnMaxY = fp[3];
    Frame::mfGridElementWidthInv = fp[4]; Frame::mfGridElementHeightInv = fp[5];
    GeometricCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    const int n = nL + nR;
    Cf->N = n; Cf->Nleft = nL; Cf->Nright = nR; Cf->mbf = fp[6]; Cf->mb = fp[7]; Cf->mpCamera = &cam;
    Cf->mTcw = Sophus::SE3f(Tcw, Tcw + 9);
    Cf->mTrl = Sophus::SE3f(Trl, Trl + 9);
    Cf->mvKeys.resize(nL); Cf->mvKeysRight.resize(nR);
    for (int i = 0; i < nL; i++) { Cf->mvKeys[i].pt.x = kpsL[2 * i]; Cf->mvKeys[i].pt.y = kpsL[2 * i + 1]; Cf->mvKeys[i].octave = octL[i]; Cf->mvKeys[i].angle = angL[i]; }
    for (int i = 0; i < nR; i++) {
        Cf->mvKeysRight[i].pt.x = kpsR[2 * i]; Cf->mvKeysRight[i].pt.y = kpsR[2 * i + 1]; Cf->mvKeysRight[i].octave = octR[i]; Cf->mvKeysRight[i].angle = angR[i];
    }
    Cf->AssignFeaturesToGrid();
    Cf->mDescriptors = to_descriptors(desc, n);
    Cf->mvuRight.assign(n, -1.0f);
    Cf->mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    MapPoint oldObs, oldNoObs;
    oldObs.nObs = 1; oldNoObs.nObs = 0;
    Cf->mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++) if (curState && curState[i]) Cf->mvpMapPoints[i] = curState[i] == 1 ? &oldObs : &oldNoObs;
    Lf->N = nLast; Lf->Nleft = lastNL; Lf->Nright = nLast - lastNL; Lf->mTcw = Sophus::SE3f(Tlw, Tlw + 9);
    Lf->mvKeys.resize(lastNL); Lf->mvKeysRight.resize(nLast - lastNL);
    std::vector<MapPoint> mps(nLast);
    Lf->mvpMapPoints.assign(nLast, nullptr);
    Lf->mvbOutlier.assign(nLast, false);
    for (int j = 0; j < nLast; j++) {
        cv::KeyPoint& kp = j < lastNL ? Lf->mvKeys[j] : Lf->mvKeysRight[j - lastNL];
        kp.octave = lastOct[j]; kp.angle = lastAngle[j];
        Lf->mvbOutlier[j] = lastOutlier[j] != 0;
        if (lastState[j]) {
            mps[j].nObs = lastState[j] == 1 ? 1 : 0;
            mps[j].mWorldPos = Eigen::Vector3f(lastPos[3 * j], lastPos[3 * j + 1], lastPos[3 * j + 2]);
            mps[j].mDescriptor = to_descriptors(lastDesc + (size_t)32 * j, 1);
            Lf->mvpMapPoints[j] = &mps[j];
        }
    }
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int nmatches = matcher.SearchByProjection(*Cf, *Lf, th, bMono != 0);
    for (int i = 0; i < n; i++) {
        MapPoint* p = Cf->mvpMapPoints[i];
        matchOf[i] = (p && p != &oldObs && p != &oldNoObs) ? (int)(p - mps.data()) : -1;
    }
    delete Cf;
    delete Lf;
    return nmatches;
}

// ORBmatcher(nnratio, checkOri).SearchByBoW(pKF, F, vpMapPointMatches) (Tracking::TrackReferenceKeyFrame / Relocalization) on a
// monocular key frame / frame pair.  Feature vectors as (node, start, feature list) triples in ascending node order (what
// refbow_transform returns); kfHasPoint[i] != 0: key-frame feature i holds a (good) map point.  -> matchOf[nF] = key-frame feature
// whose map point was matched to frame feature i (-1 none); returns nmatches.
int refcut_search_by_bow(const float* kfAngle, const uint8_t* kfDesc, const uint8_t* kfHasPoint, int nK, const int32_t* kfNode,
                         const int32_t* kfStart, const int32_t* kfFeat, int kfNodes, int kfFeats, const float* fAngle, const uint8_t* fDesc, int nF,
                         const int32_t* fNode, const int32_t* fStart, const int32_t* fFeat, int fNodes, int fFeats, float nnratio, int checkOri,
                         int32_t* matchOf) {
    using namespace ORB_SLAM3;
    KeyFrame kf;
    Frame* F = new Frame();
    std::vector<MapPoint> mps(nK);
    kf.mvKeysUn.resize(nK); kf.mvpMapPoints.assign(nK, nullptr);
    for (int i = 0; i < nK; i++) { kf.mvKeysUn[i].angle = kfAngle[i]; if (kfHasPoint[i]) kf.mvpMapPoints[i] = &mps[i]; }
    kf.mDescriptors = to_descriptors(kfDesc, nK);
    for (int g = 0; g < kfNodes; g++)
        for (int f = kfStart[g]; f < (g + 1 < kfNodes ? kfStart[g + 1] : kfFeats); f++) kf.mFeatVec.addFeature(kfNode[g], kfFeat[f]);
    F->N = nF; F->Nleft = -1;
    F->mvKeys.resize(nF);
    for (int i = 0; i < nF; i++) F->mvKeys[i].angle = fAngle[i];
    F->mDescriptors = to_descriptors(fDesc, nF);
    for (int g = 0; g < fNodes; g++)
        for (int f = fStart[g]; f < (g + 1 < fNodes ? fStart[g + 1] : fFeats); f++) F->mFeatVec.addFeature(fNode[g], fFeat[f]);
    std::vector<MapPoint*> matches;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int nm = matcher.SearchByBoW(&kf, *F, matches);
    for (int i = 0; i < nF; i++) matchOf[i] = matches[i] ? (int)(matches[i] - mps.data()) : -1;
    delete F;
    return nm;
}

// ORBmatcher(nnratio, checkOri).SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:765-905; LoopClosing's candidate check, LoopClosing.cc:1680)
// on two monocular key frames.  state1 / state2 per feature: 0 no map point / 1 good / 2 bad.  -> matchOf[n1] = feature of key frame 2 whose
// map point was matched to feature i of key frame 1 (-1 none); returns nmatches.
int refcut_search_by_bow_kf(const float* angle1, const uint8_t* desc1, const uint8_t* state1, int n1, const int32_t* node1, const int32_t* start1,
                            const int32_t* feat1, int nodes1, int feats1, const float* angle2, const uint8_t* desc2, const uint8_t* state2, int n2,
                            const int32_t* node2, const int32_t* start2, const int32_t* feat2, int nodes2, int feats2, float nnratio, int checkOri,
                            int32_t* matchOf) {
    using namespace ORB_SLAM3;
    KeyFrame k1, k2;
    std::vector<MapPoint> mps1(n1), mps2(n2);
    k1.mvKeysUn.resize(n1); k1.mvpMapPoints.assign(n1, nullptr);
    for (int i = 0; i < n1; i++) { k1.mvKeysUn[i].angle = angle1[i]; if (state1[i]) { mps1[i].mbBad = state1[i] == 2; k1.mvpMapPoints[i] = &mps1[i]; } }
    k1.mDescriptors = to_descriptors(desc1, n1);
    for (int g = 0; g < nodes1; g++)
        for (int f = start1[g]; f < (g + 1 < nodes1 ? start1[g + 1] : feats1); f++) k1.mFeatVec.addFeature(node1[g], feat1[f]);
    k2.mvKeysUn.resize(n2); k2.mvpMapPoints.assign(n2, nullptr);
    for (int i = 0; i < n2; i++) { k2.mvKeysUn[i].angle = angle2[i]; if (state2[i]) { mps2[i].mbBad = state2[i] == 2; k2.mvpMapPoints[i] = &mps2[i]; } }
    k2.mDescriptors = to_descriptors(desc2, n2);
    for (int g = 0; g < nodes2; g++)
        for (int f = start2[g]; f < (g + 1 < nodes2 ? start2[g + 1] : feats2); f++) k2.mFeatVec.addFeature(node2[g], feat2[f]);
    std::vector<MapPoint*> matches;
    ORBmatcher matcher(nnratio, checkOri != 0);
    const int nm = matcher.SearchByBoW(&k1, &k2, matches);
    for (int i = 0; i < n1; i++) matchOf[i] = matches[i] ? (int)(matches[i] - mps2.data()) : -1;
    return nm;
}

}  // extern "C"

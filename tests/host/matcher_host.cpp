// TEST HARNESS: the compiled GPU matcher adapter (orb_slam3_ros_b200/host/ORBmatcherGPU.cc) behind the SAME flat C entry points as the
// reference-cut glue of oracle/ref_cut_tu.cpp (refcut_search_by_projection, refcut_search_by_projection_motion, refcut_search_for_initialization), so that
// tests/test_gpu_matcher_host.py can feed both sides identical arrays and compare what they leave in mvpMapPoints.  Built as a shared
// library against tests/host/slam_stub + tests/cvstub (no Eigen / Sophus / OpenCV in this image).
#include <cstdint>
#include <cstring>
#include <set>
#include <vector>

#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "ORBmatcherGPU.h"

namespace ORB_SLAM3 {
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
}

using namespace ORB_SLAM3;

static cv::Mat to_descriptors(const uint8_t* d, int n) {
    cv::Mat m(n > 0 ? n : 1, 32, CV_8U);
    if (n) memcpy(m.data, d, (size_t)n * 32);
    return m;
}

static unsigned long g_frameId = 1;

extern "C" {

long gpuhost_rescans() { return ORBmatcherGPU::Instance().Rescans(); }

// same arguments and result as refcut_search_by_projection (oracle/ref_cut_tu.cpp)
int gpuhost_search_by_projection(const float* kps, const int32_t* oct, const uint8_t* train, int n, const float* grid4, const float* uRight,
                                 const uint8_t* hasPoint, const float* scaleFactors, int nlevels, const float* proj, const int32_t* level,
                                 const uint8_t* mpDesc, const uint8_t* inView, int nmp, float nnratio, float th, int32_t* matchOf) {
    Frame F;
    F.mnId = g_frameId++;
    Frame::mnMinX = grid4[0]; Frame::mnMinY = grid4[1]; Frame::mfGridElementWidthInv = grid4[2]; Frame::mfGridElementHeightInv = grid4[3];
    F.N = n; F.Nleft = -1;
    F.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) { F.mvKeysUn[i].pt.x = kps[2 * i]; F.mvKeysUn[i].pt.y = kps[2 * i + 1]; F.mvKeysUn[i].octave = oct[i]; }
    F.mDescriptors = to_descriptors(train, n);
    F.mvuRight.assign(n, -1.0f);
    if (uRight) for (int i = 0; i < n; i++) F.mvuRight[i] = uRight[i];
    F.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    MapPoint old;
    old.nObs = 1;
    F.mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++) if (hasPoint && hasPoint[i]) F.mvpMapPoints[i] = &old;
    std::vector<MapPoint> mps(nmp);
    std::vector<MapPoint*> vp(nmp);
    for (int j = 0; j < nmp; j++) {
        MapPoint& m = mps[j];
        m.mTrackProjX = proj[4 * j]; m.mTrackProjY = proj[4 * j + 1]; m.mTrackProjXR = proj[4 * j + 2]; m.mTrackViewCos = proj[4 * j + 3];
        m.mnTrackScaleLevel = level[j];
        m.mbTrackInView = inView[j] != 0;
        m.mDescriptor = to_descriptors(mpDesc + (size_t)32 * j, 1);
        m.nObs = 1;
        vp[j] = &m;
    }
    const int nmatches = ORBmatcherGPU::Instance().SearchByProjection(F, vp, th, false, 50.0f, nnratio);
    for (int i = 0; i < n; i++) {
        MapPoint* p = F.mvpMapPoints[i];
        matchOf[i] = (p && p != &old) ? (int)(p - mps.data()) : -1;
    }
    return nmatches;
}

// same arguments and result as refcut_search_by_projection_fisheye (oracle/ref_cut_tu.cpp)
int gpuhost_search_by_projection_fisheye(const float* kpsL, const int32_t* octL, int nL, const float* kpsR, const int32_t* octR, int nR, const uint8_t* desc,
                                         const float* fp, const int32_t* l2r, const int32_t* r2l, const uint8_t* hasPoint, const float* scaleFactors,
                                         int nlevels, const float* projL, const int32_t* levelL, const uint8_t* inViewL, const float* projR,
                                         const int32_t* levelR, const uint8_t* inViewR, const uint8_t* mpDesc, int nmp, float nnratio, float th,
                                         int32_t* matchOf) {
    Frame* F = new Frame();
    F->mnId = g_frameId++;
    Frame::mnMinX = fp[0]; Frame::mnMaxX = fp[1]; Frame::mnMinY = fp[2]; Frame::mnMaxY = fp[3];
    Frame::mfGridElementWidthInv = fp[4]; Frame::mfGridElementHeightInv = fp[5];
    const int n = nL + nR;
    F->N = n; F->Nleft = nL; F->Nright = nR;
    F->mvKeys.resize(nL); F->mvKeysRight.resize(nR);
    for (int i = 0; i < nL; i++) { F->mvKeys[i].pt.x = kpsL[2 * i]; F->mvKeys[i].pt.y = kpsL[2 * i + 1]; F->mvKeys[i].octave = octL[i]; }
    for (int i = 0; i < nR; i++) { F->mvKeysRight[i].pt.x = kpsR[2 * i]; F->mvKeysRight[i].pt.y = kpsR[2 * i + 1]; F->mvKeysRight[i].octave = octR[i]; }
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
    const int nmatches = ORBmatcherGPU::Instance().SearchByProjection(*F, vp, th, false, 50.0f, nnratio);
    for (int i = 0; i < n; i++) {
        MapPoint* p = F->mvpMapPoints[i];
        matchOf[i] = (p && p != &old) ? (int)(p - mps.data()) : -1;
    }
    delete F;
    return nmatches;
}

// same arguments and result as refcut_search_by_projection_motion (oracle/ref_cut_tu.cpp)
int gpuhost_search_by_projection_motion(const float* kps, const int32_t* oct, const float* angle, const uint8_t* desc, int n, const float* fp,
                                        const float* uRight, const uint8_t* curState, const float* scaleFactors, int nlevels, const float* Tcw,
                                        const float* cam4, int nLast, const int32_t* lastOct, const float* lastAngle, const uint8_t* lastState,
                                        const uint8_t* lastOutlier, const float* lastPos, const uint8_t* lastDesc, const float* Tlw, float th,
                                        int bMono, float nnratio, int checkOri, int32_t* matchOf) {
    (void)nnratio;
    Frame C, L;
    C.mnId = g_frameId++; L.mnId = g_frameId++;
    Frame::mnMinX = fp[0]; Frame::mnMaxX = fp[1]; Frame::mnMinY = fp[2]; Frame::mnMaxY = fp[3];
    Frame::mfGridElementWidthInv = fp[4]; Frame::mfGridElementHeightInv = fp[5];
    GeometricCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    C.N = n; C.Nleft = -1; C.mbf = fp[6]; C.mb = fp[7]; C.mpCamera = &cam;
    C.mTcw = Sophus::SE3f(Tcw, Tcw + 9);
    C.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) {
        C.mvKeysUn[i].pt.x = kps[2 * i]; C.mvKeysUn[i].pt.y = kps[2 * i + 1]; C.mvKeysUn[i].octave = oct[i]; C.mvKeysUn[i].angle = angle[i];
    }
    C.mvKeys = C.mvKeysUn;
    C.mDescriptors = to_descriptors(desc, n);
    C.mvuRight.assign(n, -1.0f);
    if (uRight) for (int i = 0; i < n; i++) C.mvuRight[i] = uRight[i];
    C.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    MapPoint oldObs, oldNoObs;
    oldObs.nObs = 1; oldNoObs.nObs = 0;
    C.mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++) if (curState && curState[i]) C.mvpMapPoints[i] = curState[i] == 1 ? &oldObs : &oldNoObs;
    L.N = nLast; L.Nleft = -1; L.mTcw = Sophus::SE3f(Tlw, Tlw + 9);
    L.mvKeysUn.resize(nLast);
    std::vector<MapPoint> mps(nLast);
    L.mvpMapPoints.assign(nLast, nullptr);
    L.mvbOutlier.assign(nLast, false);
    for (int j = 0; j < nLast; j++) {
        L.mvKeysUn[j].octave = lastOct[j]; L.mvKeysUn[j].angle = lastAngle[j];
        L.mvbOutlier[j] = lastOutlier[j] != 0;
        if (lastState[j]) {
            mps[j].nObs = lastState[j] == 1 ? 1 : 0;
            mps[j].mWorldPos = Eigen::Vector3f(lastPos[3 * j], lastPos[3 * j + 1], lastPos[3 * j + 2]);
            mps[j].mDescriptor = to_descriptors(lastDesc + (size_t)32 * j, 1);
            L.mvpMapPoints[j] = &mps[j];
        }
    }
    L.mvKeys = L.mvKeysUn;
    const int nmatches = ORBmatcherGPU::Instance().SearchByProjection(C, L, th, bMono != 0, checkOri != 0);
    for (int i = 0; i < n; i++) {
        MapPoint* p = C.mvpMapPoints[i];
        matchOf[i] = (p && p != &oldObs && p != &oldNoObs) ? (int)(p - mps.data()) : -1;
    }
    return nmatches;
}

// same arguments and result as refcut_search_for_initialization (oracle/ref_cut_tu.cpp)
int gpuhost_search_for_initialization(const int32_t* oct1, const float* angle1, const uint8_t* desc1, int n1, const float* kps2, const int32_t* oct2,
                                      const float* angle2, const uint8_t* desc2, int n2, const float* fp, float* prev, int windowSize, float nnratio,
                                      int checkOri, int32_t* matches12) {
    Frame* F1 = new Frame();      // (the grid makes a Frame too large for the stack)
    Frame* F2 = new Frame();
    F1->mnId = g_frameId++; F2->mnId = g_frameId++;
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
    F2->mvuRight.assign(n2, -1.0f);
    F2->mvpMapPoints.assign(n2, nullptr);
    std::vector<cv::Point2f> vbPrevMatched(n1);
    for (int i = 0; i < n1; i++) { vbPrevMatched[i].x = prev[2 * i]; vbPrevMatched[i].y = prev[2 * i + 1]; }
    std::vector<int> vnMatches12;
    const int nmatches = ORBmatcherGPU::Instance().SearchForInitialization(*F1, *F2, vbPrevMatched, vnMatches12, windowSize, nnratio, checkOri != 0);
    for (int i = 0; i < n1; i++) { matches12[i] = vnMatches12[i]; prev[2 * i] = vbPrevMatched[i].x; prev[2 * i + 1] = vbPrevMatched[i].y; }
    delete F1;
    delete F2;
    return nmatches;
}

// same arguments and result as refcut_search_by_projection_motion_fisheye (oracle/ref_cut_tu.cpp)
int gpuhost_search_by_projection_motion_fisheye(const float* kpsL, const int32_t* octL, const float* angL, int nL, const float* kpsR, const int32_t* octR,
                                                const float* angR, int nR, const uint8_t* desc, const float* fp, const uint8_t* curState,
                                                const float* scaleFactors, int nlevels, const float* Tcw, const float* Trl, const float* cam4, int nLast,
                                                int lastNL, const int32_t* lastOct, const float* lastAngle, const uint8_t* lastState,
                                                const uint8_t* lastOutlier, const float* lastPos, const uint8_t* lastDesc, const float* Tlw, float th,
                                                int bMono, float nnratio, int checkOri, int32_t* matchOf) {
    (void)nnratio;
    Frame* Cf = new Frame();
    Frame* Lf = new Frame();
    Cf->mnId = g_frameId++; Lf->mnId = g_frameId++;
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
    const int nmatches = ORBmatcherGPU::Instance().SearchByProjection(*Cf, *Lf, th, bMono != 0, checkOri != 0);
    for (int i = 0; i < n; i++) {
        MapPoint* p = Cf->mvpMapPoints[i];
        matchOf[i] = (p && p != &oldObs && p != &oldNoObs) ? (int)(p - mps.data()) : -1;
    }
    delete Cf;
    delete Lf;
    return nmatches;
}

// same arguments and result as refcut_search_by_projection_reloc (oracle/ref_cut_tu.cpp)
int gpuhost_search_by_projection_reloc(const float* kps, const int32_t* oct, const float* angle, const uint8_t* desc, int n, const float* fp,
                                       const uint8_t* curHolds, const float* scaleFactors, int nlevels, const float* Tcw, const float* cam4, int nK,
                                       const float* kfAngle, const uint8_t* kfState, const float* kfPos, const uint8_t* kfDesc, const float* kfMinDist,
                                       const float* kfMaxDist, float th, int ORBdist, float nnratio, int checkOri, int32_t* matchOf) {
    (void)nnratio;
    Frame* Cf = new Frame();
    Cf->mnId = g_frameId++;
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
    const int nmatches = ORBmatcherGPU::Instance().SearchByProjection(*Cf, &kf, sFound, th, ORBdist, checkOri != 0);
    for (int i = 0; i < n; i++) {
        MapPoint* p = Cf->mvpMapPoints[i];
        matchOf[i] = (p && p != &held) ? (int)(p - mps.data()) : -1;
    }
    delete Cf;
    return nmatches;
}

// same arguments and result as refcut_search_by_bow (oracle/ref_cut_tu.cpp)
int gpuhost_search_by_bow(const float* kfAngle, const uint8_t* kfDesc, const uint8_t* kfHasPoint, int nK, const int32_t* kfNode, const int32_t* kfStart,
                          const int32_t* kfFeat, int kfNodes, int kfFeats, const float* fAngle, const uint8_t* fDesc, int nF, const int32_t* fNode,
                          const int32_t* fStart, const int32_t* fFeat, int fNodes, int fFeats, float nnratio, int checkOri, int32_t* matchOf) {
    KeyFrame kf;
    Frame* F = new Frame();
    F->mnId = g_frameId++;
    std::vector<MapPoint> mps(nK);
    kf.mvKeysUn.resize(nK); kf.mvpMapPoints.assign(nK, nullptr);
    for (int i = 0; i < nK; i++) { kf.mvKeysUn[i].angle = kfAngle[i]; if (kfHasPoint[i]) kf.mvpMapPoints[i] = &mps[i]; }
    kf.mDescriptors = to_descriptors(kfDesc, nK);
    for (int g = 0; g < kfNodes; g++)
        for (int f = kfStart[g]; f < (g + 1 < kfNodes ? kfStart[g + 1] : kfFeats); f++) kf.mFeatVec[(unsigned)kfNode[g]].push_back((unsigned)kfFeat[f]);
    F->N = nF; F->Nleft = -1;
    F->mvKeys.resize(nF);
    for (int i = 0; i < nF; i++) F->mvKeys[i].angle = fAngle[i];
    F->mvKeysUn = F->mvKeys;
    F->mDescriptors = to_descriptors(fDesc, nF);
    for (int g = 0; g < fNodes; g++)
        for (int f = fStart[g]; f < (g + 1 < fNodes ? fStart[g + 1] : fFeats); f++) F->mFeatVec[(unsigned)fNode[g]].push_back((unsigned)fFeat[f]);
    std::vector<MapPoint*> matches;
    const int nm = ORBmatcherGPU::Instance().SearchByBoW(&kf, *F, matches, nnratio, checkOri != 0);
    for (int i = 0; i < nF; i++) matchOf[i] = matches[i] ? (int)(matches[i] - mps.data()) : -1;
    delete F;
    return nm;
}

// same arguments and result as refcut_search_by_projection_sim3 (oracle/ref_cut_tu.cpp)
static int gpuhost_sim3(int withKFs, const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held,
                        const float* scaleFactors, int nlevels, const float* sim3, const float* cam4, int nP, const uint8_t* pState, const float* pPos,
                        const float* pNormal, const uint8_t* pDesc, const float* pMinDist, const float* pMaxDist, int th, float ratioHamming,
                        int32_t* matchOf, int32_t* matchKF);
int gpuhost_search_by_projection_sim3(const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held,
                                      const float* scaleFactors, int nlevels, const float* sim3, const float* cam4, int nP, const uint8_t* pState,
                                      const float* pPos, const float* pNormal, const uint8_t* pDesc, const float* pMinDist, const float* pMaxDist, int th,
                                      float ratioHamming, int32_t* matchOf) {
    return gpuhost_sim3(0, kps, oct, desc, n, fp, held, scaleFactors, nlevels, sim3, cam4, nP, pState, pPos, pNormal, pDesc, pMinDist, pMaxDist, th,
                        ratioHamming, matchOf, nullptr);
}
// same arguments and result as refcut_search_by_projection_sim3_kfs
int gpuhost_search_by_projection_sim3_kfs(const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held,
                                          const float* scaleFactors, int nlevels, const float* sim3, const float* cam4, int nP, const uint8_t* pState,
                                          const float* pPos, const float* pNormal, const uint8_t* pDesc, const float* pMinDist, const float* pMaxDist,
                                          int th, float ratioHamming, int32_t* matchOf, int32_t* matchKF) {
    return gpuhost_sim3(1, kps, oct, desc, n, fp, held, scaleFactors, nlevels, sim3, cam4, nP, pState, pPos, pNormal, pDesc, pMinDist, pMaxDist, th,
                        ratioHamming, matchOf, matchKF);
}
static int gpuhost_sim3(int withKFs, const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held,
                        const float* scaleFactors, int nlevels, const float* sim3, const float* cam4, int nP, const uint8_t* pState, const float* pPos,
                        const float* pNormal, const uint8_t* pDesc, const float* pMinDist, const float* pMaxDist, int th, float ratioHamming,
                        int32_t* matchOf, int32_t* matchKF) {
    KeyFrame kf;
    GeometricCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    kf.fx = cam4[0]; kf.fy = cam4[1]; kf.cx = cam4[2]; kf.cy = cam4[3];
    kf.mpCamera = &cam;
    kf.NLeft = -1;
    kf.mnMinX = (int)fp[0]; kf.mnMaxX = (int)fp[1]; kf.mnMinY = (int)fp[2]; kf.mnMaxY = (int)fp[3];
    kf.mfGridElementWidthInv = fp[4]; kf.mfGridElementHeightInv = fp[5];
    kf.mnScaleLevels = (int)fp[8]; kf.mfLogScaleFactor = fp[9];
    kf.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    kf.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) { kf.mvKeysUn[i].pt.x = kps[2 * i]; kf.mvKeysUn[i].pt.y = kps[2 * i + 1]; kf.mvKeysUn[i].octave = oct[i]; }
    kf.mDescriptors = to_descriptors(desc, n);
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
    int nmatches;
    if (withKFs) {
        KeyFrame pool[7];
        std::vector<KeyFrame*> vpPointsKFs(nP), vpMatchedKF(n, nullptr);
        for (int j = 0; j < nP; j++) vpPointsKFs[j] = &pool[j % 7];
        nmatches = ORBmatcherGPU::Instance().SearchByProjection(&kf, Scw, vpPoints, vpPointsKFs, vpMatched, vpMatchedKF, th, ratioHamming);
        for (int i = 0; i < n; i++) matchKF[i] = vpMatchedKF[i] ? (int)(vpMatchedKF[i] - pool) : -1;
    } else {
        nmatches = ORBmatcherGPU::Instance().SearchByProjection(&kf, Scw, vpPoints, vpMatched, th, ratioHamming);
    }
    for (int i = 0; i < n; i++) matchOf[i] = (vpMatched[i] && vpMatched[i] != &other) ? (int)(vpMatched[i] - mps.data()) : -1;
    return nmatches;
}

// same arguments and result as refcut_search_for_triangulation (oracle/ref_cut_tu.cpp)
int gpuhost_search_for_triangulation(const float* kps1, const int32_t* oct1, const float* angle1, const uint8_t* desc1, const uint8_t* has1, const float* ur1,
                                     int n1, const int32_t* node1, const int32_t* start1, const int32_t* feat1, int nodes1, int feats1, const float* T1,
                                     const float* kps2, const int32_t* oct2, const float* angle2, const uint8_t* desc2, const uint8_t* has2, const float* ur2,
                                     int n2, const int32_t* node2, const int32_t* start2, const int32_t* feat2, int nodes2, int feats2, const float* T2,
                                     const float* sigma2, const float* scaleFactors, int nlevels, const float* cam4, int onlyStereo, int coarse, int checkOri,
                                     int32_t* pairs) {
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
            for (int f = start[g]; f < (g + 1 < nodes ? start[g + 1] : feats); f++) k.mFeatVec[(unsigned)node[g]].push_back((unsigned)feat[f]);
    };
    fill(k1, kps1, oct1, angle1, desc1, has1, ur1, n1, node1, start1, feat1, nodes1, feats1, T1);
    fill(k2, kps2, oct2, angle2, desc2, has2, ur2, n2, node2, start2, feat2, nodes2, feats2, T2);
    std::vector<std::pair<size_t, size_t> > vMatchedPairs;
    const int nm = ORBmatcherGPU::Instance().SearchForTriangulation(&k1, &k2, vMatchedPairs, onlyStereo != 0, coarse != 0, checkOri != 0);
    for (size_t k = 0; k < vMatchedPairs.size(); k++) { pairs[2 * k] = (int)vMatchedPairs[k].first; pairs[2 * k + 1] = (int)vMatchedPairs[k].second; }
    return nm;
}

// same arguments and result as refcut_fuse_kf (oracle/ref_cut_tu.cpp)
int gpuhost_fuse_kf(const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held, const int32_t* heldObs,
                    const float* uRight, const float* invSigma2, const float* scaleFactors, int nlevels, const float* Tcw, const float* cam4, int nP,
                    const uint8_t* pState, const int32_t* pObs, const float* pPos, const float* pNormal, const uint8_t* pDesc, const float* pMinDist,
                    const float* pMaxDist, float th, int32_t* kpHolds, uint8_t* ownBad, uint8_t* ptBad) {
    KeyFrame kf;
    GeometricCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    kf.fx = cam4[0]; kf.fy = cam4[1]; kf.cx = cam4[2]; kf.cy = cam4[3];
    kf.mpCamera = &cam;
    kf.NLeft = -1;
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
    kf.AssignFeaturesToGrid();
    std::vector<MapPoint> own(n);
    kf.mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; i++)
        if (held && held[i]) {
            own[i].mbBad = held[i] == 2;
            own[i].nObs = heldObs[i];
            own[i].mObservations[&kf] = i;
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
    const int nFused = ORBmatcherGPU::Instance().Fuse(&kf, vpPoints, th);
    for (int i = 0; i < n; i++) {
        MapPoint* p = kf.mvpMapPoints[i];
        kpHolds[i] = !p ? -1 : (p >= own.data() && p < own.data() + n) ? (int)(p - own.data()) : 1000000 + (int)(p - mps.data());
        ownBad[i] = own[i].mbBad;
    }
    for (int j = 0; j < nP; j++) ptBad[j] = mps[j].mbBad;
    return nFused;
}

// same arguments and result as refcut_fuse_sim3 (oracle/ref_cut_tu.cpp)
int gpuhost_fuse_sim3(const float* kps, const int32_t* oct, const uint8_t* desc, int n, const float* fp, const uint8_t* held, const float* scaleFactors,
                      int nlevels, const float* sim3, const float* cam4, int nP, const uint8_t* pState, const float* pPos, const float* pNormal,
                      const uint8_t* pDesc, const float* pMinDist, const float* pMaxDist, float th, int32_t* replaceOf, int32_t* addedAt) {
    KeyFrame kf;
    GeometricCamera cam;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    kf.fx = cam4[0]; kf.fy = cam4[1]; kf.cx = cam4[2]; kf.cy = cam4[3];
    kf.mpCamera = &cam;
    kf.NLeft = -1;
    kf.mnMinX = (int)fp[0]; kf.mnMaxX = (int)fp[1]; kf.mnMinY = (int)fp[2]; kf.mnMaxY = (int)fp[3];
    kf.mfGridElementWidthInv = fp[4]; kf.mfGridElementHeightInv = fp[5];
    kf.mnScaleLevels = (int)fp[8]; kf.mfLogScaleFactor = fp[9];
    kf.mvScaleFactors.assign(scaleFactors, scaleFactors + nlevels);
    kf.mvKeysUn.resize(n);
    for (int i = 0; i < n; i++) { kf.mvKeysUn[i].pt.x = kps[2 * i]; kf.mvKeysUn[i].pt.y = kps[2 * i + 1]; kf.mvKeysUn[i].octave = oct[i]; }
    kf.mDescriptors = to_descriptors(desc, n);
    std::vector<MapPoint> own(n);
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
    const int nFused = ORBmatcherGPU::Instance().Fuse(&kf, Scw, vpPoints, th, vpReplace);
    for (int j = 0; j < nP; j++) {
        // (a point can be replaced by one that an earlier iteration of this call added to the key frame: reported as 1000000 + its index)
        replaceOf[j] = !vpReplace[j] ? -1 : (vpReplace[j] >= own.data() && vpReplace[j] < own.data() + n) ? (int)(vpReplace[j] - own.data())
                                                                                                    : 1000000 + (int)(vpReplace[j] - mps.data());
        addedAt[j] = mps[j].mObservations.count(&kf) ? mps[j].mObservations[&kf] : -1;
    }
    return nFused;
}

// same arguments and result as refcut_search_by_bow_kf (oracle/ref_cut_tu.cpp)
int gpuhost_search_by_bow_kf(const float* angle1, const uint8_t* desc1, const uint8_t* state1, int n1, const int32_t* node1, const int32_t* start1,
                             const int32_t* feat1, int nodes1, int feats1, const float* angle2, const uint8_t* desc2, const uint8_t* state2, int n2,
                             const int32_t* node2, const int32_t* start2, const int32_t* feat2, int nodes2, int feats2, float nnratio, int checkOri,
                             int32_t* matchOf) {
    KeyFrame k1, k2;
    std::vector<MapPoint> mps1(n1), mps2(n2);
    k1.mvKeysUn.resize(n1); k1.mvpMapPoints.assign(n1, nullptr);
    for (int i = 0; i < n1; i++) { k1.mvKeysUn[i].angle = angle1[i]; if (state1[i]) { mps1[i].mbBad = state1[i] == 2; k1.mvpMapPoints[i] = &mps1[i]; } }
    k1.mDescriptors = to_descriptors(desc1, n1);
    for (int g = 0; g < nodes1; g++)
        for (int f = start1[g]; f < (g + 1 < nodes1 ? start1[g + 1] : feats1); f++) k1.mFeatVec[(unsigned)node1[g]].push_back((unsigned)feat1[f]);
    k2.mvKeysUn.resize(n2); k2.mvpMapPoints.assign(n2, nullptr);
    for (int i = 0; i < n2; i++) { k2.mvKeysUn[i].angle = angle2[i]; if (state2[i]) { mps2[i].mbBad = state2[i] == 2; k2.mvpMapPoints[i] = &mps2[i]; } }
    k2.mDescriptors = to_descriptors(desc2, n2);
    for (int g = 0; g < nodes2; g++)
        for (int f = start2[g]; f < (g + 1 < nodes2 ? start2[g + 1] : feats2); f++) k2.mFeatVec[(unsigned)node2[g]].push_back((unsigned)feat2[f]);
    std::vector<MapPoint*> matches;
    const int nm = ORBmatcherGPU::Instance().SearchByBoW(&k1, &k2, matches, nnratio, checkOri != 0);
    for (int i = 0; i < n1; i++) matchOf[i] = matches[i] ? (int)(matches[i] - mps2.data()) : -1;
    return nm;
}

}  // extern "C"

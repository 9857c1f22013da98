"""GPU: the compiled matcher replacements of the host adapter (orb_slam3_ros_b200/host/ORBmatcherGPU.cc: the two per-frame
ORBmatcher::SearchByProjection overloads on top of orbb_search_area_topk, frames uploaded once) against the REFERENCE's own function
bodies (cut out of ORBmatcher.cc at build time and compiled into oracle/_ref, see oracle/ref_cut_tu.cpp) on identical inputs: what the
two leave in Frame::mvpMapPoints and the match counts must be identical.  Plus the device-resident frame views of the C ABI against
the oracle port."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import port, ref
from orb_slam3_ros_b200 import capi, synth
from orb_slam3_ros_b200.extractor import ORBextractor
from orb_slam3_ros_b200.matcher import ORBmatcher
from scenes import bow_scene, fisheye_local_points_scene, fisheye_motion_scene, fuse_scene, init_scene, local_points_scene, motion_scene, reloc_scene, sim3_scene, triangulation_scene

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
SO = ROOT / "tests" / "models" / "_build" / "libmatcher_host.so"


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def declare(lib):
    """argument types of the harness entry points (tests/host/matcher_host.cpp), shared with tests/test_matcher_host_cpu.py"""
    lib.gpuhost_rescans.restype = C.c_long
    lib.gpuhost_search_by_projection.restype = C.c_int
    lib.gpuhost_search_by_projection.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
    lib.gpuhost_search_by_projection_fisheye.restype = C.c_int
    lib.gpuhost_search_by_projection_fisheye.argtypes = ref.FISHEYE_ARGTYPES
    lib.gpuhost_search_by_projection_motion_fisheye.restype = C.c_int
    lib.gpuhost_search_by_projection_motion_fisheye.argtypes = ref.MOTION_FISHEYE_ARGTYPES
    lib.gpuhost_search_by_projection_motion.restype = C.c_int
    lib.gpuhost_search_by_projection_motion.argtypes = [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int] + \
        [C.c_void_p] * 7 + [C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p]
    lib.gpuhost_search_by_projection_reloc.restype = C.c_int
    lib.gpuhost_search_by_projection_reloc.argtypes = ref.RELOC_ARGTYPES
    lib.gpuhost_search_by_projection_sim3.restype = C.c_int
    lib.gpuhost_search_by_projection_sim3.argtypes = ref.SIM3_ARGTYPES
    lib.gpuhost_search_by_projection_sim3_kfs.restype = C.c_int
    lib.gpuhost_search_by_projection_sim3_kfs.argtypes = ref.SIM3_ARGTYPES + [C.c_void_p]
    lib.gpuhost_search_for_triangulation.restype = C.c_int
    lib.gpuhost_search_for_triangulation.argtypes = ref.TRI_ARGTYPES
    lib.gpuhost_fuse_kf.restype = C.c_int
    lib.gpuhost_fuse_kf.argtypes = ref.FUSE_KF_ARGTYPES
    lib.gpuhost_fuse_sim3.restype = C.c_int
    lib.gpuhost_fuse_sim3.argtypes = ref.FUSE_SIM3_ARGTYPES
    lib.gpuhost_search_by_bow_kf.restype = C.c_int
    lib.gpuhost_search_by_bow_kf.argtypes = ref.BOW_KF_ARGTYPES
    lib.gpuhost_search_by_bow.restype = C.c_int
    lib.gpuhost_search_by_bow.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 2 + [C.c_int] + \
        [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
    lib.gpuhost_search_for_initialization.restype = C.c_int
    lib.gpuhost_search_for_initialization.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 2 + \
        [C.c_int, C.c_float, C.c_int, C.c_void_p]


@pytest.fixture(scope="module")
def host():
    """the adapter + harness compiled as C++14 against the stand-in headers (tests/host/slam_stub, tests/cvstub, oracle/cvshim/mini_geom.hpp)"""
    from orb_slam3_ros_b200 import build
    build.build_library()
    SO.parent.mkdir(exist_ok=True)
    pkg = ROOT / "orb_slam3_ros_b200"
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", f"-I{ROOT / 'tests' / 'cvstub'}",
                           f"-I{ROOT / 'tests' / 'host' / 'slam_stub'}", f"-I{ROOT / 'oracle' / 'cvshim'}", f"-I{ROOT / 'include'}", f"-I{pkg / 'host'}",
                           str(ROOT / "tests" / "host" / "matcher_host.cpp"), str(pkg / "host" / "ORBmatcherGPU.cc"), f"-L{pkg}", "-lorbb200",
                           f"-Wl,-rpath,{pkg}", "-o", str(SO)])
    lib = C.CDLL(str(SO))
    declare(lib)
    return lib


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")
@pytest.mark.parametrize("th,orb_dist,check,seed", [(10.0, 100, True, 5), (3.0, 64, True, 5), (10.0, 100, False, 6)])
def test_relocalization_search_equals_reference(host, th, orb_dist, check, seed):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1889-2010, Tracking::Relocalization with
    (10, 100) and (3, 64), Tracking.cc:3765 / :3779)"""
    cur, kf = reloc_scene(seed)
    nm_ref, match_ref = ref.search_by_projection_reloc(cur, kf, th, orb_dist, 0.9, check)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    k, o, a, d = f32(cur["kps_xy"]), i32(cur["octaves"]), f32(cur["angles"]), u8(cur["desc"])
    holds, fp, sf, tcw, cam = u8(cur["holds"]), f32(cur["fp"]), f32(cur["scale_factors"]), f32(cur["Tcw"]), f32(cur["cam4"])
    ka, ks, kp, kd, kmin, kmax = f32(kf["angles"]), u8(kf["state"]), f32(kf["pos"]), u8(kf["desc"]), f32(kf["min_dist"]), f32(kf["max_dist"])
    match = np.full(len(k), -1, np.int32)
    nm = host.gpuhost_search_by_projection_reloc(_p(k), _p(o), _p(a), _p(d), len(k), _p(fp), _p(holds), _p(sf), len(sf), _p(tcw), _p(cam), len(ka), _p(ka),
                                                 _p(ks), _p(kp), _p(kd), _p(kmin), _p(kmax), th, orb_dist, 0.9, int(check), _p(match))
    assert nm == nm_ref and np.array_equal(match, match_ref)
    assert nm_ref > 30


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")
@pytest.mark.parametrize("levelsup,nnratio,check", [(2, 0.7, True), (3, 0.75, True), (2, 0.9, False)])
def test_search_by_bow_equals_reference(host, levelsup, nnratio, check):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) (ORBmatcher.cc:223-421, Tracking::TrackReferenceKeyFrame / Relocalization):
    the map points that the compiled GPU replacement leaves in vpMapPointMatches against the reference's own body over the vendored
    DBoW2::FeatureVector, on the same key frame / frame pair"""
    sc = bow_scene(levelsup=levelsup)
    nm_ref, match_ref = ref.search_by_bow(sc["ang_k"], sc["dk"], sc["has_point"], sc["fv_k"], sc["ang_f"], sc["df"], sc["fv_f"], nnratio, check)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    ka, kd, hp, fa, fd = f32(sc["ang_k"]), u8(sc["dk"]), u8(sc["has_point"]), f32(sc["ang_f"]), u8(sc["df"])
    kn, ks, kfe = (i32(a) for a in sc["fv_k"])
    fn, fs, ffe = (i32(a) for a in sc["fv_f"])
    match = np.full(len(fa), -1, np.int32)
    r0 = host.gpuhost_rescans()
    nm = host.gpuhost_search_by_bow(_p(ka), _p(kd), _p(hp), len(ka), _p(kn), _p(ks), _p(kfe), len(kn), len(kfe), _p(fa), _p(fd), len(fa), _p(fn), _p(fs),
                                    _p(ffe), len(fn), len(ffe), nnratio, int(check), _p(match))
    assert nm == nm_ref and np.array_equal(match, match_ref)
    assert nm_ref > 100 and host.gpuhost_rescans() > r0      # matches and in-call collisions both occurred


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")
@pytest.mark.parametrize("seed,th,ratio", [(5, 5, 1.0), (5, 3, 1.5), (6, 8, 1.5)])
def test_sim3_projection_search_equals_reference(host, seed, th, ratio):
    """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (ORBmatcher.cc:427-530; LoopClosing.cc:1795 / :1982) with
    the reference's own KeyFrame::GetFeaturesInArea on the oracle side"""
    k, pts, sim3 = sim3_scene(seed)
    nm_ref, match_ref = ref.search_by_projection_sim3(k, pts, sim3, th, ratio)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    kx, o, d, held, fp, sf, cam = f32(k["kps_xy"]), i32(k["octaves"]), u8(k["desc"]), u8(k["held"]), f32(k["fp"]), f32(k["scale_factors"]), f32(k["cam4"])
    ps, pp, pn, pd, pmin, pmax = u8(pts["state"]), f32(pts["pos"]), f32(pts["normal"]), u8(pts["desc"]), f32(pts["min_dist"]), f32(pts["max_dist"])
    s3 = f32(sim3)
    match = np.full(len(kx), -1, np.int32)
    nm = host.gpuhost_search_by_projection_sim3(_p(kx), _p(o), _p(d), len(kx), _p(fp), _p(held), _p(sf), len(sf), _p(s3), _p(cam), len(ps), _p(ps), _p(pp),
                                                _p(pn), _p(pd), _p(pmin), _p(pmax), th, ratio, _p(match))
    assert nm == nm_ref and np.array_equal(match, match_ref)
    assert nm_ref > 100
    # the overload that also returns the source key frames (ORBmatcher.cc:532-646, LoopClosing.cc:1773: th 8, ratio 1.5); it projects by hand
    nm_ref2, match_ref2, kf_ref2 = ref.search_by_projection_sim3(k, pts, sim3, th, ratio, with_kfs=True)
    match2, kf2 = np.full(len(kx), -1, np.int32), np.full(len(kx), -1, np.int32)
    nm2 = host.gpuhost_search_by_projection_sim3_kfs(_p(kx), _p(o), _p(d), len(kx), _p(fp), _p(held), _p(sf), len(sf), _p(s3), _p(cam), len(ps), _p(ps), _p(pp),
                                                     _p(pn), _p(pd), _p(pmin), _p(pmax), th, ratio, _p(match2), _p(kf2))
    assert nm2 == nm_ref2 and np.array_equal(match2, match_ref2) and np.array_equal(kf2, kf_ref2)
    assert np.array_equal(kf_ref2 >= 0, match_ref2 >= 0) and np.array_equal(kf_ref2[match_ref2 >= 0], match_ref2[match_ref2 >= 0] % 7)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")
@pytest.mark.parametrize("stereo,only_stereo,coarse,check,levelsup,ties", [(False, False, False, False, 2, False), (True, False, False, True, 2, False),
                                                                           (True, True, False, False, 3, False), (False, False, True, True, 2, False),
                                                                           (False, False, True, False, 2, True), (False, False, False, False, 3, True)])
def test_search_for_triangulation_equals_reference(host, stereo, only_stereo, coarse, check, levelsup, ties):
    """ORBmatcher::SearchForTriangulation (ORBmatcher.cc:906-1146; LocalMapping::CreateNewMapPoints): the matched pairs"""
    k1, k2, common = triangulation_scene(levelsup=levelsup, stereo=stereo, ties=ties)
    nm_ref, pairs_ref = ref.search_for_triangulation(k1, k2, common, only_stereo, coarse, check)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    keep, args = [], []
    for k in (k1, k2):
        arrs = [f32(k["kps_xy"]), i32(k["octaves"]), f32(k["angles"]), u8(k["desc"]), u8(k["has_point"]), None if k["u_right"] is None else f32(k["u_right"])]
        fv = [i32(a) for a in k["fv"]]
        t = f32(k["Tcw"])
        keep += arrs + fv + [t]
        args += [_p(a) for a in arrs] + [len(arrs[0])] + [_p(a) for a in fv] + [len(fv[0]), len(fv[2]), _p(t)]
    sg, sf, cam = f32(common["sigma2"]), f32(common["scale_factors"]), f32(common["cam4"])
    pairs = np.full((len(keep[0]), 2), -1, np.int32)
    nm = host.gpuhost_search_for_triangulation(*args, _p(sg), _p(sf), len(sf), _p(cam), int(only_stereo), int(coarse), int(check), _p(pairs))
    assert nm == nm_ref and np.array_equal(pairs[:max(nm, 0)], pairs_ref)
    assert nm_ref > (15 if only_stereo else 60)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")
@pytest.mark.parametrize("seed,th,stereo", [(5, 3.0, False), (6, 3.0, True), (7, 6.0, True), (8, 12.0, True)])
def test_fuse_equals_reference(host, seed, th, stereo):
    """ORBmatcher::Fuse(pKF, vpMapPoints, th) (ORBmatcher.cc:1148-1338; LocalMapping::SearchInNeighbors): what every key point holds after the
    call, which key-frame points and which list points went bad (replaced), and the count"""
    k, pts = fuse_scene(seed, stereo)
    nf_ref, holds_ref, ob_ref, pb_ref = ref.fuse_kf(k, pts, th)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    kx, o, d, held, hobs = f32(k["kps_xy"]), i32(k["octaves"]), u8(k["desc"]), u8(k["held"]), i32(k["held_obs"])
    ur = None if k["u_right"] is None else f32(k["u_right"])
    isg, fp, sf, tcw, cam = f32(k["inv_sigma2"]), f32(k["fp"]), f32(k["scale_factors"]), f32(k["Tcw"]), f32(k["cam4"])
    ps, pobs, pp, pn, pd, pmin, pmax = u8(pts["state"]), i32(pts["obs"]), f32(pts["pos"]), f32(pts["normal"]), u8(pts["desc"]), f32(pts["min_dist"]), f32(pts["max_dist"])
    holds, ob, pb = np.full(len(kx), -1, np.int32), np.zeros(len(kx), np.uint8), np.zeros(len(ps), np.uint8)
    r0 = host.gpuhost_rescans()
    nf = host.gpuhost_fuse_kf(_p(kx), _p(o), _p(d), len(kx), _p(fp), _p(held), _p(hobs), _p(ur), _p(isg), _p(sf), len(sf), _p(tcw), _p(cam), len(ps), _p(ps),
                              _p(pobs), _p(pp), _p(pn), _p(pd), _p(pmin), _p(pmax), th, _p(holds), _p(ob), _p(pb))
    assert nf == nf_ref and np.array_equal(holds, holds_ref) and np.array_equal(ob, ob_ref) and np.array_equal(pb, pb_ref)
    went_bad_own = int((ob_ref > 0).sum() - (held == 2).sum())
    went_bad_pts = int((pb_ref > 0).sum() - (ps == 2).sum())
    assert nf_ref > 60 and went_bad_own > 5 and went_bad_pts > 5 and (holds_ref >= 1000000).sum() > 20      # both replacement directions, new observations
    if th >= 12:
        assert host.gpuhost_rescans() > r0      # wide windows: some points lose all eight candidates to the chi-square test and walk the window on the host


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")
@pytest.mark.parametrize("seed,th", [(5, 4.0), (6, 4.0), (7, 8.0)])
def test_sim3_fuse_equals_reference(host, seed, th):
    """ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (ORBmatcher.cc:1340-1455; LoopClosing::SearchAndFuse with th = 4): which points
    replace an existing map point, which become new observations of which key point, and the count"""
    k, pts, sim3 = sim3_scene(seed)
    rng = np.random.default_rng(300 + seed)
    k = dict(k, held=rng.choice([0, 1, 2], len(k["octaves"]), p=[0.5, 0.4, 0.1]).astype(np.uint8))
    nf_ref, rep_ref, add_ref = ref.fuse_sim3(k, pts, sim3, th)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    kx, o, d, held, fp, sf, cam = f32(k["kps_xy"]), i32(k["octaves"]), u8(k["desc"]), u8(k["held"]), f32(k["fp"]), f32(k["scale_factors"]), f32(k["cam4"])
    ps, pp, pn, pd, pmin, pmax = u8(pts["state"]), f32(pts["pos"]), f32(pts["normal"]), u8(pts["desc"]), f32(pts["min_dist"]), f32(pts["max_dist"])
    s3 = f32(sim3)
    rep, add = np.full(len(ps), -1, np.int32), np.full(len(ps), -1, np.int32)
    nf = host.gpuhost_fuse_sim3(_p(kx), _p(o), _p(d), len(kx), _p(fp), _p(held), _p(sf), len(sf), _p(s3), _p(cam), len(ps), _p(ps), _p(pp), _p(pn), _p(pd),
                                _p(pmin), _p(pmax), th, _p(rep), _p(add))
    assert nf == nf_ref and np.array_equal(rep, rep_ref) and np.array_equal(add, add_ref)
    assert nf_ref > 100 and (rep_ref >= 0).sum() > 20 and (add_ref >= 0).sum() > 20


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")
@pytest.mark.parametrize("levelsup,nnratio,check,seed", [(2, 0.8, True, 17), (3, 0.75, True, 18), (2, 0.9, False, 19)])
def test_search_by_bow_between_key_frames_equals_reference(host, levelsup, nnratio, check, seed):
    """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:765-905, LoopClosing.cc:1680): both sides need good map points, a key
    point of the second key frame is matched at most once"""
    sc = bow_scene(seed=seed, levelsup=levelsup)
    rng = np.random.default_rng(seed)
    st1 = np.where(sc["has_point"] > 0, rng.choice([1, 2], sc["nk"], p=[0.9, 0.1]), 0).astype(np.uint8)
    st2 = rng.choice([0, 1, 2], sc["nf"], p=[0.2, 0.7, 0.1]).astype(np.uint8)
    nm_ref, match_ref = ref.search_by_bow_kf(sc["ang_k"], sc["dk"], st1, sc["fv_k"], sc["ang_f"], sc["df"], st2, sc["fv_f"], nnratio, check)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    a1, d1, a2, d2 = f32(sc["ang_k"]), u8(sc["dk"]), f32(sc["ang_f"]), u8(sc["df"])
    n1_, s1_, f1_ = (i32(a) for a in sc["fv_k"])
    n2_, s2_, f2_ = (i32(a) for a in sc["fv_f"])
    match = np.full(len(a1), -1, np.int32)
    r0 = host.gpuhost_rescans()
    nm = host.gpuhost_search_by_bow_kf(_p(a1), _p(d1), _p(st1), len(a1), _p(n1_), _p(s1_), _p(f1_), len(n1_), len(f1_), _p(a2), _p(d2), _p(st2), len(a2),
                                       _p(n2_), _p(s2_), _p(f2_), len(n2_), len(f2_), nnratio, int(check), _p(match))
    assert nm == nm_ref and np.array_equal(match, match_ref)
    assert nm_ref > 60 and host.gpuhost_rescans() > r0


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")
@pytest.mark.parametrize("crowd,jitter,window,check", [(False, 0.0, 100, True), (True, 0.0, 100, True), (False, 6.0, 40, True), (True, 3.0, 100, False)])
def test_search_for_initialization_equals_reference(host, crowd, jitter, window, check):
    """ORBmatcher::SearchForInitialization (ORBmatcher.cc:648-766, Tracking::MonocularInitialization): matches, match count and the updated
    vbPrevMatched of the compiled GPU replacement against the reference's own body on the same two frames"""
    f1, f2, prev = init_scene(crowd=crowd, jitter=jitter)
    nm_ref, m_ref, prev_ref = ref.search_for_initialization(f1, f2, prev, window, 0.9, check)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    o1, a1, d1 = i32(f1["octaves"]), f32(f1["angles"]), u8(f1["desc"])
    k2, o2, a2, d2, fp = f32(f2["kps_xy"]), i32(f2["octaves"]), f32(f2["angles"]), u8(f2["desc"]), f32(f2["fp"])
    pv = f32(prev).copy()
    m12 = np.full(len(o1), -1, np.int32)
    nm = host.gpuhost_search_for_initialization(_p(o1), _p(a1), _p(d1), len(o1), _p(k2), _p(o2), _p(a2), _p(d2), len(o2), _p(fp), _p(pv), window, 0.9,
                                                int(check), _p(m12))
    assert nm == nm_ref and np.array_equal(m12, m_ref) and np.array_equal(pv, prev_ref)
    assert nm_ref > 20
    # a second round from the updated centres, as Tracking does frame after frame until the map is initialised (Tracking.cc:2527)
    nm_ref2, m_ref2, prev_ref2 = ref.search_for_initialization(f1, f2, prev_ref, window, 0.9, check)
    m12b = np.full(len(o1), -1, np.int32)
    nm2 = host.gpuhost_search_for_initialization(_p(o1), _p(a1), _p(d1), len(o1), _p(k2), _p(o2), _p(a2), _p(d2), len(o2), _p(fp), _p(pv), window, 0.9,
                                                 int(check), _p(m12b))
    assert nm2 == nm_ref2 and np.array_equal(m12b, m_ref2) and np.array_equal(pv, prev_ref2)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")
@pytest.mark.parametrize("with_stereo,th", [(False, 1.0), (True, 1.0), (False, 3.0)])
def test_search_local_points_equals_reference(host, with_stereo, th):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:43-213): 1600 projected map points compete for ~1000 key points"""
    s = local_points_scene(with_stereo)
    k, d = s["k"], s["d"]
    kxy = np.ascontiguousarray(np.stack([k["x"], k["y"]], 1), np.float32)
    octs = np.ascontiguousarray(k["octave"], np.int32)
    nm_ref, match_ref = ref.search_by_projection(kxy, octs, d, s["grid4"], s["sf"], s["proj"], s["level"], s["mp_desc"], s["in_view"], s["u_right"],
                                                 s["has_point"], nnratio=0.8, th=th)
    match = np.full(len(k), -1, np.int32)
    sf = np.ascontiguousarray(s["sf"], np.float32)
    r0 = host.gpuhost_rescans()
    nm = host.gpuhost_search_by_projection(_p(kxy), _p(octs), _p(d), len(k), _p(s["grid4"]), _p(s["u_right"]), _p(s["has_point"]), _p(sf), len(sf),
                                           _p(s["proj"]), _p(s["level"]), _p(s["mp_desc"]), _p(s["in_view"]), len(s["proj"]), 0.8, th, _p(match))
    assert nm == nm_ref and np.array_equal(match, match_ref)
    assert nm > 300
    assert host.gpuhost_rescans() - r0 < 0.2 * len(s["proj"])      # the four-candidate lists resolve most collisions on the host


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")
@pytest.mark.parametrize("stereo,direction,dense", [(False, 0, False), (True, 0, False), (True, 1, False), (True, -1, True), (False, 0, True)])
def test_motion_model_search_equals_reference(host, stereo, direction, dense):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1676-1887, Tracking::TrackWithMotionModel)"""
    cur, last = motion_scene(stereo, direction, dense=dense)
    th = 7 if stereo else 15
    nm_ref, match_ref = ref.search_by_projection_motion(cur, last, th, mono=not stereo)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    a = [f32(cur["kps_xy"]), i32(cur["octaves"]), f32(cur["angles"]), u8(cur["desc"])]
    n = len(a[1])
    ur = None if cur["u_right"] is None else f32(cur["u_right"])
    b = [f32(cur["fp"]), ur, u8(cur["state"]), f32(cur["scale_factors"])]
    c = [f32(cur["Tcw"]), f32(cur["cam4"])]
    l = [i32(last["octaves"]), f32(last["angles"]), u8(last["state"]), u8(last["outlier"]), f32(last["pos"]), u8(last["desc"]), f32(last["Tlw"])]
    match = np.full(n, -1, np.int32)
    nm = host.gpuhost_search_by_projection_motion(*[_p(x) for x in a], n, *[_p(x) for x in b], len(b[3]), *[_p(x) for x in c], len(l[0]),
                                                  *[_p(x) for x in l], th, int(not stereo), 0.9, 1, _p(match))
    assert nm == nm_ref and np.array_equal(match, match_ref)
    assert nm_ref > 30


def _frame_view(kxy, kstride, octs, ostride, desc, ur, n, on_device):
    v = capi.FrameView()
    v.kps_xy, v.kps_stride, v.octaves, v.oct_stride, v.desc, v.u_right, v.n, v.on_device = kxy, kstride, octs, ostride, desc, ur, n, on_device
    return v


def test_topk_scan_host_uploaded_and_extractor_resident_frames_agree_with_the_oracle():
    """orbb_search_area_topk: (a) a host frame view, (b) the same frame uploaded once with orbb_frame_upload, (c) the extractor's own
    device buffers (orbb_batch_device_ptrs: 24-byte key point records, the descriptors never leave the GPU) give the same lists; their
    first two entries are the oracle's best / second best; the lists are sorted by (distance, scan order) and hold no duplicates"""
    lib = capi.load()
    img = synth.frame(480, 752, 12)
    ge = ORBextractor(1000, 1.2, 8)
    mono, k, d = ge(img, None, (0, 0))
    n = len(k)
    rng = np.random.default_rng(3)
    nq = 700
    src = rng.integers(0, n, nq)
    queries = np.stack([k["x"][src] + rng.normal(0, 3, nq), k["y"][src] + rng.normal(0, 3, nq), rng.uniform(4, 30, nq), k["x"][src] - 10], 1).astype(np.float32)
    qlev = np.stack([np.maximum(k["octave"][src] - 1, 0), k["octave"][src] + 1], 1).astype(np.int32)
    qdesc = d[src].copy()
    qdesc[:, 5] ^= 0x3c
    skip = (rng.random(n) < 0.2).astype(np.uint8)
    grid4 = np.float32([0, 0, 64 / 752, 48 / 480])
    kxy = np.ascontiguousarray(np.stack([k["x"], k["y"]], 1), np.float32)
    octs = np.ascontiguousarray(k["octave"], np.int32)
    want = port.search_area_best2(kxy, octs, d, grid4, queries, qlev, qdesc, skip, None, 256)
    m = ORBmatcher()
    outs = []
    hv = _frame_view(kxy.ctypes.data, 8, octs.ctypes.data, 4, d.ctypes.data, None, n, 0)
    dv = capi.FrameView()
    capi.check(lib.orbb_frame_upload(m._m, 1, C.byref(hv), C.byref(dv)), m._m, matcher=True)
    kp_dev, desc_dev, cnt_dev = C.c_void_p(), C.c_void_p(), C.c_void_p()
    capi.check(lib.orbb_batch_device_ptrs(ge._h, C.byref(kp_dev), C.byref(desc_dev), C.byref(cnt_dev)), ge._h)
    ev = _frame_view(kp_dev.value, 24, kp_dev.value + 20, 24, desc_dev.value, None, n, 1)      # (no distortion: mvKeysUn == mvKeys)
    for view in (hv, dv, ev):
        for kk in (2, 4, 8):
            out = np.zeros((nq, kk, 2), np.int32)
            capi.check(lib.orbb_search_area_topk(m._m, C.byref(view), _p(grid4), _p(queries), _p(qlev), _p(qdesc), nq, _p(skip), 256, kk, _p(out)),
                       m._m, matcher=True)
            assert np.array_equal(out[:, :2].reshape(nq, 4), want), kk
            valid = out[:, :, 1] >= 0
            assert (np.diff(np.where(valid, out[:, :, 0], 256), axis=1) >= 0).all()          # ascending distances
            for q in range(0, nq, 37):
                ids = out[q, valid[q], 1]
                assert len(set(ids.tolist())) == len(ids) and not skip[ids].any()
            outs.append(out)
    assert all(np.array_equal(outs[i], outs[i % 3]) for i in range(len(outs)))               # the three views agree for every k
    assert (outs[2][:, :, 1] >= 0).sum() > (outs[0][:, :, 1] >= 0).sum()                     # k = 8 really returns more


def test_best2_csr_with_device_resident_train_descriptors():
    lib = capi.load()
    ge = ORBextractor(600, 1.2, 6)
    mono, k, d = ge(synth.frame(300, 400, 5), None, (0, 0))
    n = len(k)
    rng = np.random.default_rng(1)
    nq = 200
    q = d[rng.integers(0, n, nq)].copy()
    q[:, 0] ^= 0x81
    rowptr = np.concatenate([[0], np.cumsum(rng.integers(0, 40, nq))]).astype(np.int32)
    cand = rng.integers(0, n, rowptr[-1]).astype(np.int32)
    want = port.best2_csr(q, d, cand, rowptr, 256)
    kp_dev, desc_dev, cnt_dev = C.c_void_p(), C.c_void_p(), C.c_void_p()
    capi.check(lib.orbb_batch_device_ptrs(ge._h, C.byref(kp_dev), C.byref(desc_dev), C.byref(cnt_dev)), ge._h)
    m = ORBmatcher()
    out = np.zeros((nq, 4), np.int32)
    capi.check(lib.orbb_best2_csr_dev(m._m, _p(q), nq, desc_dev, n, _p(cand), _p(rowptr), 256, _p(out)), m._m, matcher=True)
    assert np.array_equal(out, want)


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")
@pytest.mark.parametrize("seed,th", [(41, 1.0), (42, 3.0), (43, 1.0)])
def test_search_local_points_fisheye_stereo_equals_reference(host, seed, th):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) on a fisheye-stereo frame (Nleft != -1, ORBmatcher.cc:43-213 with its right-eye half
    :139-208): two key point sets with their own grids, stereo partners claimed across the eyes"""
    f, mp = fisheye_local_points_scene(seed)
    nm_ref, match_ref = ref.search_by_projection_fisheye(f, mp, 0.8, th)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    kl, ol, kr, orr, d, fp = f32(f["kps_l"]), i32(f["oct_l"]), f32(f["kps_r"]), i32(f["oct_r"]), u8(f["desc"]), f32(f["fp"])
    l2r, r2l, hp, sf = i32(f["l2r"]), i32(f["r2l"]), u8(f["has_point"]), f32(f["scale_factors"])
    pl, ll, il, pr, lr, ir, md = f32(mp["proj_l"]), i32(mp["level_l"]), u8(mp["in_view_l"]), f32(mp["proj_r"]), i32(mp["level_r"]), u8(mp["in_view_r"]), u8(mp["desc"])
    match = np.full(len(kl) + len(kr), -1, np.int32)
    r0 = host.gpuhost_rescans()
    nm = host.gpuhost_search_by_projection_fisheye(_p(kl), _p(ol), len(kl), _p(kr), _p(orr), len(kr), _p(d), _p(fp), _p(l2r), _p(r2l), _p(hp), _p(sf), len(sf),
                                                   _p(pl), _p(ll), _p(il), _p(pr), _p(lr), _p(ir), _p(md), len(pl), 0.8, th, _p(match))
    assert nm == nm_ref and np.array_equal(match, match_ref)
    nL = len(kl)
    assert (match_ref[:nL] >= 0).sum() > 100 and (match_ref[nL:] >= 0).sum() > 100
    if th >= 3:
        assert host.gpuhost_rescans() > r0      # wide windows: some four-candidate lists are used up by earlier matches


def fisheye_motion_case(host, direction, dense, check):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) on fisheye-stereo frames (ORBmatcher.cc:1676-1887 with the right-eye half
    :1798-1860).  Not collected here: the branch was written after this round's GPU budget was spent; tests/test_matcher_host_cpu.py runs it
    against the reference's body over the CPU test double of the scans (the double has predicted the GPU result of every other case in this file)."""
    cur, last = fisheye_motion_scene(direction, dense=dense)
    th = 7
    nm_ref, match_ref = ref.search_by_projection_motion_fisheye(cur, last, th, mono=False, check_orientation=check)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    kl, ol, al, kr, orr, ar = f32(cur["kps_l"]), i32(cur["oct_l"]), f32(cur["ang_l"]), f32(cur["kps_r"]), i32(cur["oct_r"]), f32(cur["ang_r"])
    d, fp, cs, sf, tcw, trl, cam = u8(cur["desc"]), f32(cur["fp"]), u8(cur["state"]), f32(cur["scale_factors"]), f32(cur["Tcw"]), f32(cur["Trl"]), f32(cur["cam4"])
    lo, la, ls, lout, lp, ld, tlw = i32(last["octaves"]), f32(last["angles"]), u8(last["state"]), u8(last["outlier"]), f32(last["pos"]), u8(last["desc"]), f32(last["Tlw"])
    match = np.full(len(kl) + len(kr), -1, np.int32)
    nm = host.gpuhost_search_by_projection_motion_fisheye(_p(kl), _p(ol), _p(al), len(kl), _p(kr), _p(orr), _p(ar), len(kr), _p(d), _p(fp), _p(cs), _p(sf), len(sf),
                                                          _p(tcw), _p(trl), _p(cam), len(lo), int(last["n_left"]), _p(lo), _p(la), _p(ls), _p(lout), _p(lp), _p(ld),
                                                          _p(tlw), th, 0, 0.9, int(check), _p(match))
    assert nm == nm_ref and np.array_equal(match, match_ref)
    nL = len(kl)
    assert (match_ref[:nL] >= 0).sum() > 30 and (match_ref[nL:] >= 0).sum() > 30

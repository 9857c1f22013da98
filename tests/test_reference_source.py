"""CPU: the reference's OWN extractor -- orb_slam3/src/ORBextractor.cc compiled unmodified into oracle/_ref/liborbref.so
(OpenCV replaced by the pinned stand-in of oracle/cvshim/) -- against the oracle port, the cv2-backed restatement and the
committed golden fixtures.  This pins the restated ORB-SLAM3 control flow (cell loop + threshold fallback,
DistributeOctTree, IC_Angle, steered BRIEF, lapping assembly, constructor tables) to the reference's object code; the
OpenCV primitives are pinned separately against python-cv2 (test_oracle_primitives.py)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import port, ref
from orb_slam3_ros_b200 import synth

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/liborbref.so not built (needs /root/reference)")
GOLD = Path(__file__).resolve().parent / "golden"


def _same(k0, d0, m0, k1, d1, m1):
    assert len(k0) == len(k1) and m0 == m1
    for f in ("x", "y", "size", "angle", "response", "octave"):
        assert np.array_equal(k0[f], k1[f]), f
    assert np.array_equal(d0, d1)


@pytest.mark.parametrize("nf,sf,nl,ini,mn", [(1000, 1.2, 8, 20, 7), (2000, 1.2, 8, 20, 7), (1250, 1.2, 8, 12, 7), (500, 1.5, 4, 20, 7),
                                             (300, 2.0, 3, 30, 10), (1500, 1.1, 10, 20, 7), (1, 1.2, 8, 20, 7)])
def test_constructor_tables_match_reference(nf, sf, nl, ini, mn):
    a, b = port.PortExtractor(nf, sf, nl, ini, mn), ref.RefExtractor(nf, sf, nl, ini, mn)
    for f in ("scale_factors", "inv_scale_factors", "level_sigma2", "inv_level_sigma2", "features_per_level", "umax"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f


@pytest.mark.parametrize("shape,nf,nl,lap,seed", [((480, 752), 1000, 8, (0, 1000), 1), ((480, 752), 1000, 8, (0, 0), 5),
                                                  ((376, 1241), 2000, 8, (0, 0), 2), ((240, 320), 300, 4, (100, 200), 3),
                                                  ((480, 640), 1000, 8, (250, 400), 4), ((200, 260), 5000, 4, (0, 0), 6),
                                                  ((200, 260), 500, 8, (0, 100), 3), ((150, 400), 300, 8, (0, 0), 3)])      # top levels without a cell
def test_port_equals_reference_source(shape, nf, nl, lap, seed):
    img = synth.frame(shape[0], shape[1], seed)
    pe, re_ = port.PortExtractor(nf, 1.2, nl), ref.RefExtractor(nf, 1.2, nl)
    rc, k0, d0, m0 = pe.extract(img, lap)
    rc2, k1, d1, m1 = re_.extract(img, lap)
    assert rc == 0 and rc2 == 0 and len(k1) > 0
    for l in range(nl):
        assert np.array_equal(pe.level(l, bordered=True), re_.level(l, bordered=True)), l
    _same(k0, d0, m0, k1, d1, m1)


@pytest.mark.parametrize("sf,nl,ini,mn", [(1.5, 4, 20, 7), (2.0, 3, 20, 7), (1.1, 6, 20, 7), (1.2, 8, 40, 5), (1.2, 8, 7, 7)])
def test_port_equals_reference_source_other_settings(sf, nl, ini, mn):
    img = synth.frame(480, 640, 1)
    rc, k0, d0, m0 = port.PortExtractor(500, sf, nl, ini, mn).extract(img)
    rc2, k1, d1, m1 = ref.RefExtractor(500, sf, nl, ini, mn).extract(img)
    assert rc == 0 and rc2 == 0
    _same(k0, d0, m0, k1, d1, m1)


def test_reference_source_on_noise_flat_and_strided_images():
    rng = np.random.default_rng(11)
    noise = rng.integers(0, 256, (300, 420), dtype=np.uint8)                      # far more corners than wanted: deep quadtree, sort ties
    flat = np.full((240, 320), 127, np.uint8)                                     # no corner anywhere: empty output, descriptors released
    wide = synth.frame(300, 500, 9)
    view = np.ascontiguousarray(np.pad(wide, ((0, 0), (0, 37))))[:, :500]         # row stride != width
    for img, nf in ((noise, 1000), (flat, 500), (view, 700)):
        rc, k0, d0, m0 = port.PortExtractor(nf, 1.2, 8).extract(img, (50, 120))
        rc2, k1, d1, m1 = ref.RefExtractor(nf, 1.2, 8).extract(img, (50, 120))
        assert rc == rc2 == 0
        _same(k0, d0, m0, k1, d1, m1)
    assert len(port.PortExtractor(500, 1.2, 8).extract(flat)[1]) == 0
    assert ref.RefExtractor().extract(np.zeros((0, 0), np.uint8))[0] == -1        # ORBextractor.cc:1090


@pytest.mark.parametrize("name", ["mono_320x240", "wide_400x200", "noise_176x144", "fallback_260x200"])
def test_reference_source_reproduces_golden(name):
    """the fixtures were generated through python-cv2 (tools/make_golden.py): reference object code + stand-in == real OpenCV run"""
    g = np.load(GOLD / f"{name}.npz")
    nf, nl, ini, mn, l0, l1 = [int(v) for v in g["params"]]
    rc, k, d, m = ref.RefExtractor(nf, 1.2, nl, ini, mn).extract(g["image"], (l0, l1))
    assert rc == 0
    _same(g["kps"].view(port.KP_DTYPE).reshape(-1), g["desc"], int(g["mono"]), k, d, m)


def test_reference_source_batch_matches_single_calls():
    imgs = np.stack([synth.frame(240, 320, s) for s in range(5)])
    counts, kps, desc = ref.extract_batch(imgs, 400, 1.2, 6, lapping=(0, 100), nthreads=3)
    e = ref.RefExtractor(400, 1.2, 6)
    for f in range(5):
        rc, k, d, m = e.extract(imgs[f], (0, 100))
        assert counts[f, 0] == len(k) and counts[f, 1] == m
        assert np.array_equal(kps[f, :len(k)], k) and np.array_equal(desc[f, :len(k)], d)


@pytest.mark.parametrize("k,depth,ragged,levelsup,n", [(10, 3, False, 2, 700), (10, 4, False, 4, 1500), (6, 5, True, 4, 900), (8, 3, True, 1, 300),
                                                       (10, 2, False, 4, 50), (5, 3, False, 3, 0)])
def test_bow_port_equals_vendored_dbow2(tmp_path, k, depth, ragged, levelsup, n):
    """Frame::ComputeBoW: the port's restatement against the reference's vendored DBoW2 (loadFromTextFile + transform), on a seeded
    vocabulary written in the ORBvoc.txt format; word ids, L1-normalised tf-idf values (bit-exact doubles), feature-vector nodes and
    their feature lists, all in std::map order."""
    from orb_slam3_ros_b200.bow import synthetic_vocabulary
    vocab = synthetic_vocabulary(k, depth, seed=k * 10 + depth, stop_fraction=0.1, ragged=ragged)
    ref.write_vocabulary_text(vocab, tmp_path / "voc.txt")
    rv = ref.RefVocabulary(tmp_path / "voc.txt")
    assert rv.words == int((vocab["node_word"] >= 0).sum())
    rng = np.random.default_rng(n + depth)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    if n > 10:                                                     # some descriptors equal to node descriptors: distance ties, exact hits
        desc[:10] = vocab["node_desc"][rng.integers(1, len(vocab["node_desc"]), 10)]
    got = port.bow_transform(vocab, desc, levelsup, 1)
    want = rv.transform(desc, levelsup)
    for g, w in zip(got[:5], want[:5]):
        assert np.array_equal(g, w)
    assert got[5] == want[5]


def test_descriptor_distance_equals_forb_distance():
    rng = np.random.default_rng(3)
    a, b = rng.integers(0, 256, (200, 32), dtype=np.uint8), rng.integers(0, 256, (200, 32), dtype=np.uint8)
    b[:5] = a[:5]
    a[5], b[5] = 0, 255
    for x, y in zip(a, b):
        assert port.hamming(x, y) == ref.descriptor_distance(x, y) == int(np.unpackbits(x ^ y).sum())


def test_random_settings_sweep_port_equals_reference_source():
    """seeded sweep over image sizes, level counts, scale factors, thresholds, budgets, lapping areas and image statistics (natural-like,
    low contrast = threshold-fallback cells, pure noise = deep tie-heavy quadtrees); geometry the reference cannot handle is skipped
    when the port says so (the reference itself would assert or throw there)"""
    rng = np.random.default_rng(20241018)
    checked = 0
    for it in range(60):
        h, w = int(rng.integers(70, 420)), int(rng.integers(70, 640))
        nl = int(rng.integers(1, 10))
        sf = float(rng.choice([1.2, 1.2, 1.1, 1.3, 1.5, 2.0, 1.25]))
        nf = int(rng.integers(1, 2500))
        ini = int(rng.integers(5, 70))
        mn = int(rng.integers(1, ini + 1))
        lap = (int(rng.integers(0, w)), int(rng.integers(0, w)))
        kind = int(rng.integers(0, 3))
        img = (rng.integers(0, 256, (h, w), dtype=np.uint8) if kind == 0 else
               np.clip(synth.frame(h, w, it).astype(np.int32) // 4 + 100, 0, 255).astype(np.uint8) if kind == 1 else synth.frame(h, w, 500 + it))
        rc, k0, d0, m0 = port.PortExtractor(nf, sf, nl, ini, mn).extract(img, lap)
        if rc == -2:
            continue
        assert rc == 0
        rc2, k1, d1, m1 = ref.RefExtractor(nf, sf, nl, ini, mn).extract(img, lap)
        assert rc2 == 0, (h, w, nl, sf, nf, ini, mn)
        _same(k0, d0, m0, k1, d1, m1)
        checked += 1
    assert checked >= 30


# ---- matcher / stereo / grid rows: the reference's own definitions, cut out of Frame.cc / ORBmatcher.cc at build time and compiled
# ---- inside stand-in classes (oracle/ref_cut_tu.cpp) ------------------------------------------------------------------------------
def test_matcher_constants_and_descriptor_distance_are_the_reference_ones():
    assert ref.matcher_constants() == (50, 100, 30)                       # TH_LOW, TH_HIGH, HISTO_LENGTH (ORBmatcher.cc:35-37)
    rng = np.random.default_rng(4)
    a, b = rng.integers(0, 256, (300, 32), dtype=np.uint8), rng.integers(0, 256, (300, 32), dtype=np.uint8)
    b[:4] = a[:4]
    a[4], b[4] = 0, 255
    for x, y in zip(a, b):
        assert ref.matcher_descriptor_distance(x, y) == port.hamming(x, y) == int(np.unpackbits(x ^ y).sum())


@pytest.mark.parametrize("shape,nf,dmax,bf,b", [((376, 1241), 2000, 60, 386.1448, 0.53716), ((480, 752), 1200, 40, 47.9, 0.11),
                                                ((240, 400), 600, 30, 380.0, 0.5)])
def test_stereo_port_equals_reference_compute_stereo_matches(shape, nf, dmax, bf, b):
    """Frame::ComputeStereoMatches (Frame.cc:811-981): row table, descriptor search, 11x11 SAD sliding window on the two extractors'
    pyramids, parabola refinement, median cut -- the reference's text on the reference extractor's pyramids vs the port"""
    left, right = synth.stereo_pair(shape[0], shape[1], nf % 7, dmax=dmax)
    pl, pr = port.PortExtractor(nf, 1.2, 8), port.PortExtractor(nf, 1.2, 8)
    _, kl, dl, _ = pl.extract(left)
    _, kr, dr, _ = pr.extract(right)
    rl, rr = ref.RefExtractor(nf, 1.2, 8), ref.RefExtractor(nf, 1.2, 8)
    assert np.array_equal(rl.extract(left)[1], kl) and np.array_equal(rr.extract(right)[1], kr)
    ur, dp, _, _, kept = port.stereo(pl, pr, kl, dl, kr, dr, np.float32(bf), np.float32(b))
    ur2, dp2 = ref.stereo(rl, rr, kl, dl, kr, dr, bf, b)
    assert kept == int((ur2 >= 0).sum()) and kept > 0.2 * len(kl)
    assert np.array_equal(ur.view(np.uint32), ur2.view(np.uint32)) and np.array_equal(dp.view(np.uint32), dp2.view(np.uint32))


def test_rotation_histogram_maxima_equal_reference_compute_three_maxima():
    rng = np.random.default_rng(6)
    cases = [rng.integers(0, 40, 30) for _ in range(40)]
    cases += [np.zeros(30, np.int64), np.full(30, 7), np.eye(30, dtype=np.int64)[3] * 9, np.r_[np.full(15, 10), np.zeros(15, np.int64)],
              np.r_[100, 9, 10, np.zeros(27, np.int64)], np.r_[100, 10, 9, 11, np.zeros(26, np.int64)]]       # ties, 10 % rule edges
    for counts in cases:
        counts = np.asarray(counts, np.int64)
        assert ref.three_maxima(counts) == port.three_maxima(counts), counts.tolist()
    # through the whole filter: the bins of real angle differences (round(rot / 30), ORBmatcher.cc:236,:345-352) and their maxima
    for seed in range(5):
        r = np.random.default_rng(seed)
        a, b = r.uniform(0, 360, 500).astype(np.float32), r.uniform(0, 360, 500).astype(np.float32)
        b[:300] = (a[:300] + r.normal(25, 6, 300)).astype(np.float32) % np.float32(360)
        keep, inds = port.rotation_check(a, b)
        rot = np.where(a - b < 0, a - b + np.float32(360), a - b).astype(np.float32)
        bins = np.floor((rot * (np.float32(1) / np.float32(30))).astype(np.float64) + 0.5).astype(np.int64) % 30
        assert ref.three_maxima(np.bincount(bins, minlength=30)) == tuple(int(v) for v in inds)


def test_rgbd_port_equals_reference_compute_stereo_from_rgbd():
    rng = np.random.default_rng(12)
    h, w, n = 480, 640, 900
    depth = (rng.integers(0, 40000, (h, w)).astype(np.float32) * (np.float32(1) / np.float32(5000))).astype(np.float32)
    depth[rng.random((h, w)) < 0.2] = 0
    xy = np.stack([rng.uniform(16, w - 17, n), rng.uniform(16, h - 17, n)], 1).astype(np.float32)
    K4 = (517.306408, 516.469215, 318.643040, 255.313989)
    for dist in ([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], [0.0, 0.0, 0.0, 0.0]):
        ur, dp = port.rgbd_stereo(xy, depth, K4, dist, np.float32(40.0))
        xu = port.undistort_points(xy, K4, dist)[:, 0] if dist[0] else xy[:, 0]
        ur2, dp2 = ref.rgbd_stereo(xy, xu, depth, 40.0)
        assert np.array_equal(dp, dp2) and np.array_equal(ur.view(np.uint32), ur2.view(np.uint32))


@pytest.mark.parametrize("with_stereo", [False, True])
def test_search_area_port_equals_reference_grid_functions(with_stereo):
    """Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (reference text) + best/second scan vs the port, on projected points
    near key points, outside the image and exactly on cell borders"""
    h, w = 480, 752
    _, k, d, _ = port.PortExtractor(1000, 1.2, 8).extract(synth.frame(h, w, 3))
    n = len(k)
    rng = np.random.default_rng(5)
    kps_xy = np.stack([k["x"], k["y"]], 1)
    grid4 = np.float32([0.0, 0.0, np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)])      # Frame.cc:251-252
    nq = 1200
    src = rng.integers(0, n, nq)
    qx = k["x"][src] + rng.normal(0, 3, nq).astype(np.float32)
    qy = k["y"][src] + rng.normal(0, 3, nq).astype(np.float32)
    qx[::50] = rng.choice([-40.0, w + 60.0, 0.0, w / 64 * 7], len(qx[::50])).astype(np.float32)
    qy[::70] = rng.choice([-30.0, h + 50.0, 0.0, h / 48 * 5], len(qy[::70])).astype(np.float32)
    r = (rng.choice([2.5, 4.0], nq) * rng.choice([1.0, 1.2, 1.44, 3.0, 15.0], nq)).astype(np.float32)
    lvl = k["octave"][src]
    qlev = np.stack([lvl - 1, lvl], 1).astype(np.int32)
    qlev[::9] = (-1, -1)
    qlev[5::9, 1] = -1
    qdesc = d[src].copy()
    qdesc[:, :3] ^= rng.integers(0, 256, (nq, 3), dtype=np.uint8)
    skip = (rng.random(n) < 0.2).astype(np.uint8)
    u_right = np.where(rng.random(n) < 0.5, k["x"] - rng.uniform(1, 40, n), -1).astype(np.float32) if with_stereo else None
    queries = np.stack([qx, qy, r, qx - rng.uniform(0, 45, nq).astype(np.float32)], 1).astype(np.float32)
    for init in (256, 100):
        want = ref.search_area_best2(kps_xy, k["octave"], d, grid4, queries, qlev, qdesc, skip, u_right, init)
        got = port.search_area_best2(kps_xy, k["octave"], d, grid4, queries, qlev, qdesc, skip, u_right, init)
        assert np.array_equal(want, got), init
    assert (want[:, 1] >= 0).mean() > 0.5


def test_distinctive_port_equals_reference_compute_distinctive_descriptors():
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:329-403, reference text): observations with a left and/or right index,
    bad key frames skipped, least median distance incl. the self distance, first minimum on ties"""
    rng = np.random.default_rng(21)
    for case in range(60):
        nkf = int(rng.integers(1, 40))
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        kf_desc = np.repeat(base[None, None, :], nkf * 2, 0).reshape(nkf, 2, 32).copy()
        flips = rng.integers(0, 256, (nkf, 2, 32), dtype=np.uint8) & rng.integers(0, 256, (nkf, 2, 32), dtype=np.uint8) & rng.integers(0, 256, (nkf, 2, 32), dtype=np.uint8)
        kf_desc ^= flips                                                   # observations of one point: close to a common descriptor
        if case % 5 == 0:
            kf_desc[:, 0] = kf_desc[0, 0]                                  # many identical descriptors: ties
        lr = np.stack([rng.choice([0, -1], nkf, p=[0.8, 0.2]), rng.choice([1, -1], nkf, p=[0.3, 0.7])], 1).astype(np.int32)
        bad = (rng.random(nkf) < 0.15).astype(np.uint8)
        rows = [kf_desc[i, j] for i in range(nkf) if not bad[i] for j in (0, 1) if lr[i, j] != -1]      # order of vDescriptors (:347-361)
        got = ref.distinctive(kf_desc, lr, bad)
        if not rows:
            assert got is None
            continue
        group = np.stack(rows)
        best = port.distinctive(group, np.int32([0, len(group)]))[0]
        assert np.array_equal(got, group[best]), case


def _batched_search_local_points(k, d, grid4, sf, proj, level, mp_desc, in_view, u_right, has_point, nnratio, th, scan):
    """one batched grid lookup + best/second scan for all map points of the call (`scan` = the oracle's or the
    library's search_area_best2), then the reference's per-point decision in list order.  The reference updates F.mvpMapPoints while it
    walks the list, and a key point that got a map point is skipped by the later ones (ORBmatcher.cc:88-90); the batched scan saw
    the key points as they were BEFORE the call, so a point whose best or second candidate was taken meanwhile is scanned again with
    the current mask (excluding any other candidate cannot change its best two)."""
    kps_xy = np.stack([k["x"], k["y"]], 1)
    oct_ = k["octave"].astype(np.int32)
    r = np.where(proj[:, 3] > np.float32(0.998), np.float32(2.5), np.float32(4.0)).astype(np.float32)      # RadiusByViewingCos :215-221
    if th != 1.0:
        r = (r * np.float32(th)).astype(np.float32)
    rq = (r * sf[level]).astype(np.float32)
    queries = np.stack([proj[:, 0], proj[:, 1], rq, proj[:, 2]], 1).astype(np.float32)
    qlev = np.stack([level - 1, level], 1).astype(np.int32)
    skip = np.zeros(len(k), np.uint8) if has_point is None else has_point.astype(np.uint8).copy()
    out = scan(kps_xy, oct_, d, grid4, queries, qlev, mp_desc, skip, u_right, 256)
    match_of = np.full(len(k), -1, np.int32)
    nmatches = rescans = 0
    for j in range(len(proj)):
        if not in_view[j]:
            continue
        d1, i1, d2, i2 = (int(v) for v in out[j])
        if (i1 >= 0 and skip[i1]) or (i2 >= 0 and skip[i2]):
            d1, i1, d2, i2 = (int(v) for v in scan(kps_xy, oct_, d, grid4, queries[j:j + 1], qlev[j:j + 1], mp_desc[j:j + 1], skip, u_right, 256)[0])
            rescans += 1
        if i1 < 0 or d1 > 100:                                             # TH_HIGH :122
            continue
        l1, l2 = int(oct_[i1]), (int(oct_[i2]) if i2 >= 0 else -1)
        if l1 == l2 and np.float32(d1) > np.float32(nnratio) * np.float32(d2):      # :124-125
            continue
        match_of[i1] = j                                                   # :127-128
        skip[i1] = 1
        nmatches += 1
    return nmatches, match_of, rescans


@pytest.mark.parametrize("with_stereo,th", [(False, 1.0), (True, 1.0), (False, 3.0)])
def test_batched_search_local_points_equals_reference_search_by_projection(with_stereo, th):
    """the reference's own SearchByProjection(F, vpMapPoints, th) (Tracking::SearchLocalPoints) against the batched formulation of the
    integration notes, on a scene where many map points compete for the same key points"""
    h, w = 480, 752
    pe = port.PortExtractor(1000, 1.2, 8)
    _, k, d, _ = pe.extract(synth.frame(h, w, 8))
    n = len(k)
    rng = np.random.default_rng(31)
    grid4 = np.float32([0.0, 0.0, np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)])
    nmp = 1600                                                             # more map points than key points: collisions guaranteed
    src = rng.integers(0, n, nmp)
    proj = np.stack([k["x"][src] + rng.normal(0, 2.5, nmp), k["y"][src] + rng.normal(0, 2.5, nmp),
                     k["x"][src] - rng.uniform(0, 40, nmp), rng.choice([0.9, 0.999, 0.9985], nmp)], 1).astype(np.float32)
    level = np.clip(k["octave"][src] + rng.integers(-1, 2, nmp), 0, 7).astype(np.int32)
    mp_desc = d[src].copy()
    mp_desc[:, :4] ^= rng.integers(0, 256, (nmp, 4), dtype=np.uint8) & rng.integers(0, 256, (nmp, 4), dtype=np.uint8)
    in_view = (rng.random(nmp) < 0.9).astype(np.uint8)
    has_point = (rng.random(n) < 0.25).astype(np.uint8)
    u_right = np.where(rng.random(n) < 0.5, k["x"] - rng.uniform(1, 40, n), -1).astype(np.float32) if with_stereo else None
    nm_ref, match_ref = ref.search_by_projection(np.stack([k["x"], k["y"]], 1), k["octave"], d, grid4, pe.scale_factors, proj, level, mp_desc,
                                                 in_view, u_right, has_point, nnratio=0.8, th=th)
    nm, match, rescans = _batched_search_local_points(k, d, grid4, pe.scale_factors, proj, level, mp_desc, in_view, u_right, has_point, 0.8, th,
                                                      port.search_area_best2)
    assert nm == nm_ref and np.array_equal(match, match_ref)
    assert nm > 300 and rescans > 0                                       # collisions really happened (extreme here: 1600 points for ~1000 key points)


def test_batched_search_by_bow_equals_reference_search_by_bow():
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ..) (TrackReferenceKeyFrame; reference text over the vendored DBoW2::FeatureVector)
    against the batched formulation: feature vectors from the BoW transform, one best-two scan over the candidate lists of all
    key-frame features (init 256), the in-order decision (frame features matched earlier in the call are skipped, :276-277; TH_LOW
    and the strict float ratio test, :322-324) and the rotation-histogram filter (:329-341, :396-415)"""
    from scenes import bow_scene
    sc = bow_scene()
    ang_k, dk, has_point, fv_k, df, fv_f = sc["ang_k"], sc["dk"], sc["has_point"], sc["fv_k"], sc["df"], sc["fv_f"]
    kk, kf = np.zeros(sc["nk"]), np.zeros(sc["nf"], dtype=[("angle", np.float32)])
    kf["angle"] = sc["ang_f"]
    nnratio = 0.7
    nm_ref, match_ref = ref.search_by_bow(ang_k, dk, has_point, fv_k, kf["angle"], df, fv_f, nnratio, True)

    def groups(fv, n):
        node, start, feat = fv
        ends = list(start[1:]) + [len(feat)]
        return {int(nd): feat[s:e] for nd, s, e in zip(node, start, ends)}
    gk, gf = groups(fv_k, len(kk)), groups(fv_f, len(kf))
    queries, cand, rowptr = [], [], [0]
    for nd in sorted(set(gk) & set(gf)):                                  # the merge loop :243-394 visits common nodes in ascending order
        for q in gk[nd]:
            if has_point[q]:
                queries.append(int(q))
                cand.extend(int(c) for c in gf[nd])
                rowptr.append(len(cand))
    cand, rowptr = np.int32(cand), np.int32(rowptr)
    best = port.best2_csr(dk[queries], df, cand, rowptr, 256)
    taken = np.zeros(len(kf), bool)
    match = np.full(len(kf), -1, np.int32)
    pairs, rescans = [], 0
    for j, q in enumerate(queries):
        d1, i1, d2, i2 = (int(v) for v in best[j])
        if (i1 >= 0 and taken[i1]) or (i2 >= 0 and taken[i2]):
            c = np.int32([x for x in cand[rowptr[j]:rowptr[j + 1]] if not taken[x]])
            d1, i1, d2, i2 = (int(v) for v in port.best2_csr(dk[q:q + 1], df, c, np.int32([0, len(c)]), 256)[0])
            rescans += 1
        if i1 < 0 or d1 > 50:                                              # TH_LOW :320
            continue
        if not (np.float32(d1) < np.float32(nnratio) * np.float32(d2)):    # :322
            continue
        match[i1] = q
        taken[i1] = True
        pairs.append((q, i1))
    keep, _ = port.rotation_check(ang_k[[p[0] for p in pairs]], kf["angle"][[p[1] for p in pairs]])
    for (q, i1), kp in zip(pairs, keep):
        if not kp:
            match[i1] = -1
    assert int((match >= 0).sum()) == nm_ref and np.array_equal(match, match_ref)
    assert nm_ref > 100 and rescans > 0 and (~keep).sum() > 0              # matches, collisions and rotation rejects all occurred


def _batched_motion_model(cur, last, th, mono, scan, check_orientation=True):
    """ORBmatcherGPU::SearchByProjection(CurrentFrame, LastFrame, th, bMono) restated in numpy (orb_slam3_ros_b200/host/ORBmatcherGPU.cc):
    all projections first, ONE batched grid lookup + best scan (`scan` = the oracle's search_area_best2), then the reference's
    per-point decisions in order on the live state -- a key point that received a point WITH observations is skipped by later
    points (ORBmatcher.cc:1749-1751), one that received a temporal point is not and can be overwritten."""
    f32 = np.float32
    R, t = cur["Tcw"][:9].astype(f32), cur["Tcw"][9:].astype(f32)
    Rl, tl = last["Tlw"][:9].astype(f32), last["Tlw"][9:].astype(f32)

    def apply(Rm, tv, p):
        return [f32(f32(f32(f32(Rm[3 * i] * p[0]) + f32(Rm[3 * i + 1] * p[1])) + f32(Rm[3 * i + 2] * p[2])) + tv[i]) for i in range(3)]
    Ri = [R[3 * j + i] for i in range(3) for j in range(3)]                     # inverse: (R^T, -R^T t)
    twc = [f32(-f32(f32(f32(Ri[3 * i] * t[0]) + f32(Ri[3 * i + 1] * t[1])) + f32(Ri[3 * i + 2] * t[2]))) for i in range(3)]
    tlc = apply(Rl, tl, twc)
    mb, mbf = f32(cur["fp"][7]), f32(cur["fp"][6])
    forward, backward = (tlc[2] > mb and not mono), (-tlc[2] > mb and not mono)
    fx, fy, cx, cy = (f32(v) for v in cur["cam4"])
    sf = cur["scale_factors"]
    queries, qlev, qdesc, src = [], [], [], []
    for i in range(len(last["octaves"])):
        if not last["state"][i] or last["outlier"][i]:
            continue
        x, y, z = apply(R, t, last["pos"][i].astype(f32))
        invz = f32(1.0 / float(z))
        if invz < 0:
            continue
        u, v = f32(f32(f32(fx * x) / z) + cx), f32(f32(f32(fy * y) / z) + cy)
        if u < cur["fp"][0] or u > cur["fp"][1] or v < cur["fp"][2] or v > cur["fp"][3]:
            continue
        o = int(last["octaves"][i])
        queries.append([u, v, f32(f32(th) * sf[o]), f32(u - f32(mbf * invz))])
        qlev.append([o, -1] if forward else [0, o] if backward else [o - 1, o + 1])
        qdesc.append(last["desc"][i])
        src.append(i)
    queries, qlev, qdesc = np.float32(queries).reshape(-1, 4), np.int32(qlev).reshape(-1, 2), np.uint8(qdesc).reshape(-1, 32)
    grid4 = np.float32([cur["fp"][0], cur["fp"][2], cur["fp"][4], cur["fp"][5]])
    holds = cur["state"].astype(np.int64).copy()                                 # 0 none, 1 with observations, 2 without
    assigned = np.full(len(holds), -1, np.int32)
    skip = (holds == 1).astype(np.uint8)
    out = scan(cur["kps_xy"], cur["octaves"], cur["desc"], grid4, queries, qlev, qdesc, skip, cur["u_right"], 256)
    nmatches, rescans, votes = 0, 0, []
    for j, i in enumerate(src):
        d1, i1 = int(out[j][0]), int(out[j][1])
        if i1 >= 0 and holds[i1] == 1:                                           # its best candidate was taken meanwhile: scan this point again
            skip = (holds == 1).astype(np.uint8)
            d1, i1 = (int(v) for v in scan(cur["kps_xy"], cur["octaves"], cur["desc"], grid4, queries[j:j + 1], qlev[j:j + 1], qdesc[j:j + 1], skip,
                                           cur["u_right"], 256)[0][:2])
            rescans += 1
        if i1 < 0 or d1 > 100:
            continue
        holds[i1] = last["state"][i]
        assigned[i1] = i
        nmatches += 1
        if check_orientation:
            rot = f32(last["angles"][i] - cur["angles"][i1])
            if rot < 0:
                rot = f32(rot + f32(360))
            b = int(np.round(f32(rot * f32(1.0 / 30))))                          # round(): half away from zero; rot >= 0 here
            b = int(np.floor(float(f32(rot * f32(1.0 / 30))) + 0.5))
            votes.append((0 if b == 30 else b, i1))
    if check_orientation:
        hist = np.bincount([b for b, _ in votes], minlength=30)
        keep3 = set(int(v) for v in port.three_maxima(hist) if v >= 0)
        for b, i1 in votes:
            if b not in keep3:
                assigned[i1] = -1
                nmatches -= 1
    return nmatches, assigned, rescans


@pytest.mark.parametrize("stereo,direction,dense", [(False, 0, False), (True, 0, False), (True, 1, False), (True, -1, True), (False, 0, True)])
def test_batched_motion_model_search_equals_reference_search_by_projection(stereo, direction, dense):
    """the reference's own SearchByProjection(CurrentFrame, LastFrame, th, bMono) (Tracking::TrackWithMotionModel; its definition cut out of
    ORBmatcher.cc:1676-1887 and compiled against the Eigen / Sophus stand-ins of oracle/cvshim/mini_geom.hpp) against the batched
    formulation that the GPU host adapter implements, with the oracle's scan in the place of the device scan"""
    from scenes import motion_scene
    cur, last = motion_scene(stereo, direction, dense=dense)
    th = 7 if stereo else 15                                                     # Tracking.cc:2918-2923
    nm_ref, match_ref = ref.search_by_projection_motion(cur, last, th, mono=not stereo)
    nm, match, rescans = _batched_motion_model(cur, last, th, not stereo, port.search_area_best2)
    assert nm == nm_ref and np.array_equal(match, match_ref)
    assert nm_ref > 30
    if dense:
        assert rescans > 0


def _batched_search_for_initialization(f1, f2, prev, window, nnratio, check_orientation, scan, k=4):
    """ORBmatcherGPU::SearchForInitialization restated in numpy (orb_slam3_ros_b200/host/ORBmatcherGPU.cc): ONE batched scan for the k best
    candidates of every level-0 key point (`scan` = the oracle's best-two scan, applied twice with the first two masked to get four), then
    the reference's decisions in order on the live vMatchedDistance; a key point that loses more than k - 2 of a full list walks its window
    again with the reference's own loop."""
    f32 = np.float32
    n1, n2 = len(f1["octaves"]), len(f2["octaves"])
    grid4 = f32([f2["fp"][0], f2["fp"][2], f2["fp"][4], f2["fp"][5]])
    src = [i for i in range(n1) if f1["octaves"][i] == 0]
    queries = f32([[prev[i][0], prev[i][1], window, -1] for i in src]).reshape(-1, 4)
    qlev = np.zeros((len(src), 2), np.int32)
    qdesc = f1["desc"][src]
    args = (f2["kps_xy"], f2["octaves"], f2["desc"], grid4)
    first = scan(*args, queries, qlev, qdesc, None, None, 257)
    lists = []
    for j in range(len(src)):
        lst = [(int(first[j][0]), int(first[j][1])), (int(first[j][2]), int(first[j][3]))]
        if lst[1][1] >= 0 and k > 2:
            skip = np.zeros(n2, np.uint8)
            skip[[lst[0][1], lst[1][1]]] = 1
            nxt = scan(*args, queries[j:j + 1], qlev[j:j + 1], qdesc[j:j + 1], skip, None, 257)[0]
            lst += [(int(nxt[0]), int(nxt[1])), (int(nxt[2]), int(nxt[3]))]
        lists.append([c for c in lst if c[1] >= 0])
    INF = 2 ** 31 - 1
    matched_dist = np.full(n2, INF, np.int64)
    m12, m21 = np.full(n1, -1, np.int32), np.full(n2, -1, np.int32)
    nm, fallbacks, votes = 0, 0, []
    gx = np.floor((f2["kps_xy"][:, 0] - grid4[0]) * grid4[2] + f32(0.5)).astype(np.int64)       # round(): the coordinates are >= 0 here
    gy = np.floor((f2["kps_xy"][:, 1] - grid4[1]) * grid4[3] + f32(0.5)).astype(np.int64)
    in_grid = (gx >= 0) & (gx < 64) & (gy >= 0) & (gy < 48) & (f2["octaves"] == 0)
    for j, i1 in enumerate(src):
        live = [(d, i) for d, i in lists[j] if matched_dist[i] > d]
        if len(live) < 2 and len(lists[j]) == k:                                # the reference's loop for this one (ORBmatcher.cc:668-700)
            x, y = f32(prev[i1][0]), f32(prev[i1][1])
            cand = np.flatnonzero(in_grid & (np.abs(f2["kps_xy"][:, 0] - x) < window) & (np.abs(f2["kps_xy"][:, 1] - y) < window))
            cand = sorted(cand, key=lambda i: (gx[i], gy[i], i))                 # cells column by column, index order inside a cell
            best, best2, bidx = INF, INF, -1
            for i2 in cand:
                d = port.hamming(f1["desc"][i1], f2["desc"][i2])
                if matched_dist[i2] <= d:
                    continue
                if d < best:
                    best2, best, bidx = best, d, i2
                elif d < best2:
                    best2 = d
            fallbacks += 1
        else:
            best, bidx = live[0] if live else (INF, -1)
            best2 = live[1][0] if len(live) > 1 else INF
        if best <= 50 and f32(best) < f32(f32(best2) * f32(nnratio)):
            if m21[bidx] >= 0:
                m12[m21[bidx]] = -1
                nm -= 1
            m12[i1], m21[bidx], matched_dist[bidx] = bidx, i1, best
            nm += 1
            if check_orientation:
                rot = f32(f1["angles"][i1] - f2["angles"][bidx])
                if rot < 0:
                    rot = f32(rot + f32(360))
                b = int(np.floor(float(f32(rot * f32(1.0 / 30))) + 0.5))
                votes.append((0 if b == 30 else b, i1))
    if check_orientation:
        hist = np.bincount([b for b, _ in votes], minlength=30)
        keep3 = set(int(v) for v in port.three_maxima(hist) if v >= 0)
        for b, i1 in votes:
            if b not in keep3 and m12[i1] >= 0:
                m12[i1] = -1
                nm -= 1
    prev = np.array(prev, np.float32, copy=True)
    for i1 in range(n1):
        if m12[i1] >= 0:
            prev[i1] = f2["kps_xy"][m12[i1]]
    return nm, m12, prev, fallbacks


@pytest.mark.parametrize("crowd,jitter,window,k", [(False, 0.0, 100, 4), (True, 0.0, 100, 4), (False, 6.0, 40, 4), (True, 0.0, 100, 2)])
def test_batched_search_for_initialization_equals_reference(crowd, jitter, window, k):
    """the reference's own ORBmatcher::SearchForInitialization (Tracking::MonocularInitialization; its definition cut out of
    ORBmatcher.cc:648-766) against the batched formulation that the GPU host adapter implements, with the oracle's scan in the place of
    the device scan.  k = 2 forces the fallback path (a list of two is exhausted by one take-over)."""
    from scenes import init_scene
    f1, f2, prev = init_scene(crowd=crowd, jitter=jitter)
    nm_ref, m_ref, prev_ref = ref.search_for_initialization(f1, f2, prev, window, 0.9, True)
    nm, m12, prev_out, fallbacks = _batched_search_for_initialization(f1, f2, prev, window, 0.9, True, port.search_area_best2, k)
    assert nm == nm_ref and np.array_equal(m12, m_ref) and np.array_equal(prev_out, prev_ref)
    assert nm_ref > 20
    if k == 2:
        assert fallbacks > 0
    nm_ref2, m_ref2, _ = ref.search_for_initialization(f1, f2, prev, window, 0.9, False)
    nm2, m122, _, _ = _batched_search_for_initialization(f1, f2, prev, window, 0.9, False, port.search_area_best2, k)
    assert nm2 == nm_ref2 and np.array_equal(m122, m_ref2) and nm_ref2 >= nm_ref

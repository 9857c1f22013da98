"""CPU: pin the OpenCV-free port (oracle/orb_port.cpp) against the real OpenCV primitives (python cv2) and
the known-answer constants of SURVEY.md §8c.  These closed forms are the specification of the CUDA kernels."""
import cv2
import numpy as np
import pytest

from oracle import port
from oracle import orb_ref
from orb_slam3_ros_b200 import synth

cv2.setNumThreads(1)


def test_tables_match_survey():
    e = port.PortExtractor(1000, 1.2, 8, 20, 7)
    assert list(e.features_per_level) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(e.umax) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    np.testing.assert_array_equal(
        e.scale_factors,
        np.array([1, 1.20000005, 1.44000006, 1.72800016, 2.07360029, 2.48832035, 2.98598456, 3.58318162], np.float32))
    assert list(port.PortExtractor(2000, 1.2, 8, 20, 7).features_per_level) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert list(port.PortExtractor(5000, 1.2, 8, 20, 7).features_per_level) == [1086, 905, 754, 628, 524, 436, 364, 303]
    assert list(port.PortExtractor(8000, 1.2, 12, 20, 7).features_per_level) == [1502, 1251, 1043, 869, 724, 604, 503, 419,
                                                                                   349, 291, 243, 202]


def test_cv_round_half_even():
    got = [port.lib().port_cv_round(v) for v in (0.5, 1.5, 2.5, -0.5, -1.5)]
    assert got == [0, 2, 2, 0, -2]


@pytest.mark.parametrize("shape", [(480, 752), (376, 1241), (480, 640), (97, 131)])
def test_resize_chain_matches_cv2(shape):
    h, w = shape
    img = synth.frame(h, w, 3)
    e = port.PortExtractor(1000, 1.2, 8, 20, 7)
    cur = img
    for l in range(1, 8):
        s = np.float32(e.inv_scale_factors[l])
        lw = port.lib().port_cv_round(float(np.float32(w) * s))
        lh = port.lib().port_cv_round(float(np.float32(h) * s))
        want = cv2.resize(cur, (lw, lh), interpolation=cv2.INTER_LINEAR)
        got = port.resize_linear(cur, lw, lh)
        assert np.array_equal(want, got), f"level {l}"
        cur = want


def test_resize_random_sizes_matches_cv2():
    rng = np.random.default_rng(5)
    for _ in range(40):
        sh, sw = int(rng.integers(8, 200)), int(rng.integers(8, 200))
        dh, dw = int(rng.integers(4, 260)), int(rng.integers(4, 260))   # includes up-scaling (clamped taps)
        src = rng.integers(0, 256, size=(sh, sw), dtype=np.uint8)
        if sw == 2 * dw and sh == 2 * dh:
            continue  # cv::resize silently switches INTER_LINEAR to INTER_AREA for exact 2x decimation
        assert np.array_equal(cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR), port.resize_linear(src, dw, dh))


def test_fast9_matches_cv2():
    rng = np.random.default_rng(11)
    for trial in range(60):
        h, w = int(rng.integers(7, 60)), int(rng.integers(7, 60))
        if trial % 3 == 0:
            img = rng.integers(0, 256, size=(h, w), dtype=np.uint8)       # noise: dense, adjacent equal scores
        elif trial % 3 == 1:
            img = synth.frame(max(h, 16), max(w, 16), trial)[:h, :w].copy()
        else:
            img = (rng.integers(0, 4, size=(h, w)) * 60).astype(np.uint8)  # plateaus: ties kill each other
        for th in (7, 20, 0, 100):
            det = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
            want = np.array([(k.pt[0], k.pt[1], k.response) for k in det.detect(img, None)], np.float32).reshape(-1, 3)
            got = port.fast9(img, th)
            assert np.array_equal(want, got), (trial, th, len(want), len(got))


def test_fast9_known_corner():
    img = np.full((7, 7), 100, np.uint8)
    img[3, 3] = 200
    got = port.fast9(img, 20)
    assert got.tolist() == [[3.0, 3.0, 99.0]]          # M = 100 -> response = M - 1


def test_gaussian7_matches_cv2():
    rng = np.random.default_rng(2)
    for trial in range(12):
        h, w = int(rng.integers(8, 150)), int(rng.integers(8, 150))
        img = rng.integers(0, 256, size=(h, w), dtype=np.uint8) if trial % 2 else synth.frame(h + 8, w + 8, trial)[:h, :w].copy()
        want = cv2.GaussianBlur(img, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
        assert np.array_equal(want, port.gaussian7(img))
    imp = np.zeros((15, 15), np.uint8)
    imp[7, 7] = 255
    assert np.array_equal(cv2.GaussianBlur(imp, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101), port.gaussian7(imp))


def test_fast_atan2_matches_cv2():
    assert port.fast_atan2(np.float32([1]), np.float32([2]))[0] == np.float32(26.56710433959961)
    assert port.fast_atan2(np.float32([0]), np.float32([0]))[0] == 0
    rng = np.random.default_rng(4)
    y = np.concatenate([rng.integers(-200000, 200000, 20000), [0, 0, 5, -5, 1, -1]]).astype(np.float32)
    x = np.concatenate([rng.integers(-200000, 200000, 20000), [7, -7, 0, 0, 1, -1]]).astype(np.float32)
    want = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(want, port.fast_atan2(y, x))


def test_hamming_known_answers():
    z = np.zeros(32, np.uint8)
    f = np.full(32, 255, np.uint8)
    assert port.hamming(z, f) == 256 and port.hamming(z, z) == 0
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 256, (2, 32), dtype=np.uint8)
    assert port.hamming(a, b) == int(np.unpackbits(a ^ b).sum())


def test_knn2_tie_rule_and_cv2():
    # distances 3,1,1,2,1,1 -> (idx 1, idx 2): stable ascending = lowest train index first
    q = np.zeros((1, 32), np.uint8)
    tr = np.zeros((6, 32), np.uint8)
    for i, d in enumerate([3, 1, 1, 2, 1, 1]):
        tr[i, 0] = (1 << d) - 1
    idx, dist = port.knn2(q, tr)
    assert idx.tolist() == [[1, 2]] and dist.tolist() == [[1, 1]]
    rng = np.random.default_rng(9)
    db, qq = synth.descriptor_db(3000, 200, seed=3, dup_every=97)
    i1, d1 = port.knn2(qq, db, nthreads=2)
    i2, d2 = orb_ref.bf_knn2(qq, db)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    # fewer than two train rows -> shorter list
    i3, d3 = port.knn2(qq[:4], db[:1])
    i4, d4 = orb_ref.bf_knn2(qq[:4], db[:1])
    assert np.array_equal(i3, i4) and np.array_equal(d3, d4) and (i3[:, 1] == -1).all()


def test_matcher_constants():
    # ORBmatcher.cc:35-37, Frame.cc:816
    from orb_slam3_ros_b200 import constants as c
    assert (c.TH_LOW, c.TH_HIGH, c.HISTO_LENGTH, (c.TH_HIGH + c.TH_LOW) // 2) == (50, 100, 30, 75)


def test_gray_matches_cv2_cvtcolor():
    """Tracking.cc:1498-1525 converts colour input with cv::cvtColor before building a Frame ("next" row)."""
    rng = np.random.default_rng(12)
    for c, rgb_code, bgr_code in ((3, cv2.COLOR_RGB2GRAY, cv2.COLOR_BGR2GRAY), (4, cv2.COLOR_RGBA2GRAY, cv2.COLOR_BGRA2GRAY)):
        img = rng.integers(0, 256, (61, 83, c), dtype=np.uint8)
        assert np.array_equal(port.gray(img, True), cv2.cvtColor(img, rgb_code))
        assert np.array_equal(port.gray(img, False), cv2.cvtColor(img, bgr_code))
    ramp = np.stack(np.meshgrid(np.arange(256), np.arange(256)), -1).astype(np.uint8)
    ramp = np.concatenate([ramp, (ramp[..., :1] // 2 + ramp[..., 1:] // 3)], -1)
    assert np.array_equal(port.gray(ramp, True), cv2.cvtColor(ramp, cv2.COLOR_RGB2GRAY))


def test_distinctive_descriptor_reference_semantics():
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:329-403) restated in numpy vs the port"""
    rng = np.random.default_rng(13)
    sizes = [1, 2, 3, 4, 7, 20, 33, 64, 0, 5]
    rowptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    base = rng.integers(0, 256, (len(sizes), 32), dtype=np.uint8)
    desc = np.zeros((rowptr[-1], 32), np.uint8)
    for g, n in enumerate(sizes):
        for i in range(n):
            d = base[g].copy()
            flips = rng.integers(0, 256, rng.integers(0, 30))
            bits = np.unpackbits(d)
            bits[flips] ^= 1
            desc[rowptr[g] + i] = np.packbits(bits)
    got = port.distinctive(desc, rowptr)
    for g, n in enumerate(sizes):
        if n == 0:
            assert got[g] == -1
            continue
        d = desc[rowptr[g]:rowptr[g + 1]]
        dist = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(2)
        med = np.sort(dist, axis=1)[:, int(0.5 * (n - 1))]
        assert got[g] == int(np.argmin(med))          # first minimum


def test_bow_oracle_against_plain_python():
    """oracle/port.bow_transform (std::map restatement of DBoW2 transform) against a dict-based Python walk"""
    from orb_slam3_ros_b200.bow import synthetic_vocabulary
    vocab = synthetic_vocabulary(5, 3, seed=9, ragged=True)
    rng = np.random.default_rng(4)
    desc = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    desc[::3] = vocab["node_desc"][rng.integers(1, len(vocab["child_begin"]), 100)]
    levelsup = 1
    bow, fv = {}, {}
    for f, d in enumerate(desc):
        node, level, nid = 0, 0, 0
        while vocab["child_count"][node] > 0:
            level += 1
            b, c = vocab["child_begin"][node], vocab["child_count"][node]
            ids = vocab["child_list"][b:b + c]
            dist = np.unpackbits(vocab["node_desc"][ids] ^ d, axis=1).sum(1)
            node = int(ids[int(np.argmin(dist))])           # argmin = first minimum
            if level == vocab["depth"] - levelsup:
                nid = node
        w = float(vocab["node_weight"][node])
        if w > 0:
            bow[int(vocab["node_word"][node])] = bow.get(int(vocab["node_word"][node]), 0.0) + w
            fv.setdefault(nid, []).append(f)
    ids = sorted(bow)
    nrm = 0.0
    for i in ids:
        nrm += abs(bow[i])
    got = port.bow_transform(vocab, desc, levelsup, 1)
    assert got[0].tolist() == ids
    assert got[1].tolist() == [bow[i] / nrm for i in ids]
    assert got[2].tolist() == sorted(fv)
    assert got[4].tolist() == [f for n in sorted(fv) for f in fv[n]]
    assert got[5] == sum(len(v) for v in fv.values())


def _rectify_maps(h, w, seed=0):
    """maps of a mild radial + rotational warp (the shape initUndistortRectifyMap produces), partly leaving the source"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    cx, cy = np.float32(w / 2 + 3.3), np.float32(h / 2 - 2.1)
    r2 = ((xx - cx) ** 2 + (yy - cy) ** 2) / np.float32(w * w)
    th = np.float32(0.02)
    mx = cx + (xx - cx) * (1 + np.float32(0.25) * r2) * np.cos(th) - (yy - cy) * np.sin(th) + np.float32(1.7)
    my = cy + (xx - cx) * np.sin(th) + (yy - cy) * (1 + np.float32(0.25) * r2) * np.cos(th) - np.float32(0.9)
    mx[::7, ::5] = np.round(mx[::7, ::5])                     # exact integer positions (the {32767,0,0,1} table entry)
    my[::7, ::5] = np.round(my[::7, ::5])
    mx[:3] -= w                                               # rows mapped far outside
    my[:, -2:] += np.float32(0.5) + h
    mx[5, :40] = np.linspace(-1.5, 0.5, 40, dtype=np.float32)     # straddling the left / top border
    my[6, :40] = np.linspace(-1.5, 0.5, 40, dtype=np.float32)
    return mx.astype(np.float32), my.astype(np.float32)


def test_remap_oracle_matches_cv2():
    """oracle/port.remap_linear (restatement of cv::remap INTER_LINEAR, System.cc:239) against the real cv2.remap"""
    cv2 = pytest.importorskip("cv2")
    for (h, w, seed) in ((120, 160, 0), (376, 1241, 1), (97, 131, 2)):
        src = synth.frame(h + 16, w + 24, seed)
        mx, my = _rectify_maps(h, w, seed)
        want = cv2.remap(src, mx, my, cv2.INTER_LINEAR)
        assert np.array_equal(port.remap_linear(src, mx, my), want), (h, w)


def test_undistort_oracle_matches_cv2():
    """oracle/port.undistort_points against cv2.undistortPoints (Frame::UndistortKeyPoints, Frame.cc:766), bit for bit"""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    xy = np.stack([rng.uniform(0, 752, 4000), rng.uniform(0, 480, 4000)], 1).astype(np.float32)
    K4 = (458.654, 457.296, 367.215, 248.375)                                    # config/.../EuRoC.yaml
    for dist in ([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05],         # EuRoC cam0
                 [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.011],   # with k3
                 [0.35, -0.6, 0.01, -0.02, 0.4]):                                # strong, some points diverge
        K = np.float32([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]])
        D = np.float32(dist).reshape(-1, 1)
        want = cv2.undistortPoints(xy.reshape(-1, 1, 2).copy(), K, D, None, K).reshape(-1, 2)
        got = port.undistort_points(xy, K4, dist)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), dist
    assert np.array_equal(port.undistort_points(xy, K4, [0.0, 0.1, 0, 0]), xy)   # Frame.cc:749: k1 == 0 -> copy

"""CPU: whole-extractor agreement of the two oracles -- the OpenCV-free port vs the cv2-backed restatement -- and of
both with the committed golden fixtures; stereo / best-2 sanity of the port."""
import zlib
from pathlib import Path

import cv2
import numpy as np
import pytest

from oracle import orb_ref, port
from orb_slam3_ros_b200 import synth

GOLD = Path(__file__).resolve().parent / "golden"


def _same(k0, d0, m0, k1, d1, m1):
    assert len(k0) == len(k1) and m0 == m1
    for f in ("x", "y", "size", "angle", "response", "octave"):
        assert np.array_equal(k0[f], k1[f]), f
    assert np.array_equal(d0, d1)


@pytest.mark.parametrize("shape,nf,nl,lap", [((480, 752), 1000, 8, (0, 1000)), ((376, 1241), 2000, 8, (0, 0)),
                                             ((240, 320), 300, 4, (100, 200))])
def test_port_equals_cv2_reference(shape, nf, nl, lap):
    img = synth.frame(shape[0], shape[1], 1)
    pe, re_ = port.PortExtractor(nf, 1.2, nl), orb_ref.RefExtractor(nf, 1.2, nl)
    rc, k0, d0, m0 = pe.extract(img, lap)
    rc2, k1, d1, m1 = re_.extract(img, lap)
    assert rc == 0 and rc2 == 0
    for l in range(nl):
        assert np.array_equal(pe.level(l, bordered=True), re_.pyramid[l])
        assert np.array_equal(pe.raw_keys(l), re_.raw[l])
    _same(k0, d0, m0, k1, d1, m1)


@pytest.mark.parametrize("name", ["mono_320x240", "wide_400x200", "noise_176x144", "fallback_260x200"])
def test_port_reproduces_golden(name):
    g = np.load(GOLD / f"{name}.npz")
    nf, nl, ini, mn, l0, l1 = [int(v) for v in g["params"]]
    pe = port.PortExtractor(nf, 1.2, nl, ini, mn)
    rc, k, d, m = pe.extract(g["image"], (l0, l1))
    assert rc == 0
    for l in range(nl):
        assert zlib.crc32(pe.level(l, bordered=True).tobytes()) == int(g["pyr_crc"][l])
        assert len(pe.raw_keys(l)) == int(g["raw_n"][l])
    _same(g["kps"].view(port.KP_DTYPE).reshape(-1), g["desc"], int(g["mono"]), k, d, m)


def test_golden_still_matches_installed_cv2():
    """canary: the committed fixtures were produced with the recorded cv2 version; the installed one must agree."""
    g = np.load(GOLD / "mono_320x240.npz")
    nf, nl, ini, mn, l0, l1 = [int(v) for v in g["params"]]
    rc, k, d, m = orb_ref.RefExtractor(nf, 1.2, nl, ini, mn).extract(g["image"], (l0, l1))
    _same(g["kps"].view(port.KP_DTYPE).reshape(-1), g["desc"], int(g["mono"]), k, d, m)
    gk = np.load(GOLD / "knn_300x4000.npz")
    i, dd = orb_ref.bf_knn2(gk["q"], gk["db"])
    assert np.array_equal(i, gk["idx"]) and np.array_equal(dd, gk["dist"]), (str(gk["cv2_version"]), cv2.__version__)


def test_fma_build_variance_is_within_tolerance():
    """The reference builds with GCC's default -ffp-contract=fast (-O3 -march=native): report how many descriptor bits
    that changes relative to the un-fused truth -- the reference's own build-to-build variance."""
    img = synth.frame(480, 752, 2)
    _, k0, d0, _ = port.PortExtractor().extract(img)
    _, k1, d1, _ = port.PortExtractor(fma=True).extract(img)
    assert len(k0) == len(k1)
    assert np.abs(k0["angle"] - k1["angle"]).max() <= 1e-3
    assert np.unpackbits(d0 ^ d1).sum() <= 1e-4 * d0.size * 8


def test_port_stereo_and_best2_sanity():
    left, right = synth.stereo_pair(240, 400, 0, dmax=40)
    pl, pr = port.PortExtractor(600, 1.2, 6), port.PortExtractor(600, 1.2, 6)
    _, kl, dl, _ = pl.extract(left)
    _, kr, dr, _ = pr.extract(right)
    ur, dp, br, sad, kept = port.stereo(pl, pr, kl, dl, kr, dr, np.float32(380.0), np.float32(0.5))
    ok = ur >= 0
    assert ok.sum() == kept and kept > 0.25 * len(kl)
    assert (ur[ok] <= kl["x"][ok]).all() and (dp[ok] > 0).all()
    # best-2 over explicit candidate lists == brute force restricted to the list
    rng = np.random.default_rng(0)
    rowptr = np.arange(0, 50 * 21, 50, dtype=np.int32)
    cand = rng.integers(0, len(dr), rowptr[-1]).astype(np.int32)
    out = port.best2_csr(dl[:20], dr, cand, rowptr, 256)
    for i in range(20):
        ds = [port.hamming(dl[i], dr[c]) for c in cand[rowptr[i]:rowptr[i + 1]]]
        assert out[i, 0] == min(ds) and out[i, 1] == cand[rowptr[i] + int(np.argmin(ds))]


@pytest.mark.parametrize("sf,nl", [(1.5, 4), (2.0, 3), (1.1, 6)])
def test_port_equals_cv2_reference_other_scale_factors(sf, nl):
    img = synth.frame(480, 640, 1)
    pe, re_ = port.PortExtractor(500, sf, nl), orb_ref.RefExtractor(500, sf, nl)
    rc, k0, d0, m0 = pe.extract(img)
    rc2, k1, d1, m1 = re_.extract(img)
    for l in range(nl):
        assert np.array_equal(pe.level(l, bordered=True), re_.pyramid[l])
    _same(k0, d0, m0, k1, d1, m1)


def test_port_rejects_degenerate_pyramids_without_work():
    """levels that shrink to nothing (scale 2.0 x 10 levels on a 236x192 frame: level 8 would be 1x1, level 9 0x0) -- cv::resize
    asserts dsize.area() > 0 in the reference; the port must say "unsupported" at once, not loop"""
    assert port.PortExtractor(90, 2.0, 10, 17, 16).extract(synth.frame(236, 192, 1))[0] == -2
    assert port.PortExtractor(100, 1.2, 8).extract(synth.frame(90, 120, 0))[0] == -2          # level 7 = 33x25: inside the FAST border


def test_cv2_baseline_equals_port():
    """oracle/cv2_baseline.py (bench.py's honest CPU baseline: the extractor's control flow over python-cv2's SIMD resize /
    FAST / GaussianBlur + the port's C++ glue) gives bit for bit the port's result, so the two CPU arms time the same work"""
    from oracle import cv2_baseline
    for (h, w, nf, nl, lap, seed) in [(240, 320, 300, 4, (0, 1000), 1), (200, 333, 500, 5, (100, 220), 2), (480, 752, 1000, 8, (0, 0), 3)]:
        img = synth.frame(h, w, seed)
        rc0, k0, d0, m0 = port.PortExtractor(nf, 1.2, nl).extract(img, lap)
        rc1, k1, d1, m1 = cv2_baseline.Cv2Extractor(nf, 1.2, nl).extract(img, lap)
        assert rc0 == rc1 == 0 and m0 == m1 and len(k0) == len(k1)
        for f in k0.dtype.names:
            assert np.array_equal(k0[f], k1[f]), f
        assert np.array_equal(d0, d1)
    a, b = cv2_baseline.rates(synth.sequence(240, 320, 4, canvas=512), (300, 1.2, 4, 20, 7), processes=2)
    assert a > 0 and b >= a * 0.5

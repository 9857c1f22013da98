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

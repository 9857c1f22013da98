"""GPU parity tests of the CUDA ORB extractor, through the C ABI (orbb_* via ctypes), against
  (a) the oracle port (oracle/orb_port.cpp, pinned to python-cv2 by tests/test_oracle_*.py) on seeded inputs, and
  (b) the committed golden fixtures produced with the cv2-backed oracle (tools/make_golden.py), and
  (c) the reference's own ORBextractor.cc compiled unmodified (oracle/_ref/liborbref.so, see oracle/cvshim/cvshim.hpp) whenever that
      library is present (it is built in the development container and travels to the GPU box with the other built files).
Bar (BASELINE.json north_star): pyramid pixels, keypoints (pt, octave, size, response), their order and the mono index
bit-exact; angles within 1e-3 deg; descriptor bit disagreement <= 1e-4.
"""
import zlib
from pathlib import Path

import numpy as np
import pytest

from oracle import port, ref
from orb_slam3_ros_b200 import capi, synth
from orb_slam3_ros_b200.extractor import ORBextractor

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"
ANGLE_TOL_DEG = 1e-3
DESC_BIT_TOL = 1e-4


def assert_same_features(k0, d0, m0, k1, d1, m1, ctx=""):
    assert len(k0) == len(k1), f"{ctx}: keypoint count {len(k0)} vs {len(k1)}"
    assert m0 == m1, f"{ctx}: mono index {m0} vs {m1}"
    for f in ("x", "y", "size", "response", "octave"):
        assert np.array_equal(k0[f], k1[f]), f"{ctx}: field {f} differs"
    if len(k0):
        assert np.abs(k0["angle"] - k1["angle"]).max() <= ANGLE_TOL_DEG, ctx
        assert np.unpackbits(d0 ^ d1).sum() <= DESC_BIT_TOL * d0.size * 8, ctx


def check_against_port(img, nf, nl, ini=20, mn=7, lap=(0, 0), stages=True):
    pe = port.PortExtractor(nf, 1.2, nl, ini, mn)
    rc, k0, d0, m0 = pe.extract(img, lap)
    assert rc == 0
    ge = ORBextractor(nf, 1.2, nl, ini, mn)
    m1, k1, d1 = ge(img, None, lap)
    if stages:
        for l in range(nl):
            assert np.array_equal(pe.level(l, bordered=True), ge.debug_level(0, l, bordered=True)), f"pyramid level {l}"
            assert np.array_equal(pe.raw_keys(l), ge.debug_raw_keys(0, l)), f"vToDistributeKeys level {l}"
            s = pe.selected(l)
            want = np.stack([s["x"], s["y"], s["response"]], 1) if len(s) else np.zeros((0, 3), np.float32)
            assert np.array_equal(want, ge.debug_selected(0, l)), f"DistributeOctTree level {l}"
            if len(s):
                assert np.array_equal(pe.level(l, blurred=True), ge.debug_level(0, l, blurred=True)), f"blur level {l}"
    assert_same_features(k0, d0, m0, k1, d1, m1)
    if ref.available():      # the reference's object code on the same input
        rc, kr, dr, mr = ref.RefExtractor(nf, 1.2, nl, ini, mn).extract(img, lap)
        assert rc == 0
        assert_same_features(kr, dr, mr, k1, d1, m1, "vs the reference's ORBextractor.cc")
    return ge, k1, d1


@pytest.mark.parametrize("shape,nf,nl,lap", [
    ((480, 752), 1000, 8, (0, 1000)),        # BASELINE config 1: EuRoC mono (Frame.cc:311 passes {0,1000})
    ((376, 1241), 2000, 8, (0, 0)),          # config 2 shape: KITTI stereo eye
    ((480, 640), 1000, 8, (0, 0)),           # config 3 shape: TUM RGB-D
    ((480, 752), 5000, 8, (0, 1000)),        # monocular initialisation extractor (Tracking.cc:637: 5 x nFeatures)
    ((134, 210), 100, 3, (0, 0)),
    ((720, 1280), 3000, 12, (300, 900)),     # 12 levels, fisheye-style lapping split
])
def test_extract_matches_oracle(shape, nf, nl, lap):
    check_against_port(synth.frame(shape[0], shape[1], nf % 17), nf, nl, lap=lap)


def test_4k_12_levels_matches_oracle():
    # BASELINE config 5: 3840x2160, 8000 features, 12 levels
    check_against_port(synth.frame(2160, 3840, 5), 8000, 12, lap=(0, 0), stages=False)


@pytest.mark.parametrize("name", ["mono_320x240", "wide_400x200", "noise_176x144", "fallback_260x200"])
def test_golden_fixture(name):
    g = np.load(GOLD / f"{name}.npz")
    nf, nl, ini, mn, l0, l1 = [int(v) for v in g["params"]]
    ge = ORBextractor(nf, 1.2, nl, ini, mn)
    m1, k1, d1 = ge(g["image"], None, (l0, l1))
    for l in range(nl):
        assert zlib.crc32(ge.debug_level(0, l, bordered=True).tobytes()) == int(g["pyr_crc"][l]), f"pyramid level {l}"
        assert len(ge.debug_raw_keys(0, l)) == int(g["raw_n"][l]), f"raw key count level {l}"
        if int(g["blur_crc"][l]):
            assert zlib.crc32(ge.debug_level(0, l, blurred=True).tobytes()) == int(g["blur_crc"][l]), f"blur level {l}"
    sel = np.concatenate([np.concatenate([ge.debug_selected(0, l), np.full((len(ge.debug_selected(0, l)), 1), l, np.float32)], 1)
                          for l in range(nl)])
    assert np.array_equal(sel, g["sel"])
    assert_same_features(g["kps"].view(capi.KP_DTYPE).reshape(-1), g["desc"], int(g["mono"]), k1, d1, m1, name)


def test_empty_image_returns_minus_one():
    ge = ORBextractor()
    mono, k, d = ge(np.zeros((0, 0), np.uint8))
    assert mono == -1 and len(k) == 0 and d.shape == (0, 32)              # ORBextractor.cc:1090-1091
    lib = capi.load()
    import ctypes as C
    n, m = C.c_int(), C.c_int()
    assert lib.orbb_extract(ge._h, None, 0, 0, 0, 0, 0, None, None, 0, C.byref(n), C.byref(m)) == capi.ORBB_ERR_EMPTY


def test_flat_image_has_no_keypoints():
    ge = ORBextractor(500, 1.2, 4)
    mono, k, d = ge(np.full((200, 300), 90, np.uint8))
    assert mono == 0 and len(k) == 0 and d.shape == (0, 32)               # _descriptors.release() :1108-1109
    assert port.PortExtractor(500, 1.2, 4).extract(np.full((200, 300), 90, np.uint8))[1].size == 0


def test_strided_input_and_handle_reuse_across_sizes():
    big = synth.frame(300, 500, 2)
    view = big[10:250, 20:420]                      # non-contiguous rows
    ge, k1, d1 = check_against_port(np.ascontiguousarray(view), 400, 5)
    m2, k2, d2 = ge(view, None, (0, 0))             # same pixels through a strided view
    assert np.array_equal(k1, k2) and np.array_equal(d1, d2)
    img2 = synth.frame(240, 320, 3)                 # the plan is rebuilt for a new image size
    m3, k3, d3 = ge(img2, None, (0, 0))
    rc, k0, d0, m0 = port.PortExtractor(400, 1.2, 5).extract(img2)
    assert_same_features(k0, d0, m0, k3, d3, m3, "after resize")


def test_noise_image_dense_corners_and_tie_heavy_quadtree():
    rng = np.random.default_rng(8)
    check_against_port(rng.integers(0, 256, (240, 320), dtype=np.uint8), 600, 5, lap=(100, 220))
    check_against_port((rng.integers(0, 3, (200, 280)) * 100).astype(np.uint8), 300, 4)       # plateaus


def test_thresholds_and_small_feature_budget():
    img = synth.frame(240, 320, 9)
    check_against_port(img, 40, 6, ini=12, mn=7)    # KITTI04-12 uses iniThFAST 12; N per level < 4*nIni at the top levels
    check_against_port(img, 300, 4, ini=40, mn=3)


def test_too_small_level_is_rejected():
    ge = ORBextractor(100, 1.2, 8)
    with pytest.raises(capi.OrbbError) as e:
        ge(synth.frame(90, 120, 0))                 # level 7 would be 33x25: a negative row span, the reference's root count (:559) is undefined
    assert e.value.code == capi.ORBB_ERR_UNSUPPORTED
    with pytest.raises(capi.OrbbError) as e:
        ORBextractor(90, 2.0, 10, 17, 16)(synth.frame(236, 192, 1))      # levels shrink to 1x1 and 0x0: cv::resize asserts in the reference
    assert e.value.code == capi.ORBB_ERR_UNSUPPORTED


@pytest.mark.parametrize("shape,nf,nl", [((200, 260), 500, 8), ((150, 400), 300, 8), ((120, 160), 200, 6)])
def test_levels_smaller_than_one_cell_contribute_nothing(shape, nf, nl):
    """nCols or nRows == 0 on the top levels: the reference's cell loops do not run and the level stays empty (its image is
    still built); checked against the port and the reference's own object code."""
    ge, k, d = check_against_port(synth.frame(shape[0], shape[1], 3), nf, nl, lap=(0, 100))
    assert len(k) > 0 and k["octave"].max() < nl - 1


def test_pyramid_accessor_matches_mvImagePyramid():
    img = synth.frame(240, 320, 4)
    pe = port.PortExtractor(300, 1.2, 4)
    pe.extract(img)
    ge = ORBextractor(300, 1.2, 4)
    ge(img)
    for l in range(4):
        assert np.array_equal(ge.image_pyramid(l), pe.level(l))
        assert np.array_equal(ge.image_pyramid(l, with_border=True), pe.level(l, bordered=True))
    assert np.array_equal(ge.GetScaleFactors(), pe.scale_factors)
    assert np.array_equal(ge.GetInverseScaleFactors(), pe.inv_scale_factors)
    assert np.array_equal(ge.GetScaleSigmaSquares(), pe.level_sigma2)
    assert np.array_equal(ge.GetInverseScaleSigmaSquares(), pe.inv_level_sigma2)
    assert np.array_equal(ge.features_per_level, pe.features_per_level)


def test_batch_equals_single_and_host_equals_device_path():
    import torch
    NF = 9                                          # >= 8 frames: the host path pipelines two chunks (5 + 4)
    frames = synth.sequence(240, 320, NF, canvas=512)
    ge = ORBextractor(300, 1.2, 4, max_batch=NF)
    counts, kps, desc = ge.extract_batch_host(frames, (0, 1000))
    single = ORBextractor(300, 1.2, 4)
    for f in range(NF):
        m1, k1, d1 = single(frames[f], None, (0, 1000))
        n = counts[f, 0]
        assert n == len(k1) and counts[f, 1] == m1
        assert np.array_equal(kps[f, :n], k1) and np.array_equal(desc[f, :n], d1)
    dev = torch.from_numpy(frames).cuda()
    ge.extract_batch_device(dev, NF, 320, 240, lapping=(0, 1000))
    c2, k2, d2 = ge.fetch(NF)
    assert np.array_equal(c2, counts)
    for f in range(NF):
        n = counts[f, 0]
        assert np.array_equal(k2[f, :n], kps[f, :n]) and np.array_equal(d2[f, :n], desc[f, :n])


def test_full_size_batch_properties():
    """BASELINE-size batch (256 x 752x480): size-independent properties, and ALL 256 frames against the oracle."""
    import torch
    B = 256
    frames = synth.sequence(480, 752, B)
    ge = ORBextractor(1000, 1.2, 8, max_batch=B)
    dev = torch.from_numpy(frames).cuda()
    ge.extract_batch_device(dev, B, 752, 480, lapping=(0, 1000))
    c1, k1, d1 = ge.fetch(B)
    ge.extract_batch_device(dev, B, 752, 480, lapping=(0, 1000))
    c2, k2, d2 = ge.fetch(B)
    assert np.array_equal(c1, c2) and np.array_equal(k1, k2) and np.array_equal(d1, d2)       # deterministic / idempotent
    npl = ge.features_per_level
    for f in range(B):
        n = c1[f, 0]
        assert 0 < n <= 1000 + 2 * 8 and c1[f, 1] == 0                   # width <= 1000: every keypoint is "lapping"
        k = k1[f, :n]
        assert (k["x"] >= 19).all() and (k["x"] <= 752 - 19).all() and (k["y"] >= 19).all() and (k["y"] <= 480 - 19).all()
        assert (np.diff(k["octave"]) <= 0).all()                         # written from the back: octaves descend
        per =np.bincount(k["octave"], minlength=8)
        assert (per <= npl + 2).all()
        assert ((k["angle"] >= 0) & (k["angle"] < 360)).all()
    # every frame of the batch against the oracle port (all host threads)
    import os
    c0, k0, d0 = port.extract_batch(frames, lapping=(0, 1000), nthreads=max(1, os.cpu_count() or 1))
    assert np.array_equal(c0, c1)
    for f in range(B):
        n = c1[f, 0]
        assert_same_features(k0[f, :n], d0[f, :n], int(c0[f, 1]), k1[f, :n], d1[f, :n], int(c1[f, 1]), f"frame {f}")


def test_submit_wait_on_two_handles_equals_sync_call():
    """orbb_extract_batch_host_submit/_wait: two handles overlapping their batches give the results of the sync call"""
    NF = 12
    frames = [synth.sequence(240, 320, NF, canvas=512, base_seed=100 + i) for i in range(3)]
    a, b = ORBextractor(300, 1.2, 4, max_batch=NF), ORBextractor(300, 1.2, 4, max_batch=NF)
    cap = a.max_keypoints

    def bufs():
        return (np.zeros((NF, cap), capi.KP_DTYPE), np.zeros((NF, cap, 32), np.uint8), np.zeros((NF, 2), np.int32))

    want = [a.extract_batch_host(f, (0, 1000), out=bufs()) for f in frames]
    outs = [bufs() for _ in frames]
    a.submit_batch_host(frames[0], (0, 1000), out=outs[0])
    b.submit_batch_host(frames[1], (0, 1000), out=outs[1])
    a.wait_batch_host()
    a.submit_batch_host(frames[2], (0, 1000), out=outs[2])
    b.wait_batch_host()
    a.wait_batch_host()
    for (c0, k0, d0), (k1, d1, c1) in zip(want, outs):
        assert np.array_equal(c0, c1)
        for f in range(NF):
            n = c0[f, 0]
            assert np.array_equal(k0[f, :n], k1[f, :n]) and np.array_equal(d0[f, :n], d1[f, :n])


@pytest.mark.parametrize("sf,nl", [(1.5, 4), (2.0, 3), (1.1, 6)])
def test_other_scale_factors(sf, nl):
    """scaleFactor is a user setting (ORBextractor.scaleFactor in the YAML files); 2.0 is the exact-decimation case in
    which cv::resize silently switches INTER_LINEAR to INTER_AREA (same numbers as the fixed-point bilinear path)."""
    img = synth.frame(480, 640, 1)
    pe = port.PortExtractor(500, sf, nl)
    rc, k0, d0, m0 = pe.extract(img)
    ge = ORBextractor(500, sf, nl)
    m1, k1, d1 = ge(img)
    for l in range(nl):
        assert np.array_equal(pe.level(l, bordered=True), ge.debug_level(0, l, bordered=True)), l
    assert_same_features(k0, d0, m0, k1, d1, m1, f"scale {sf}")
    assert np.array_equal(ge.GetScaleFactors(), pe.scale_factors)


def test_frame_partition_is_result_invariant():
    """BASELINE config 3 semantics at reduced size: a TUM-shape sequence partitioned over ranks (contiguous blocks, no
    collective) gives, frame for frame, what one extractor gives for the whole sequence."""
    import torch
    from orb_slam3_ros_b200 import sharding
    NF, world = 24, 3
    frames = synth.sequence(480, 640, NF, canvas=1024)
    whole = ORBextractor(1000, 1.2, 8, max_batch=NF)
    whole.extract_batch_device(torch.from_numpy(frames).cuda(), NF, 640, 480)
    c0, k0, d0 = whole.fetch(NF)
    for rank in range(world):
        lo, hi = sharding.block_bounds(NF, world, rank)
        part = ORBextractor(1000, 1.2, 8, max_batch=hi - lo)
        part.extract_batch_device(torch.from_numpy(frames[lo:hi]).cuda(), hi - lo, 640, 480)
        c1, k1, d1 = part.fetch(hi - lo)
        assert np.array_equal(c0[lo:hi], c1)
        for f in range(hi - lo):
            n = c1[f, 0]
            assert np.array_equal(k0[lo + f, :n], k1[f, :n]) and np.array_equal(d0[lo + f, :n], d1[f, :n])


def test_color_input_matches_cvtcolor_then_extract():
    """Tracking.cc:1498-1525 ("next" row): colour frames are converted with cv::cvtColor(..2GRAY) before extraction; the
    device conversion + extraction must equal the oracle's conversion + the gray extraction."""
    rng = np.random.default_rng(4)
    gray = synth.frame(240, 320, 6).astype(np.int32)
    for channels in (3, 4):
        for rgb in (True, False):
            img = np.clip(np.stack([gray + rng.integers(-20, 21, gray.shape) for _ in range(channels)], -1), 0, 255).astype(np.uint8)
            g = port.gray(img, rgb)
            rc, k0, d0, m0 = port.PortExtractor(300, 1.2, 4).extract(g, (0, 1000))
            ge = ORBextractor(300, 1.2, 4)
            m1, k1, d1 = ge.extract_color(img, rgb, (0, 1000))
            assert np.array_equal(ge.debug_level(0, 0), g), (channels, rgb)
            assert_same_features(k0, d0, m0, k1, d1, m1, f"channels={channels} rgb={rgb}")


def test_random_geometries_match_oracle():
    """seeded sweep over odd image sizes / level counts / budgets: every stage (bordered pyramid, raw FAST keys, quadtree
    selection, blur) and the final features against the oracle -- exercises the alignment-dependent paths (TMA row staging,
    16-byte apron chunks, word-wise resize) on widths that are not multiples of 4 or 16"""
    rng = np.random.default_rng(2024)
    sizes = [(131, 257), (203, 333), (97, 401), (311, 190), (480, 641), (255, 1000), (150, 150), (402, 599)]
    for i, (h, w) in enumerate(sizes):
        nl = int(rng.integers(2, 9))
        while min(h, w) / 1.2 ** (nl - 1) < 80:            # every level must hold at least one 35-px cell row and column
            nl -= 1
        nf = int(rng.integers(150, 1600))
        lap = (0, 0) if i % 2 else (int(w * 0.3), int(w * 0.6))
        check_against_port(synth.frame(h, w, 40 + i), nf, nl, lap=lap)


def test_single_frame_graph_is_rebuilt_when_the_call_changes():
    """the single-frame path replays a captured CUDA graph; changing the lapping area, the image size or toggling the stage
    profiling on one handle must give the same results as a fresh handle"""
    ge = ORBextractor(500, 1.2, 5)
    img_a, img_b = synth.frame(240, 320, 7), synth.frame(200, 400, 8)
    seq = [(img_a, (0, 0)), (img_a, (100, 220)), (img_a, (100, 220)), (img_b, (0, 0)), (img_a, (0, 0))]
    for step, (img, lap) in enumerate(seq):
        if step == 2:
            ge.set_profiling(True)
        if step == 3:
            ge.set_profiling(False)
        m1, k1, d1 = ge(img, None, lap)
        rc, k0, d0, m0 = port.PortExtractor(500, 1.2, 5).extract(img, lap)
        assert_same_features(k0, d0, m0, k1, d1, m1, f"step {step}")


def test_two_extractors_on_two_threads():
    """the reference runs the left and the right extractor on two std::threads (Frame.cc:122-125): two handles used
    concurrently from two host threads (graph capture, plan building and result copies included) give the serial results"""
    import threading
    imgs = [synth.frame(240, 320, 30 + i) for i in range(4)]
    want = [port.PortExtractor(400, 1.2, 5).extract(im, (0, 0)) for im in imgs]
    errors = []

    def worker(offset):
        try:
            ge = ORBextractor(400, 1.2, 5)
            for rep in range(6):
                i = (offset + rep) % len(imgs)
                m1, k1, d1 = ge(imgs[i], None, (0, 0))
                rc, k0, d0, m0 = want[i]
                assert_same_features(k0, d0, m0, k1, d1, m1, f"thread {offset} rep {rep}")
        except Exception as e:                                 # noqa: BLE001 -- reported below, in the main thread
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(t,)) for t in (0, 2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_single_and_batch_calls_mixed_on_one_handle():
    """the single-frame CUDA graph captures the staging pointer; a batch call that regrows the staging area must invalidate it:
    single -> batch of 8 -> single on one handle, each checked against a fresh handle"""
    NF = 8
    frames = synth.sequence(240, 320, NF, canvas=512, base_seed=300)
    img = synth.frame(240, 320, 77)
    ge = ORBextractor(300, 1.2, 4, max_batch=NF)
    want1 = ORBextractor(300, 1.2, 4)(img, None, (0, 0))
    wantB = ORBextractor(300, 1.2, 4, max_batch=NF).extract_batch_host(frames, (0, 0))
    for step in range(2):
        m1, k1, d1 = ge(img, None, (0, 0))
        assert m1 == want1[0] and np.array_equal(k1, want1[1]) and np.array_equal(d1, want1[2]), f"single, round {step}"
        c, k, d = ge.extract_batch_host(frames, (0, 0))
        assert np.array_equal(c, wantB[0])
        for f in range(NF):
            n = c[f, 0]
            assert np.array_equal(k[f, :n], wantB[1][f, :n]) and np.array_equal(d[f, :n], wantB[2][f, :n]), f"batch frame {f}, round {step}"
    m1, k1, d1 = ge(img, None, (0, 0))
    assert m1 == want1[0] and np.array_equal(k1, want1[1]) and np.array_equal(d1, want1[2])


@pytest.mark.parametrize("ini,mn", [(140, 130), (200, 128), (130, 20), (20, 1), (3, 0)])
def test_extreme_fast_thresholds(ini, mn):
    """thresholds >= 128 take the exact byte compare of the quick reject; minThFAST 0/1 exercise the dark-polarity bookkeeping of
    the packed score network (a corner whose best arc minimum is 1 has response 0 and can never survive NMS)"""
    rng = np.random.default_rng(ini * 7 + mn)
    img = rng.integers(0, 256, (160, 224), dtype=np.uint8)
    img[40:90, 60:140] = 0
    img[100:130, 30:200] = 255
    check_against_port(img, 300, 3, ini=ini, mn=mn)


_SWITCH_SCRIPT = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
from oracle import port
from orb_slam3_ros_b200 import synth
from orb_slam3_ros_b200.extractor import ORBextractor
import torch
def same(a, b):
    return len(a[1]) == len(b[1]) and all(np.array_equal(a[1][f], b[1][f]) for f in ("x", "y", "size", "response", "octave")) and \
        np.abs(a[1]["angle"] - b[1]["angle"]).max() <= 1e-3 and np.unpackbits(a[2] ^ b[2]).sum() <= 1e-4 * a[2].size * 8
img = synth.frame(300, 417, 3)
ge = ORBextractor(600, 1.2, 6)
m1, k1, d1 = ge(img, None, (0, 1000))
rc, k0, d0, m0 = port.PortExtractor(600, 1.2, 6).extract(img, (0, 1000))
assert m0 == m1 and same((m0, k0, d0), (m1, k1, d1)), "single frame"
pe = port.PortExtractor(600, 1.2, 6); pe.extract(img, (0, 1000))
for l in range(6):
    assert np.array_equal(pe.level(l, bordered=True), ge.debug_level(0, l, bordered=True)), l
    assert np.array_equal(pe.raw_keys(l), ge.debug_raw_keys(0, l)), l
NF = 40
frames = synth.sequence(240, 320, NF, canvas=512, base_seed=11)
gb = ORBextractor(300, 1.2, 4, max_batch=NF)
c, k, d = gb.extract_batch_host(frames, (0, 0))
c0, k0, d0 = port.extract_batch(frames, 300, 1.2, 4, lapping=(0, 0), nthreads=4)
assert np.array_equal(c, c0)
for f in range(NF):
    n = c[f, 0]
    assert same((0, k0[f, :n], d0[f, :n]), (0, k[f, :n], d[f, :n])), f
gb.extract_batch_device(torch.from_numpy(frames).cuda(), NF, 320, 240)
c2, k2, d2 = gb.fetch(NF)
assert np.array_equal(c2, c) and all(np.array_equal(k2[f, :c[f, 0]], k[f, :c[f, 0]]) for f in range(NF))
from orb_slam3_ros_b200.matcher import ORBmatcher
db, q = synth.descriptor_db(6000, 700, seed=3)
i1, dd1 = ORBmatcher().knn2(q, db)
i0, dd0 = port.knn2(q, db)
assert np.array_equal(i1, i0) and np.array_equal(dd1, dd0)
print("SWITCH-OK")
"""


@pytest.mark.parametrize("env", [
    {"ORBB_FAST_NO_TMAP": "1"}, {"ORBB_RESIZE_NO_TMA": "1"}, {"ORBB_RESIZE_NO_PRMT": "1"}, {"ORBB_NO_GRAPH": "1"},
    {"ORBB_NO_PDL": "1"}, {"ORBB_BLUR_EARLY": "1"}, {"ORBB_DESC_NO_STAGE": "1"}, {"ORBB_GRAPH_NO_PDL": "1"},
    {"ORBB_FAST_LATENCY_FRAMES": "0"}, {"ORBB_BRANCH_FRAMES": "0"}, {"ORBB_OCTREE_NO_SMEM": "1"}, {"ORBB_ASM_LATENCY_FRAMES": "0"}, {"ORBB_ASM_LATENCY_FRAMES": "1000"}, {"ORBB_FAST_LATENCY_FRAMES": "64", "ORBB_FAST_NO_TMAP": "1"},
    {"ORBB_LANES": "1"}, {"ORBB_LANES": "3", "ORBB_LANES_MIN": "8"}, {"ORBB_LANES_MIN": "8", "ORBB_LANES_HOST": "1"}, {"ORBB_CHUNKS": "4"}, {"ORBB_PYR_LATENCY_FRAMES": "0", "ORBB_OCTREE_LATENCY_FRAMES": "0"},
    {"ORBB_KNN_MIX": "2,2"}, {"ORBB_KNN_MIX": "4,0,3"}, {"ORBB_KNN_MIX": "0,4"},
], ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_every_shipped_switch_keeps_parity(env):
    """the A/B environment switches of DESIGN.md section 5 select alternative kernels / launch shapes; they are read once per
    process, so each setting runs the same parity script (single frame incl. pyramid and raw FAST keys, a 40-frame host batch, the
    device batch, a 2-NN scan -- all against the oracle) in its own interpreter"""
    import os
    import subprocess
    import sys
    root = str(Path(__file__).resolve().parent.parent)
    r = subprocess.run([sys.executable, "-c", _SWITCH_SCRIPT.format(root=root)], env={**os.environ, **env}, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and "SWITCH-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]

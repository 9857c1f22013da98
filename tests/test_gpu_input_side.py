"""GPU parity of the input-side "next" rows: stereo rectification (cv::remap, System.cc:233-240) fused in front of the
extraction, and Frame::UndistortKeyPoints (cv::undistortPoints, Frame.cc:747-780).  The oracle restatements are pinned to
the real cv2 functions by tests/test_oracle_primitives.py."""
import numpy as np
import pytest
import torch

from oracle import port
from orb_slam3_ros_b200 import synth
from orb_slam3_ros_b200.extractor import ORBextractor
from orb_slam3_ros_b200.matcher import ORBmatcher
from orb_slam3_ros_b200.rectify import Rectifier
from test_oracle_primitives import _rectify_maps

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("h,w,seed", [(120, 160, 0), (376, 1241, 1), (97, 131, 2), (480, 752, 3)])
def test_remap_matches_oracle(h, w, seed):
    src = synth.frame(h + 16, w + 24, seed)
    mx, my = _rectify_maps(h, w, seed)
    r = Rectifier(mx, my, src.shape)
    assert np.array_equal(r.remap(src), port.remap_linear(src, mx, my))
    r.close()


def test_rectified_extraction_equals_remap_then_extract():
    h, w = 376, 1241
    mx, my = _rectify_maps(h, w, 5)
    mx[:3] += w                                    # keep the frame mostly inside the source (undo the far-outside rows)
    raw = [synth.frame(h + 16, w + 24, 10 + i) for i in range(3)]
    r = Rectifier(mx, my, raw[0].shape)
    ge = ORBextractor(2000)
    want = [port.PortExtractor(2000, 1.2, 8, 20, 7).extract(port.remap_linear(f, mx, my), (0, 0)) for f in raw]
    # one host frame
    m1, k1, d1 = r.extract(ge, raw[0])
    rc, k0, d0, m0 = want[0]
    assert rc == 0 and m1 == m0 and len(k1) == len(k0)
    for f in ("x", "y", "size", "response", "octave"):
        assert np.array_equal(k0[f], k1[f]), f
    assert np.abs(k0["angle"] - k1["angle"]).max() <= 1e-3
    assert np.unpackbits(d0 ^ d1).sum() <= 1e-4 * d0.size * 8
    # batch of device-resident raw frames
    dev = torch.from_numpy(np.stack(raw)).cuda()
    r.extract_batch_device(ge, dev, len(raw))
    counts, kps, desc = ge.fetch(len(raw))
    for i, (rc, k0, d0, m0) in enumerate(want):
        n = counts[i, 0]
        assert n == len(k0) and counts[i, 1] == m0
        for f in ("x", "y", "size", "response", "octave"):
            assert np.array_equal(k0[f], kps[i, :n][f]), (i, f)
        assert np.unpackbits(d0 ^ desc[i, :n]).sum() <= 1e-4 * d0.size * 8
    r.close()


def test_undistort_points_matches_oracle_bit_for_bit():
    m = ORBmatcher()
    rng = np.random.default_rng(3)
    xy = np.stack([rng.uniform(0, 752, 5000), rng.uniform(0, 480, 5000)], 1).astype(np.float32)
    K4 = (458.654, 457.296, 367.215, 248.375)
    for dist in ([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.011],
                 [0.35, -0.6, 0.01, -0.02, 0.4]):
        got = m.undistort_points(xy, K4, dist)
        want = port.undistort_points(xy, K4, dist)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), dist
    assert np.array_equal(m.undistort_points(xy, K4, [0.0, 0.1, 0.0, 0.0]), xy)       # Frame.cc:749
    assert m.undistort_points(np.zeros((0, 2), np.float32), K4, [0.1, 0, 0, 0]).shape == (0, 2)

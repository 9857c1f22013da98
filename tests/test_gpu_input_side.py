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


def test_rgbd_stereo_matches_oracle():
    """Frame::ComputeStereoFromRGBD (Frame.cc:984-1005) for a TUM-shape batch: float32 depth maps and 16-bit ones with the
    depth-map factor, with and without lens distortion"""
    from orb_slam3_ros_b200.extractor import rgbd_stereo_batch, stereo_fetch
    B, H, W = 3, 480, 640
    frames = synth.sequence(H, W, B)
    ge = ORBextractor(1000, max_batch=B)
    ge.extract_batch_device(torch.from_numpy(frames).cuda(), B, W, H)
    counts, kps, _ = ge.fetch(B)
    rng = np.random.default_rng(9)
    raw = rng.integers(0, 40000, (B, H, W)).astype(np.uint16)
    raw[rng.random((B, H, W)) < 0.2] = 0                                   # holes in the depth map
    factor = np.float32(1.0) / np.float32(5000.0)                          # TUM: DepthMapFactor 5000 (Tracking.cc:617)
    depth_f = (raw.astype(np.float32) * factor).astype(np.float32)
    K4 = (517.306408, 516.469215, 318.643040, 255.313989)                  # TUM1.yaml
    bf = np.float32(40.0)
    for dist in ([0.262383, -0.953104, -0.005358, 0.002628, 1.163314], [0.0, 0.0, 0.0, 0.0]):
        for as_u16 in (False, True):
            dev = torch.from_numpy(raw.view(np.int16) if as_u16 else depth_f).cuda()
            rgbd_stereo_batch(ge, dev, B, W, H, K4, dist, float(bf), depth_is_u16=as_u16, depth_factor=float(factor))
            ur, dp = stereo_fetch(ge, B)
            for f in range(B):
                n = counts[f, 0]
                xy = np.stack([kps[f, :n]["x"], kps[f, :n]["y"]], 1)
                ur0, dp0 = port.rgbd_stereo(xy, depth_f[f], K4, dist, bf)
                assert np.array_equal(dp[f, :n], dp0) and np.array_equal(ur[f, :n].view(np.uint32), ur0.view(np.uint32)), (dist[0], as_u16, f)
            assert 0.1 < (dp[0, :counts[0, 0]] < 0).mean() < 0.3


@pytest.mark.parametrize("src_shape,new_size", [((480, 752), (600, 350)), ((376, 1241), (621, 188)), ((240, 320), (500, 400)), ((300, 401), (333, 257))])
def test_resized_extraction_equals_resize_then_extract(src_shape, new_size):
    """cv::resize(im, imToFeed, newImSize) of System::Track* (System.cc:241-244) on the device: the pyramid's level 0 must
    be the oracle's resize (pinned to cv2.resize, incl. up-scaling and the exact-2x case), the features those of
    resize-then-extract; single host frame and a batch of device-resident frames"""
    h, w = src_shape
    raw = [synth.frame(h, w, 60 + i) for i in range(2)]
    ge = ORBextractor(800, 1.2, 6)
    want_img = [port.resize_linear(f, new_size[0], new_size[1]) for f in raw]
    want = [port.PortExtractor(800, 1.2, 6).extract(im, (0, 0)) for im in want_img]
    m1, k1, d1 = ge.extract_resized(raw[0], new_size)
    assert np.array_equal(ge.debug_level(0, 0), want_img[0])
    rc, k0, d0, m0 = want[0]
    assert rc == 0 and m1 == m0 and len(k1) == len(k0)
    for f in ("x", "y", "size", "response", "octave"):
        assert np.array_equal(k0[f], k1[f]), f
    assert np.unpackbits(d0 ^ d1).sum() <= 1e-4 * d0.size * 8
    ge.extract_batch_resized_device(torch.from_numpy(np.stack(raw)).cuda(), 2, w, h, new_size)
    counts, kps, desc = ge.fetch(2)
    for i, (rc, k0, d0, m0) in enumerate(want):
        n = counts[i, 0]
        assert n == len(k0)
        for f in ("x", "y", "size", "response", "octave"):
            assert np.array_equal(k0[f], kps[i, :n][f]), (i, f)

"""GPU parity of the fused grid lookup + best/second scan ("next" row: Frame::GetFeaturesInArea, Frame.cc:657-723, inside
ORBmatcher::SearchByProjection, ORBmatcher.cc:71-120) against the oracle that builds the real 64x48 grid of
Frame::AssignFeaturesToGrid and walks it cell by cell."""
import numpy as np
import pytest

from oracle import port
from orb_slam3_ros_b200 import synth
from orb_slam3_ros_b200.extractor import ORBextractor
from orb_slam3_ros_b200.matcher import ORBmatcher

pytestmark = pytest.mark.gpu


def _frame_features(h, w, nf, idx):
    ge = ORBextractor(nf)
    _, k, d = ge(synth.frame(h, w, idx))
    return k, d


@pytest.mark.parametrize("with_stereo", [False, True])
def test_search_area_matches_grid_oracle(with_stereo):
    h, w = 480, 752
    k, d = _frame_features(h, w, 1000, 3)
    n = len(k)
    rng = np.random.default_rng(5)
    kps_xy = np.stack([k["x"], k["y"]], 1)
    grid4 = np.float32([0.0, 0.0, np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)])      # Frame.cc:251-252
    nq = 1500
    src = rng.integers(0, n, nq)
    # projected map points: near a keypoint (match expected), far outside the image, and exactly on cell borders
    qx = k["x"][src] + rng.normal(0, 3, nq).astype(np.float32)
    qy = k["y"][src] + rng.normal(0, 3, nq).astype(np.float32)
    qx[::50] = rng.choice([-40.0, w + 60.0, 0.0, w / 64 * 7], len(qx[::50])).astype(np.float32)
    qy[::70] = rng.choice([-30.0, h + 50.0, 0.0, h / 48 * 5], len(qy[::70])).astype(np.float32)
    r = (rng.choice([2.5, 4.0], nq) * rng.choice([1.0, 1.2, 1.44, 3.0, 15.0], nq)).astype(np.float32)
    lvl = k["octave"][src]
    qlev = np.stack([lvl - 1, lvl], 1).astype(np.int32)
    qlev[::9] = (-1, -1)                                   # no level check
    qlev[5::9, 1] = -1                                     # only a lower bound
    qdesc = d[src].copy()
    flips = rng.integers(0, 256, (nq, 20))
    bits = np.unpackbits(qdesc, axis=1)
    np.bitwise_xor.at(bits, (np.repeat(np.arange(nq), 20), flips.ravel()), 1)
    qdesc = np.packbits(bits, axis=1)
    skip = (rng.random(n) < 0.2).astype(np.uint8)
    u_right = np.where(rng.random(n) < 0.5, k["x"] - rng.uniform(1, 40, n), -1).astype(np.float32) if with_stereo else None
    queries = np.stack([qx, qy, r, qx - rng.uniform(0, 45, nq).astype(np.float32)], 1).astype(np.float32)
    m = ORBmatcher()
    for init in (256, 100):
        want = port.search_area_best2(kps_xy, k["octave"], d, grid4, queries, qlev, qdesc, skip, u_right, init)
        got = m.search_area_best2(kps_xy, k["octave"], d, grid4, queries, qlev, qdesc, skip, u_right, init)
        assert np.array_equal(want, got), init
    assert (want[:, 1] >= 0).mean() > 0.5                  # most queries really found something
    # duplicated keypoints / descriptors: the first minimum in (cell column, cell row, index) order must win
    kd = np.concatenate([kps_xy, kps_xy[:200] + np.float32(0.25)])
    od = np.concatenate([k["octave"], k["octave"][:200]])
    dd = np.concatenate([d, d[:200]])
    want = port.search_area_best2(kd, od, dd, grid4, queries, qlev, qdesc, None, None, 256)
    got = m.search_area_best2(kd, od, dd, grid4, queries, qlev, qdesc, None, None, 256)
    assert np.array_equal(want, got)


def test_search_area_empty_inputs():
    m = ORBmatcher()
    grid4 = np.float32([0, 0, 0.1, 0.1])
    q = np.float32([[10, 10, 5, 0]])
    out = m.search_area_best2(np.zeros((0, 2), np.float32), np.zeros(0, np.int32), np.zeros((0, 32), np.uint8), grid4, q,
                              np.int32([[-1, -1]]), np.zeros((1, 32), np.uint8))
    assert out.tolist() == [[256, -1, 256, -1]]

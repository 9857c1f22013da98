"""CPU: the array-rebuild formulation of DistributeOctTree (the algorithm the CUDA kernel runs) against the
std::list + std::sort oracle (oracle/orb_port.cpp, following ORBextractor.cc:555-779)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import port

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def model():
    out = ROOT / "tests" / "models" / "_build"
    out.mkdir(exist_ok=True)
    so = out / "liboctree_model.so"
    src = ROOT / "tests" / "models" / "octree_array_model.cpp"
    hdr = ROOT / "orb_slam3_ros_b200" / "csrc" / "introsort.cuh"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", str(src), "-o", str(so)])
    lib = C.CDLL(str(so))
    lib.model_distribute.restype = C.c_int
    lib.model_distribute.argtypes = [C.c_void_p, C.c_int] + [C.c_int] * 5 + [C.c_void_p, C.c_int]
    return lib


def _model(lib, xyr, w, h, n_feat):
    xyr = np.ascontiguousarray(xyr, np.float32)
    out = np.zeros(n_feat + 64, np.int32)
    rc = lib.model_distribute(xyr.ctypes.data, len(xyr), 16, w - 16, 16, h - 16, n_feat, out.ctypes.data, len(out))
    assert rc >= 0, rc
    return out[:rc]


def _random_keys(rng, w, h, n, clustered):
    ww, hh = w - 32 - 6, h - 32 - 6
    if clustered:
        cx = rng.integers(0, ww, 12)
        cy = rng.integers(0, hh, 12)
        k = rng.integers(0, 12, n)
        x = np.clip(cx[k] + rng.normal(0, 9, n).astype(int), 0, ww - 1)
        y = np.clip(cy[k] + rng.normal(0, 9, n).astype(int), 0, hh - 1)
    else:
        x = rng.integers(0, ww, n)
        y = rng.integers(0, hh, n)
    xy = np.unique(np.stack([x, y], 1), axis=0)          # FAST yields at most one key per pixel
    rng.shuffle(xy)
    resp = rng.integers(7, 40, len(xy))                    # few distinct responses -> first-max ties
    return np.concatenate([xy + 3, resp[:, None]], 1).astype(np.float32)


@pytest.mark.parametrize("shape", [(752, 480), (1241, 376), (640, 480), (210, 134), (3840, 2160), (300, 290)])
def test_model_matches_list_oracle(model, shape):
    w, h = shape
    rng = np.random.default_rng(w * 7 + h)
    for trial in range(40):
        n_feat = int(rng.choice([5, 60, 217, 434, 1086, 1502]))
        n = int(rng.integers(0, 8 * n_feat + 10))
        xyr = _random_keys(rng, w, h, n, clustered=bool(trial % 2))
        want = port.distribute(xyr, 16, w - 16, 16, h - 16, n_feat)
        got = _model(model, xyr, w, h, n_feat)
        assert np.array_equal(want, got), (trial, n_feat, len(xyr), len(want), len(got))


def test_model_on_real_fast_keys(model):
    from orb_slam3_ros_b200 import synth
    for idx, (h, w, nf) in enumerate([(480, 752, 1000), (376, 1241, 2000), (480, 640, 1000)]):
        img = synth.frame(h, w, idx)
        e = port.PortExtractor(nf, 1.2, 8, 20, 7)
        rc, _, _, _ = e.extract(img)
        assert rc == 0
        for l in range(8):
            lw, lh = e.level(l).shape[1], e.level(l).shape[0]
            raw = e.raw_keys(l)
            want = port.distribute(raw, 16, lw - 16, 16, lh - 16, int(e.features_per_level[l]))
            got = _model(model, raw, lw, lh, int(e.features_per_level[l]))
            assert np.array_equal(want, got), (idx, l)

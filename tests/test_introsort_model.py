"""CPU: the product's libstdc++-std::sort emulation (csrc/introsort.cuh, compiled for the host) must reproduce the
real std::sort's arrangement of equivalent elements (compareNodes ties, ORBextractor.cc:538-553,700)."""
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
    so = out / "libintrosort_model.so"
    src = ROOT / "tests" / "models" / "introsort_host.cpp"
    hdr = ROOT / "orb_slam3_ros_b200" / "csrc" / "introsort.cuh"
    if not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", str(src), "-o", str(so)])
    lib = C.CDLL(str(so))
    lib.model_sort_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.model_killer_sequence.argtypes = [C.c_int, C.c_void_p]
    lib.model_sort_nodes_pf.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    return lib


def _model_sort(lib, sizes, x0s):
    """serial emulation; the parallel formulation (what the CUDA kernel executes) must give the same arrangement"""
    sizes = np.ascontiguousarray(sizes, np.int32)
    x0s = np.ascontiguousarray(x0s, np.int32)
    perm = np.zeros(len(sizes), np.int32)
    lib.model_sort_nodes(sizes.ctypes.data, x0s.ctypes.data, len(sizes), perm.ctypes.data)
    perm_pf = np.zeros(len(sizes), np.int32)
    lib.model_sort_nodes_pf(sizes.ctypes.data, x0s.ctypes.data, len(sizes), perm_pf.ctypes.data)
    assert np.array_equal(perm, perm_pf), "parallel formulation differs from the serial emulation"
    return perm


def test_tie_heavy_random(model):
    rng = np.random.default_rng(0)
    for trial in range(400):
        n = int(rng.integers(0, 900))
        sizes = rng.integers(2, 2 + int(rng.integers(1, 6)), n)      # very few distinct sizes -> many ties
        x0s = rng.integers(0, int(rng.integers(1, 40)), n) * 11
        assert np.array_equal(_model_sort(model, sizes, x0s), port.sort_nodes(sizes, x0s)), trial


def test_structured_inputs(model):
    for n in (1, 2, 15, 16, 17, 18, 31, 32, 33, 64, 257, 1000, 2048, 5000):
        for sizes in (np.arange(n), np.arange(n)[::-1], np.zeros(n, int), np.arange(n) % 3, (np.arange(n) * 7919) % 13):
            x0s = (np.arange(n) * 31) % 17
            assert np.array_equal(_model_sort(model, sizes + 2, x0s), port.sort_nodes(sizes + 2, x0s)), n


def test_killer_adversary_hits_heapsort_path(model):
    before = model.model_heap_fallbacks()
    for n in (200, 1000, 3000):
        keys = np.zeros(n, np.int32)
        model.model_killer_sequence(n, keys.ctypes.data)
        # distinct keys (forces the depth-limit fallback), then the same with ties folded in
        for fold in (1, 7):
            sizes = keys // fold + 2
            x0s = np.zeros(n, np.int32)
            assert np.array_equal(_model_sort(model, sizes, x0s), port.sort_nodes(sizes, x0s)), (n, fold)
    assert model.model_heap_fallbacks() > before, "adversary did not reach the depth-limit fallback"

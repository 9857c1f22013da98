"""CPU: the C-ABI library loads and exports every symbol include/orbb200.h declares; without a CUDA device the
compute entry points fail loudly (no CPU fallback, no oracle behind the product)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from orb_slam3_ros_b200 import build, capi
    build.build_library()
    return capi.load()


def test_header_symbols_are_exported(lib):
    from orb_slam3_ros_b200 import capi
    hdr = (ROOT / "include" / "orbb200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(orbb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in orbb200.h but not exported"
    assert declared == {s[0] for s in capi.SYMBOLS}, "capi.SYMBOLS and the header disagree"


def test_struct_layouts():
    from orb_slam3_ros_b200 import capi
    assert capi.KP_DTYPE.itemsize == 24 and C.sizeof(capi.Params) == 28


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from orb_slam3_ros_b200 import capi
    from orb_slam3_ros_b200.extractor import ORBextractor
    from orb_slam3_ros_b200.matcher import ORBmatcher
    with pytest.raises(capi.OrbbError) as e:
        ORBextractor()
    assert e.value.code == capi.ORBB_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(capi.OrbbError):
        ORBmatcher()
    # the "next"-row handles fail the same way: vocabulary (ComputeBoW) and rectifier (cv::remap)
    from orb_slam3_ros_b200.bow import Vocabulary, synthetic_vocabulary
    from orb_slam3_ros_b200.rectify import Rectifier
    import numpy as np
    with pytest.raises(capi.OrbbError) as e:
        Vocabulary(synthetic_vocabulary(3, 2))
    assert e.value.code == capi.ORBB_ERR_CUDA
    with pytest.raises(capi.OrbbError) as e:
        Rectifier(np.zeros((8, 8), np.float32), np.zeros((8, 8), np.float32), (8, 8))
    assert e.value.code == capi.ORBB_ERR_CUDA
    # the single-pair distance is host code by design (ORBmatcher::DescriptorDistance)
    assert ORBmatcher.DescriptorDistance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_product_never_imports_oracle():
    for py in (ROOT / "orb_slam3_ros_b200").rglob("*.py"):
        src = py.read_text()
        assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), f"{py} references the oracle"
    for cu in (ROOT / "orb_slam3_ros_b200" / "csrc").iterdir():
        assert "orb_port" not in cu.read_text()


def test_argument_errors_are_reported_before_any_device_work(lib):
    """bad arguments come back as ORBB_ERR_ARG (with text) whether or not a GPU is present"""
    import ctypes as C
    import numpy as np
    from orb_slam3_ros_b200 import capi
    out = C.c_void_p()
    m = np.zeros((4, 4), np.float32)
    assert lib.orbb_rectifier_create(0, capi.ptr(m), capi.ptr(m), 4, 0, 4, 4, 4, C.byref(out)) == capi.ORBB_ERR_ARG      # dst width 0
    assert lib.orbb_rectifier_create(0, capi.ptr(m), capi.ptr(m), 2, 4, 4, 4, 4, C.byref(out)) == capi.ORBB_ERR_ARG      # stride < width
    assert lib.orbb_rectifier_create(0, None, capi.ptr(m), 4, 4, 4, 4, 4, C.byref(out)) == capi.ORBB_ERR_ARG
    assert b"rectifier" in lib.orbb_last_error(None)
    cb = np.zeros(1, np.int32)
    assert lib.orbb_vocab_create(0, 0, capi.ptr(cb), capi.ptr(cb), capi.ptr(cb), 0, capi.ptr(cb), capi.ptr(cb), capi.ptr(cb), 1, C.byref(out)) == capi.ORBB_ERR_ARG
    assert lib.orbb_create(None, C.byref(out)) == capi.ORBB_ERR_ARG
    assert lib.orbb_extract(None, None, 0, 0, 0, 0, 0, None, None, 0, None, None) == capi.ORBB_ERR_ARG
    assert lib.orbb_undistort_points(None, None, 0, None, None, 0, None, None) == capi.ORBB_ERR_ARG

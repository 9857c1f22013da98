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
    # the single-pair distance is host code by design (ORBmatcher::DescriptorDistance)
    import numpy as np
    assert ORBmatcher.DescriptorDistance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_product_never_imports_oracle():
    for py in (ROOT / "orb_slam3_ros_b200").rglob("*.py"):
        src = py.read_text()
        assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), f"{py} references the oracle"
    for cu in (ROOT / "orb_slam3_ros_b200" / "csrc").iterdir():
        assert "orb_port" not in cu.read_text()

"""CPU: the rBRIEF sampling table shipped with the product / oracle equals the reference's bit_pattern_31_ (when the
reference tree is present, i.e. in the build container)."""
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference/orb_slam3/src/ORBextractor.cc")


def _oracle_table():
    return np.loadtxt(ROOT / "oracle" / "bit_pattern_31.txt", dtype=np.int32)


def test_product_table_equals_oracle_table():
    inc = (ROOT / "orb_slam3_ros_b200" / "csrc" / "orb_pattern.inc").read_text()
    vals = [int(v) for v in re.findall(r"-?\d+", re.sub(r"//.*", "", inc))]
    assert np.array_equal(np.array(vals, np.int32).reshape(256, 4), _oracle_table())
    assert np.abs(_oracle_table()).max() <= 15


@pytest.mark.skipif(not REF.exists(), reason="reference tree not present on this machine")
def test_tables_equal_reference():
    import sys
    sys.path.insert(0, str(ROOT / "tools"))
    import gen_pattern
    assert np.array_equal(np.array(gen_pattern.parse_reference(), np.int32).reshape(256, 4), _oracle_table())

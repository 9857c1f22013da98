#!/usr/bin/env python3
"""Phase times of the level-0 quadtree CTA of one frame (debug build: nvcc -DORBB_OT_TIMING, library path in ORBB_LIB)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from orb_slam3_ros_b200 import synth                       # noqa: E402
from orb_slam3_ros_b200.extractor import ORBextractor      # noqa: E402

ext = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=1)
img = synth.frame(480, 752, 3)
for _ in range(5):
    ext(img, None, (0, 1000))
buf = (C.c_longlong * 64)()
names = {0: "start", 1: "prefix", 2: "gather", 3: "roots", 10: "elist", 11: "split", 12: "emit", 20: "p2 split", 21: "p2 sort", 22: "p2 emit", 4: "best"}
acc = {}
for rep in range(5):
    ext(img, None, (0, 1000))
    n = ext._lib.orbb_debug_ot_timing(buf)
    t = [(buf[2 * i], buf[2 * i + 1]) for i in range(n)]
    row = [(names[int(tag)], (c - t[i - 1][0])) for i, (c, tag) in enumerate(t) if i > 0]
    if rep == 4:
        print(" | ".join("%s %d" % (k, v) for k, v in row), "| total cycles", t[-1][0] - t[0][0], "| pending", buf[61], "nodes", buf[62], "keys", buf[63])

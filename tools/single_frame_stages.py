#!/usr/bin/env python3
"""Per-stage CUDA-event times of ONE frame (the SLAM thread's call shape)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from orb_slam3_ros_b200 import synth
from orb_slam3_ros_b200.extractor import ORBextractor
h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (480, 752)
nf, nl = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1000, 8)
img = synth.frame(h, w, 3)
ext = ORBextractor(nf, 1.2, nl, 20, 7, max_batch=1)
for _ in range(5):
    ext(img, None, (0, 1000))
ext.set_profiling(True)
acc = {}
for _ in range(20):
    ext(img, None, (0, 1000))
    for k, v in ext.stage_times().items():
        acc[k] = acc.get(k, 0.0) + v / 20
print({k: round(v * 1e3, 1) for k, v in acc.items()}, "us; sum", round(sum(acc.values()) * 1e3, 1))

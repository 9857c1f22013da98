#!/usr/bin/env python3
"""Small driver for ncu captures: a few batched extractions (and optionally one kNN call) with no timing logic."""
import argparse
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from orb_slam3_ros_b200 import synth                              # noqa: E402
from orb_slam3_ros_b200.extractor import ORBextractor             # noqa: E402
from orb_slam3_ros_b200.matcher import ORBmatcher                 # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--knn", action="store_true")
a = ap.parse_args()
frames = torch.from_numpy(synth.sequence(480, 752, a.batch)).cuda()
ext = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=a.batch)
for _ in range(a.iters):
    ext.extract_batch_device(frames, a.batch, 752, 480, lapping=(0, 1000))
ext.sync()
print("keypoints/frame:", ext.fetch(a.batch, with_data=False)[0][:, 0].mean())
if a.knn:
    db, q = synth.descriptor_db(400_000, 20_000, seed=77)
    m = ORBmatcher()
    d_db, d_q = torch.from_numpy(db).cuda(), torch.from_numpy(q).cuda()
    idx = torch.empty((len(q), 2), dtype=torch.int32, device="cuda")
    dst = torch.empty_like(idx)
    m.knn2_device(d_q, len(q), d_db, len(db), idx, dst)
    torch.cuda.synchronize()
    print("knn ok", int(dst[:, 0].sum()))

#!/usr/bin/env python3
"""Experiment: aggregate resident throughput of H independent extractor handles (own streams, own workspaces) fed back to back from one
host thread, frames split evenly -- the upper bound of what pipelining the lanes across batch calls could give.
usage: multi_handle.py H [frames_per_handle] [iters]"""
import json
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from orb_slam3_ros_b200 import synth                              # noqa: E402
from orb_slam3_ros_b200.extractor import ORBextractor             # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 2
per = int(sys.argv[2]) if len(sys.argv) > 2 else 256 // H
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
frames = [torch.from_numpy(synth.sequence(480, 752, per, base_seed=1234 + 1000 * i)).cuda() for i in range(H)]
exts = [ORBextractor(1000, 1.2, 8, 20, 7, max_batch=per) for _ in range(H)]
for _ in range(3):
    for e, f in zip(exts, frames):
        e.extract_batch_device(f, per, 752, 480, lapping=(0, 1000))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(iters):
    for e, f in zip(exts, frames):
        e.extract_batch_device(f, per, 752, 480, lapping=(0, 1000))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(json.dumps({"handles": H, "frames_per_handle": per, "frames_per_s": round(H * per * iters / dt, 1)}))

#!/usr/bin/env python3
"""Per-stage CUDA-event times of a resident batch (default 256 x 752x480, 1000 features) and the un-profiled step time.
usage: batch_stages.py [--batch N] [--shape H W] [--nf N] [--nl N] [--iters N]   (prints one JSON line)"""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from orb_slam3_ros_b200 import synth                              # noqa: E402
from orb_slam3_ros_b200.extractor import ORBextractor             # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--shape", type=int, nargs=2, default=[480, 752])
ap.add_argument("--nf", type=int, default=1000)
ap.add_argument("--nl", type=int, default=8)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--tag", default="")
a = ap.parse_args()
h, w = a.shape
frames = [torch.from_numpy(synth.sequence(h, w, a.batch, base_seed=1234 + 1000 * i)).cuda() for i in range(2)]
ext = ORBextractor(a.nf, 1.2, a.nl, 20, 7, max_batch=a.batch)
for i in range(3):
    ext.extract_batch_device(frames[i % 2], a.batch, w, h, lapping=(0, 1000))
ext.sync()
st = torch.cuda.ExternalStream(ext.stream)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for i in range(a.iters):
    ext.extract_batch_device(frames[i % 2], a.batch, w, h, lapping=(0, 1000))
e1.record(st)
ext.sync()
step_ms = e0.elapsed_time(e1) / a.iters
ext.set_profiling(True)
acc = {}
for i in range(5):
    ext.extract_batch_device(frames[i % 2], a.batch, w, h, lapping=(0, 1000))
    for k, v in ext.stage_times().items():
        acc[k] = acc.get(k, 0.0) + v / 5
kp = float(ext.fetch(a.batch, with_data=False)[0][:, 0].mean())
print(json.dumps({"tag": a.tag, "batch": a.batch, "shape": [h, w], "step_ms": round(step_ms, 4), "frames_per_s": round(a.batch / step_ms * 1e3, 1),
                  "stage_ms": {k: round(v, 4) for k, v in acc.items() if v > 0}, "kp_per_frame": kp}))

#!/usr/bin/env python3
"""Driver for the ncu capture of the matcher-side kernels: one call of each at the sizes
bench.py measures -- a 16-pair KITTI-shape stereo batch, the SearchLocalPoints scan (1600 map points), a SearchByBoW-shaped
candidate-list scan, ComputeBoW on a 10^5-word tree, 2000 ComputeDistinctiveDescriptors groups, one cv::remap, and (--knn) the
200k x 2M brute-force 2-NN.  No timing logic."""
import argparse
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from orb_slam3_ros_b200 import capi, synth                         # noqa: E402
from orb_slam3_ros_b200.bow import Vocabulary, synthetic_vocabulary  # noqa: E402
from orb_slam3_ros_b200.extractor import ORBextractor, stereo_match_batch  # noqa: E402
from orb_slam3_ros_b200.matcher import ORBmatcher                  # noqa: E402
from orb_slam3_ros_b200.rectify import Rectifier                   # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--knn", action="store_true")
ap.add_argument("--pairs", type=int, default=16)
ap.add_argument("--knn-nq", type=int, default=200_000)
a = ap.parse_args()
lib = capi.load()
_p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
rng = np.random.default_rng(7)
# stereo (S1)
SP = a.pairs
Lh, Rh = synth.stereo_sequence(376, 1241, SP)
eL, eR = ORBextractor(2000, max_batch=SP), ORBextractor(2000, max_batch=SP)
eL.extract_batch_device(torch.from_numpy(Lh).cuda(), SP, 1241, 376)
eR.extract_batch_device(torch.from_numpy(Rh).cuda(), SP, 1241, 376)
stereo_match_batch(eL, eR, SP, 718.856 * 0.53716, 0.53716)
eL.sync()
# one frame for the scans
ext = ORBextractor(1000, 1.2, 8, 20, 7)
_, k, d = ext(synth.frame(480, 752, 8), None, (0, 0))
n = len(k)
m = ORBmatcher()
kp_dev, desc_dev, cnt_dev = C.c_void_p(), C.c_void_p(), C.c_void_p()
capi.check(lib.orbb_batch_device_ptrs(ext._h, C.byref(kp_dev), C.byref(desc_dev), C.byref(cnt_dev)), ext._h)
ev = capi.FrameView()
ev.kps_xy, ev.kps_stride, ev.octaves, ev.oct_stride, ev.desc, ev.u_right, ev.n, ev.on_device = kp_dev.value, 24, kp_dev.value + 20, 24, desc_dev.value, None, n, 1
nmp = 1600
src = rng.integers(0, n, nmp)
lev = np.clip(k["octave"][src] + rng.integers(-1, 2, nmp), 0, 7).astype(np.int32)
queries = np.stack([k["x"][src] + rng.normal(0, 2.5, nmp), k["y"][src] + rng.normal(0, 2.5, nmp), np.float32(4.0) * ext.GetScaleFactors()[lev], k["x"][src] - 20], 1).astype(np.float32)
qlev = np.stack([lev - 1, lev], 1).astype(np.int32)
qdesc = d[src].copy()
skip = (rng.random(n) < 0.25).astype(np.uint8)
out = np.zeros((nmp, 4, 2), np.int32)
grid4 = np.float32([0, 0, 64 / 752, 48 / 480])
capi.check(lib.orbb_search_area_topk(m._m, C.byref(ev), _p(grid4), _p(queries), _p(qlev), _p(qdesc), nmp, _p(skip), 256, 4, _p(out)), m._m, matcher=True)
nq = 1000
q = d[rng.integers(0, n, nq)].copy()
rowptr = (np.arange(nq + 1) * 30).astype(np.int32)
cand = rng.integers(0, n, rowptr[-1]).astype(np.int32)
out4 = np.zeros((nq, 4), np.int32)
capi.check(lib.orbb_best2_csr_dev(m._m, _p(q), nq, desc_dev, n, _p(cand), _p(rowptr), 256, _p(out4)), m._m, matcher=True)
Vocabulary(synthetic_vocabulary(10, 5, seed=3)).transform([d], 4, 1)
ng = 2000
m.distinctive(d[rng.integers(0, n, ng * 8)].copy(), (np.arange(ng + 1) * 8).astype(np.int32))
m.rotation_check([(k["angle"][:500], k["angle"][500:1000])])
yy, xx = np.mgrid[0:480, 0:752].astype(np.float32)
Rectifier((xx + 3.0 * np.sin(yy / 57.0)).astype(np.float32), (yy + 2.0 * np.cos(xx / 91.0)).astype(np.float32), (480, 752)).remap(synth.frame(480, 752, 8))
if a.knn:
    db, qq = synth.descriptor_db(2_000_000, a.knn_nq, seed=77)
    d_db, d_q = torch.from_numpy(db).cuda(), torch.from_numpy(qq).cuda()
    idx = torch.empty((len(qq), 2), dtype=torch.int32, device="cuda")
    dst = torch.empty_like(idx)
    m.knn2_device(d_q, len(qq), d_db, len(db), idx, dst)
    torch.cuda.synchronize()
print("prof_matcher ok")

#!/usr/bin/env python3
"""time the 2-NN kernel for one (CSA, direct) mix given by ORBB_KNN_MIX; checks the result against a small oracle run"""
import os, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from orb_slam3_ros_b200 import synth
from orb_slam3_ros_b200.matcher import ORBmatcher
from oracle import port
nq, nd = 200_000, 2_000_000
db, q = synth.descriptor_db(nd, nq, seed=77)
m = ORBmatcher()
d_db, d_q = torch.from_numpy(db).cuda(), torch.from_numpy(q).cuda()
idx = torch.empty((nq, 2), dtype=torch.int32, device="cuda"); dst = torch.empty_like(idx)
st = torch.cuda.ExternalStream(m.stream)
m.knn2_device(d_q, nq, d_db, nd, idx, dst); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    e0.record()
    for _ in range(2): m.knn2_device(d_q, nq, d_db, nd, idx, dst)
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
i0, d0 = port.knn2(q[:64], db, nthreads=16)
ok = np.array_equal(idx[:64].cpu().numpy(), i0) and np.array_equal(dst[:64].cpu().numpy(), d0)
print(f"mix={os.environ.get('ORBB_KNN_MIX','default')} ms={ms:.1f} Gpairs/s={nq*nd/ms/1e6:.1f} parity={ok}")

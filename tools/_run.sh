set -x
python tools/gpu_sweep.py 11 250 --ref > gpurun_out/r2m_sweep1.log 2>&1; tail -3 gpurun_out/r2m_sweep1.log
python tools/gpu_sweep.py 12 150 --big > gpurun_out/r2m_sweep2.log 2>&1; tail -3 gpurun_out/r2m_sweep2.log
python - > gpurun_out/r2m_determinism.log 2>&1 <<'P'
# lanes + programmatic launches: 30 repeated 256-frame batches must give byte-identical results (a missing dependency would show up as a flicker)
import sys, zlib, numpy as np, torch
sys.path.insert(0, '.')
from orb_slam3_ros_b200 import synth
from orb_slam3_ros_b200.extractor import ORBextractor
fr = torch.from_numpy(synth.sequence(480, 752, 256)).cuda()
ge = ORBextractor(1000, 1.2, 8, max_batch=256)
crcs = set()
for i in range(30):
    ge.extract_batch_device(fr, 256, 752, 480, lapping=(0, 1000))
    if i % 3 == 2:
        c, k, d = ge.fetch(256)
        crcs.add((zlib.crc32(c.tobytes()), zlib.crc32(k.tobytes()), zlib.crc32(d.tobytes())))
print("distinct results over 10 fetches of 30 back-to-back batches:", len(crcs))
assert len(crcs) == 1
P
tail -2 gpurun_out/r2m_determinism.log

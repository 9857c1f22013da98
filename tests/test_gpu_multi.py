"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the database-sharded brute-force 2-NN through the C ABI
entry orbb_knn2_sharded -- per-shard scans, ONE packed ncclAllGather over NVLink, device merge -- must equal the unsharded
orbb_knn2_dev result on every rank, bit for bit (indices, distances, lowest-index ties), and the oracle port on a sample.
Run on a multi-GPU box:  gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu -q"""
import os
import sys
import tempfile
import zlib
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, nq, nd, tmp):
    sys.path.insert(0, str(ROOT))
    import time
    import torch
    from orb_slam3_ros_b200 import synth
    from orb_slam3_ros_b200.matcher import ORBmatcher
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    m = ORBmatcher(device=rank)
    idf = Path(tmp) / "nccl_id"
    if rank == 0:
        uid = ORBmatcher.nccl_unique_id()
        (Path(tmp) / "nccl_id.tmp").write_bytes(uid)
        os.replace(Path(tmp) / "nccl_id.tmp", idf)
    else:
        for _ in range(600):
            if idf.exists():
                break
            time.sleep(0.05)
        uid = idf.read_bytes()
    comm = m.nccl_comm_create(world, rank, uid)
    db, q = synth.descriptor_db(nd, nq, seed=5)
    db[nd // 2 + 3] = db[7]                                   # duplicates across shards: the lowest index must win
    db[nd - 1] = db[7]
    q[0] = db[7]
    lo, hi = rank * nd // world, (rank + 1) * nd // world
    d_q = torch.from_numpy(q).to(dev)
    d_shard = torch.from_numpy(db[lo:hi]).to(dev)
    d_all = torch.from_numpy(db).to(dev)
    idx = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    dst = torch.empty_like(idx)
    for _ in range(2):                                         # twice: scratch reuse
        m.knn2_sharded_device(comm, d_q, nq, d_shard, hi - lo, lo, idx, dst)
    ref_i = torch.empty_like(idx)
    ref_d = torch.empty_like(idx)
    m.knn2_device(d_q, nq, d_all, nd, ref_i, ref_d)
    torch.cuda.synchronize()
    ok = bool(torch.equal(idx, ref_i) and torch.equal(dst, ref_d))
    crc = zlib.crc32(idx.cpu().numpy().tobytes() + dst.cpu().numpy().tobytes())
    (Path(tmp) / f"result_{rank}").write_text(f"{int(ok)} {crc} {int(idx[0, 0])} {int(idx[0, 1])}")
    m.nccl_comm_destroy(comm)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_knn2_equals_unsharded(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    import torch.multiprocessing as mp
    nq, nd = 3000, 200_003                                     # (not divisible by the shard count)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(world, nq, nd, tmp), nprocs=world, join=True)
        res = [(Path(tmp) / f"result_{r}").read_text().split() for r in range(world)]
    assert all(r[0] == "1" for r in res), res                  # sharded == unsharded on every rank
    assert len({r[1] for r in res}) == 1, res                  # the same bytes on every rank
    assert res[0][2] == "7" and int(res[0][3]) == nd // 2 + 3  # duplicated rows: lowest global indices, in order

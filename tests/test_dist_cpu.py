"""CPU, gloo, world_size 2: the host logic of the N>1 paths -- frame partition (no collective) and the sharded 2-NN
exchange (all-gather of per-shard top-2 with global indices, then the (distance, index) merge).  The per-shard scan is
done by the oracle port here (no GPU in this container); on the GPU box tests/test_gpu_match.py checks the CUDA scan
and merge kernels against the same answers."""
import os
import sys
from pathlib import Path

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp
import torch

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port_no, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import port
    from orb_slam3_ros_b200 import sharding, synth
    nq, nd, nframes = 300, 5001, 13
    db, qq = synth.descriptor_db(nd, nq, seed=9, dup_every=41)
    lo, hi = sharding.block_bounds(nd, world, rank)
    idx, dst = port.knn2(qq, db[lo:hi])
    idx = np.where(idx >= 0, idx + lo, -1).astype(np.int32)          # index_base = shard start
    g_idx = [torch.empty((nq, 2), dtype=torch.int32) for _ in range(world)]
    g_dst = [torch.empty((nq, 2), dtype=torch.int32) for _ in range(world)]
    dist.all_gather(g_idx, torch.from_numpy(idx))
    dist.all_gather(g_dst, torch.from_numpy(dst))
    mi, md = sharding.merge_top2(torch.stack(g_idx).numpy(), torch.stack(g_dst).numpy())
    i0, d0 = port.knn2(qq, db)
    ok_knn = bool(np.array_equal(mi, i0) and np.array_equal(md, d0))
    # frame partition: contiguous, disjoint, complete -- gathered without any data-path collective
    flo, fhi = sharding.block_bounds(nframes, world, rank)
    owned = torch.zeros(nframes, dtype=torch.int32)
    owned[flo:fhi] = 1
    dist.all_reduce(owned)
    ok_frames = bool((owned == 1).all())
    q.put((rank, ok_knn, ok_frames, (flo, fhi)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_knn_and_frame_partition_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
    assert res[0][3] == (0, 6) and res[1][3] == (6, 13)


def test_merge_top2_tie_rule():
    from orb_slam3_ros_b200 import sharding
    idx = np.array([[[5, 7]], [[12, -1]], [[20, 21]]], np.int32)      # 3 shards, 1 query
    dst = np.array([[[4, 9]], [[4, sharding.INT_MAX]], [[3, 4]]], np.int32)
    mi, md = sharding.merge_top2(idx, dst)
    assert mi.tolist() == [[20, 5]] and md.tolist() == [[3, 4]]       # dist 4 tie -> lowest global index
    mi, md = sharding.merge_top2(np.full((2, 1, 2), -1, np.int32), np.full((2, 1, 2), sharding.INT_MAX, np.int32))
    assert mi.tolist() == [[-1, -1]] and md.tolist() == [[sharding.INT_MAX] * 2]

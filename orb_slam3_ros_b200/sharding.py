"""Host-side partitioning for the multi-GPU paths.

Extraction / stereo: frames are independent -> contiguous blocks of frames per rank, no collective.
Brute-force 2-NN against a large database: contiguous database shards, every rank scans all queries against its
shard and reports top-2 with GLOBAL indices (index_base = shard start); one all-gather of [nq,2] (idx, dist) per
rank, then a merge by lexicographic (distance, index).  Contiguous shards + that order reproduce cv::BFMatcher's
"lowest train index wins ties" rule exactly.
"""
import numpy as np

INT_MAX = np.iinfo(np.int32).max


def block_bounds(n, world, rank):
    """contiguous block [lo, hi) of n items owned by `rank` (frames or database rows)"""
    return rank * n // world, (rank + 1) * n // world


def merge_top2(idx_shards, dist_shards):
    """numpy statement of the shard merge (the device version is k_knn2_merge_shards); inputs [shards, nq, 2]."""
    idx = np.asarray(idx_shards).transpose(1, 0, 2).reshape(idx_shards.shape[1], -1).astype(np.int64)
    dist = np.asarray(dist_shards).transpose(1, 0, 2).reshape(idx.shape[0], -1).astype(np.int64)
    key = np.where(idx >= 0, dist * (1 << 32) + idx, np.iinfo(np.int64).max)
    order = np.argsort(key, axis=1, kind="stable")[:, :2]
    rows = np.arange(idx.shape[0])[:, None]
    out_i = idx[rows, order].astype(np.int32)
    out_d = dist[rows, order].astype(np.int32)
    missing = np.take_along_axis(key, order, 1) == np.iinfo(np.int64).max
    out_i[missing] = -1
    out_d[missing] = INT_MAX
    return out_i, out_d

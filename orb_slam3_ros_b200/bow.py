"""Python mirror of the bag-of-words transform (DBoW2 TemplatedVocabulary::transform behind Frame::ComputeBoW,
reference orb_slam3/src/Frame.cc:738-745) on liborbb200.so, plus a seeded synthetic vocabulary tree (the reference's
ORBvoc.txt.bin blob is not available offline)."""
import ctypes as C

import numpy as np

from . import capi


def synthetic_vocabulary(k=10, depth=3, seed=7, stop_fraction=0.05, ragged=False):
    """Random k-ary tree of `depth` levels in DBoW2's flat form (see include/orbb200.h).  Leaves are words with random
    positive idf weights; a few words are "stopped" (weight 0).  ragged=True prunes some children (k varies per node)."""
    rng = np.random.default_rng(seed)
    child_begin, child_count, child_list = [0], [0], []
    level_nodes = [0]
    nnodes = 1
    for _ in range(depth):
        nxt = []
        for p in level_nodes:
            kk = int(rng.integers(max(2, k - 4), k + 1)) if ragged else k
            child_begin[p], child_count[p] = len(child_list), kk
            for _c in range(kk):
                child_list.append(nnodes)
                child_begin.append(0)
                child_count.append(0)
                nxt.append(nnodes)
                nnodes += 1
        level_nodes = nxt
    node_desc = rng.integers(0, 256, (nnodes, 32), dtype=np.uint8)
    node_weight = np.zeros(nnodes, np.float64)
    node_word = np.full(nnodes, -1, np.int32)
    leaves = np.array(level_nodes)
    node_word[leaves] = np.arange(len(leaves), dtype=np.int32)
    w = rng.uniform(0.5, 12.0, len(leaves))
    w[rng.random(len(leaves)) < stop_fraction] = 0.0
    node_weight[leaves] = w
    return dict(child_begin=np.array(child_begin, np.int32), child_count=np.array(child_count, np.int32),
                child_list=np.array(child_list, np.int32), node_desc=node_desc, node_weight=node_weight, node_word=node_word,
                depth=depth)


class Vocabulary:
    def __init__(self, vocab, device=0):
        self._lib = capi.load()
        self._v = C.c_void_p()
        a = {k: np.ascontiguousarray(x) for k, x in vocab.items() if k != "depth"}
        capi.check(self._lib.orbb_vocab_create(device, len(a["child_begin"]), capi.ptr(a["child_begin"]), capi.ptr(a["child_count"]),
                                               capi.ptr(a["child_list"]), len(a["child_list"]), capi.ptr(a["node_desc"]),
                                               capi.ptr(a["node_weight"]), capi.ptr(a["node_word"]), int(vocab["depth"]), C.byref(self._v)))

    def close(self):
        if getattr(self, "_v", None):
            self._lib.orbb_vocab_destroy(self._v)
            self._v = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform(self, desc_sets, levelsup=4, norm=1):
        """desc_sets: list of [n_i,32] uint8 arrays -> list of (bow_id, bow_val, fv_node, fv_start, fv_feat, n_valid)"""
        sizes = [len(d) for d in desc_sets]
        rowptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        total = int(rowptr[-1])
        desc = np.ascontiguousarray(np.concatenate([np.asarray(d, np.uint8).reshape(-1, 32) for d in desc_sets]) if total else np.zeros((0, 32), np.uint8))
        T = max(total, 1)
        bow_id, bow_val = np.zeros(T, np.int32), np.zeros(T, np.float64)
        fv_node, fv_start, fv_feat = np.zeros(T, np.int32), np.zeros(T, np.int32), np.zeros(T, np.int32)
        counts = np.zeros((len(sizes), 3), np.int32)
        rc = self._lib.orbb_bow_transform(self._v, capi.ptr(desc), capi.ptr(rowptr), len(sizes), levelsup, norm, capi.ptr(bow_id),
                                          capi.ptr(bow_val), capi.ptr(fv_node), capi.ptr(fv_start), capi.ptr(fv_feat), capi.ptr(counts))
        if rc != capi.ORBB_OK:
            raise capi.OrbbError(rc, (self._lib.orbb_vocab_last_error(self._v) or b"").decode())
        out = []
        for s, (lo, c) in enumerate(zip(rowptr[:-1], counts)):
            out.append((bow_id[lo:lo + c[0]].copy(), bow_val[lo:lo + c[0]].copy(), fv_node[lo:lo + c[1]].copy(),
                        fv_start[lo:lo + c[1]].copy(), fv_feat[lo:lo + c[2]].copy(), int(c[2])))
        return out

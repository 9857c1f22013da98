"""Python mirror of the Hamming-matching part of ORB_SLAM3::ORBmatcher (reference orb_slam3/include/ORBmatcher.h:38-94,
orb_slam3/src/ORBmatcher.cc) and of Frame's BFMatcher use (Frame.cc:1144-1151) on liborbb200.so.

The projection geometry / MapPoint bookkeeping of the twelve Search*/Fuse methods stays in the reference's host C++
(out of scope); what moves to the GPU is the scan itself: best/second-best over candidate lists
(best2_csr) and the brute-force 2-NN (knn2) with the ratio test.
"""
import ctypes as C

import numpy as np

from . import capi
from .constants import HISTO_LENGTH, TH_HIGH, TH_LOW

INT_MAX = np.iinfo(np.int32).max


class ORBmatcher:
    TH_LOW = TH_LOW
    TH_HIGH = TH_HIGH
    HISTO_LENGTH = HISTO_LENGTH

    def __init__(self, nnratio=0.6, check_ori=True, device=0):
        self._lib = capi.load()
        self.mfNNratio = np.float32(nnratio)
        self.mbCheckOrientation = check_ori
        self.device = device
        self._m = C.c_void_p()
        capi.check(self._lib.orbb_matcher_create(device, C.byref(self._m)))

    def close(self):
        if getattr(self, "_m", None):
            self._lib.orbb_matcher_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def DescriptorDistance(a, b):
        """ORBmatcher::DescriptorDistance (ORBmatcher.cc:2058-2074): host popcount, never a GPU round trip."""
        a = np.ascontiguousarray(a, np.uint8)
        b = np.ascontiguousarray(b, np.uint8)
        return capi.load().orbb_hamming_distance(capi.ptr(a), capi.ptr(b))

    # ---- Frame::UndistortKeyPoints (Frame.cc:747-780) ----
    def undistort_points(self, xy, K4, dist, new_K4=None):
        """cv::undistortPoints(xy, out, K, dist, noArray(), newK): xy float32 [n,2]; K4 = (fx, fy, cx, cy); dist = k1 k2 p1 p2 [k3..]"""
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        k = np.ascontiguousarray(K4, np.float32)
        nk = k if new_K4 is None else np.ascontiguousarray(new_K4, np.float32)
        d = np.ascontiguousarray(dist, np.float32).ravel()
        out = np.zeros_like(xy)
        capi.check(self._lib.orbb_undistort_points(self._m, capi.ptr(xy), len(xy), capi.ptr(k), capi.ptr(d), len(d), capi.ptr(nk), capi.ptr(out)),
                   self._m, matcher=True)
        return out

    # ---- brute-force 2-NN (Frame.cc:1144) ----
    def knn2(self, query, train):
        """host arrays [nq,32], [nd,32] uint8 -> idx[nq,2], dist[nq,2] (missing = -1 / INT_MAX)"""
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        idx = np.full((len(q), 2), -1, np.int32)
        dist = np.full((len(q), 2), INT_MAX, np.int32)
        capi.check(self._lib.orbb_knn2(self._m, capi.ptr(q), len(q), capi.ptr(t), len(t), capi.ptr(idx), capi.ptr(dist)),
                   self._m, matcher=True)
        return idx, dist

    def knn2_device(self, q_dev, nq, db_dev, nd, idx_dev, dist_dev, index_base=0):
        """device pointers (ints / torch CUDA tensors); asynchronous on the matcher's stream"""
        capi.check(self._lib.orbb_knn2_dev(self._m, capi.ptr(q_dev), nq, capi.ptr(db_dev), nd, index_base, capi.ptr(idx_dev),
                                           capi.ptr(dist_dev)), self._m, matcher=True)

    # ---- the same against a database sharded over the ranks of an NCCL communicator (BASELINE config 4) ----
    def nccl_comm_create(self, nranks, rank, unique_id):
        """ncclCommInitRank through the library (it binds the process's libnccl at run time); unique_id = 128 bytes from
        nccl_unique_id() on one rank, handed to the others by any means (torch.distributed broadcast, a file, MPI ...)"""
        comm = C.c_void_p()
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        capi.check(self._lib.orbb_nccl_comm_create(self.device, nranks, rank, buf, C.byref(comm)))
        return comm

    @staticmethod
    def nccl_unique_id():
        buf = (C.c_char * 128)()
        capi.check(capi.load().orbb_nccl_unique_id(buf))
        return bytes(buf)

    def nccl_comm_destroy(self, comm):
        self._lib.orbb_nccl_comm_destroy(comm)

    def knn2_sharded_device(self, comm, q_dev, nq, db_shard_dev, nd_shard, index_base, idx_dev, dist_dev):
        """every rank: all queries x its database shard -> ONE packed all-gather of the per-shard top-2 -> device merge;
        idx_dev / dist_dev [nq,2] hold the global result on every rank (asynchronous on the matcher's stream)"""
        capi.check(self._lib.orbb_knn2_sharded(self._m, comm, capi.ptr(q_dev), nq, capi.ptr(db_shard_dev), nd_shard, index_base,
                                               capi.ptr(idx_dev), capi.ptr(dist_dev)), self._m, matcher=True)

    def merge_shards_device(self, idx_sh_dev, dist_sh_dev, nshards, nq, idx_dev, dist_dev):
        capi.check(self._lib.orbb_knn2_merge_dev(self._m, capi.ptr(idx_sh_dev), capi.ptr(dist_sh_dev), nshards, nq,
                                                 capi.ptr(idx_dev), capi.ptr(dist_dev)), self._m, matcher=True)

    def ratio_test_device(self, idx_dev, dist_dev, nq, ratio, keep_dev):
        capi.check(self._lib.orbb_ratio_test_dev(self._m, capi.ptr(idx_dev), capi.ptr(dist_dev), nq, float(ratio),
                                                 capi.ptr(keep_dev)), self._m, matcher=True)

    @property
    def stream(self):
        return self._lib.orbb_matcher_stream(self._m)

    @property
    def launch_count(self):
        return int(self._lib.orbb_matcher_launch_count(self._m))

    # ---- best / second-best candidate scans (ORBmatcher.cc:77-120 and siblings) ----
    def best2_csr(self, query, train, cand, rowptr, init=256):
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        cand = np.ascontiguousarray(cand, np.int32)
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        assert len(rowptr) == len(q) + 1
        out = np.zeros((len(q), 4), np.int32)
        capi.check(self._lib.orbb_best2_csr(self._m, capi.ptr(q), len(q), capi.ptr(t), len(t), capi.ptr(cand), capi.ptr(rowptr),
                                            int(init), capi.ptr(out)), self._m, matcher=True)
        return out

    # ---- Frame::GetFeaturesInArea + SearchByProjection scan (Frame.cc:657-723, ORBmatcher.cc:71-120), batched ----
    def search_area_best2(self, kps_xy, octaves, train, grid4, queries, qlev, qdesc, skip=None, u_right=None, init=256):
        kps_xy = np.ascontiguousarray(kps_xy, np.float32).reshape(-1, 2)
        octaves = np.ascontiguousarray(octaves, np.int32)
        train = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        grid4 = np.ascontiguousarray(grid4, np.float32)
        queries = np.ascontiguousarray(queries, np.float32).reshape(-1, 4)
        qlev = np.ascontiguousarray(qlev, np.int32).reshape(-1, 2)
        qdesc = np.ascontiguousarray(qdesc, np.uint8).reshape(-1, 32)
        sk = None if skip is None else np.ascontiguousarray(skip, np.uint8)
        ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
        out = np.zeros((len(queries), 4), np.int32)
        capi.check(self._lib.orbb_search_area_best2(self._m, capi.ptr(kps_xy), capi.ptr(octaves), capi.ptr(train), len(kps_xy),
                                                    capi.ptr(grid4), capi.ptr(queries), capi.ptr(qlev), capi.ptr(qdesc), len(queries),
                                                    capi.ptr(sk), capi.ptr(ur), int(init), capi.ptr(out)), self._m, matcher=True)
        return out

    # ---- MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:329-403), batched ----
    def distinctive(self, desc, rowptr):
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        rowptr = np.ascontiguousarray(rowptr, np.int32)
        best = np.zeros(len(rowptr) - 1, np.int32)
        capi.check(self._lib.orbb_distinctive_csr(self._m, capi.ptr(desc), len(desc), capi.ptr(rowptr), len(best), capi.ptr(best)),
                   self._m, matcher=True)
        return best

    # ---- rotation-consistency filter (ORBmatcher.cc:345-352, :405-423) ----
    def rotation_check(self, angle_sets):
        """angle_sets: list of (angle_a[n], angle_b[n]) float32 pairs, one per match set -> list of (keep[n] bool, (ind1, ind2, ind3))"""
        sizes = [len(a) for a, _ in angle_sets]
        rowptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        total = int(rowptr[-1])
        A = np.ascontiguousarray(np.concatenate([np.asarray(a, np.float32) for a, _ in angle_sets]) if total else np.zeros(0, np.float32))
        Bn = np.ascontiguousarray(np.concatenate([np.asarray(b, np.float32) for _, b in angle_sets]) if total else np.zeros(0, np.float32))
        keep = np.zeros(max(total, 1), np.uint8)
        ind3 = np.zeros((max(len(sizes), 1), 3), np.int32)
        capi.check(self._lib.orbb_rotation_check_csr(self._m, capi.ptr(A), capi.ptr(Bn), total, capi.ptr(rowptr), len(sizes), capi.ptr(keep),
                                                     capi.ptr(ind3)), self._m, matcher=True)
        return [(keep[lo:hi].astype(bool), tuple(int(v) for v in ind3[s])) for s, (lo, hi) in enumerate(zip(rowptr[:-1], rowptr[1:]))]

    # ---- ComputeThreeMaxima (ORBmatcher.cc:2012-2053), host ----
    @staticmethod
    def ComputeThreeMaxima(histo_sizes):
        max1 = max2 = max3 = 0
        ind1 = ind2 = ind3 = -1
        for i, s in enumerate(histo_sizes):
            if s > max1:
                max3, max2, max1 = max2, max1, s
                ind3, ind2, ind1 = ind2, ind1, i
            elif s > max2:
                max3, max2 = max2, s
                ind3, ind2 = ind2, i
            elif s > max3:
                max3, ind3 = s, i
        if max2 < np.float32(0.1) * np.float32(max1):
            ind2 = ind3 = -1
        elif max3 < np.float32(0.1) * np.float32(max1):
            ind3 = -1
        return ind1, ind2, ind3

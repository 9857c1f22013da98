"""GPU parity tests of the Hamming matcher and the stereo matcher, through the C ABI, against the oracle port
(brute-force 2-NN pinned to cv2.BFMatcher in tests/test_oracle_primitives.py) and the committed cv2 golden vectors.
Everything here is integer / index work: bit-exact."""
from pathlib import Path

import numpy as np
import pytest

from oracle import port
from orb_slam3_ros_b200 import capi, synth
from orb_slam3_ros_b200.extractor import ORBextractor, compute_stereo_matches, stereo_fetch, stereo_match_batch
from orb_slam3_ros_b200.matcher import INT_MAX, ORBmatcher

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def matcher():
    m = ORBmatcher()
    yield m
    m.close()


def test_descriptor_distance_known_answers():
    z, f = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    assert ORBmatcher.DescriptorDistance(z, f) == 256 and ORBmatcher.DescriptorDistance(f, f) == 0
    rng = np.random.default_rng(0)
    for _ in range(50):
        a, b = rng.integers(0, 256, (2, 32), dtype=np.uint8)
        assert ORBmatcher.DescriptorDistance(a, b) == port.hamming(a, b)


def test_knn2_golden_cv2_bfmatcher(matcher):
    g = np.load(GOLD / "knn_300x4000.npz")
    idx, dist = matcher.knn2(g["q"], g["db"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])


def test_knn2_tie_rule_lowest_index(matcher):
    q = np.zeros((3, 32), np.uint8)
    tr = np.zeros((6, 32), np.uint8)
    for i, d in enumerate([3, 1, 1, 2, 1, 1]):
        tr[i, 0] = (1 << d) - 1
    idx, dist = matcher.knn2(q, tr)
    assert idx.tolist() == [[1, 2]] * 3 and dist.tolist() == [[1, 1]] * 3


@pytest.mark.parametrize("nq,nd", [(1, 1), (5, 2), (1000, 1000), (1025, 70001), (3000, 300000), (17, 1 << 20)])
def test_knn2_matches_oracle(matcher, nq, nd):
    db, q = synth.descriptor_db(nd, nq, seed=nq + nd, dup_every=97)
    i1, d1 = matcher.knn2(q, db)
    i0, d0 = port.knn2(q, db, nthreads=8)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


def test_knn2_degenerate_sizes(matcher):
    db, q = synth.descriptor_db(10, 4, seed=1)
    i1, d1 = matcher.knn2(q, db[:0])
    assert (i1 == -1).all() and (d1 == INT_MAX).all()
    i1, d1 = matcher.knn2(q, db[:1])
    assert (i1[:, 0] == 0).all() and (i1[:, 1] == -1).all() and (d1[:, 1] == INT_MAX).all()
    i1, d1 = matcher.knn2(q[:0], db)
    assert i1.shape == (0, 2)


def test_knn2_sharded_merge_equals_unsharded_and_ratio(matcher):
    """config-4 semantics on one GPU: contiguous database shards, per-shard top-2 with global indices, merge by
    (distance, index); the ratio test of Frame.cc:1151 afterwards."""
    import torch
    nq, nd, shards = 2000, 120000, 8
    db, q = synth.descriptor_db(nd, nq, seed=4, dup_every=31)
    d_q = torch.from_numpy(q).cuda()
    g_idx = torch.empty((shards, nq, 2), dtype=torch.int32, device="cuda")
    g_dst = torch.empty((shards, nq, 2), dtype=torch.int32, device="cuda")
    parts = []
    for s in range(shards):
        lo, hi = s * nd // shards, (s + 1) * nd // shards
        parts.append(torch.from_numpy(db[lo:hi]).cuda())
        matcher.knn2_device(d_q, nq, parts[-1], hi - lo, g_idx[s], g_dst[s], index_base=lo)
    f_idx = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
    f_dst = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
    keep = torch.empty((nq,), dtype=torch.uint8, device="cuda")
    matcher.merge_shards_device(g_idx, g_dst, shards, nq, f_idx, f_dst)
    matcher.ratio_test_device(f_idx, f_dst, nq, 0.7, keep)
    torch.cuda.synchronize()
    i0, d0 = port.knn2(q, db, nthreads=8)
    assert np.array_equal(f_idx.cpu().numpy(), i0) and np.array_equal(f_dst.cpu().numpy(), d0)
    want = (i0[:, 1] >= 0) & (d0[:, 0].astype(np.float32).astype(np.float64) < d0[:, 1].astype(np.float32).astype(np.float64) * 0.7)
    got = keep.cpu().numpy().astype(bool)
    assert np.array_equal(want, got) and 0 < got.sum() < nq


def test_best2_csr_matches_oracle(matcher):
    rng = np.random.default_rng(6)
    db, q = synth.descriptor_db(5000, 800, seed=2, dup_every=13)
    lens = rng.integers(0, 120, len(q))
    lens[::7] = 0                                            # empty candidate lists
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cand = rng.integers(0, len(db), rowptr[-1]).astype(np.int32)
    for init in (256, 100, 50, INT_MAX):                     # 256 (:77), TH_HIGH, TH_LOW, INT_MAX starts
        assert np.array_equal(matcher.best2_csr(q, db, cand, rowptr, init), port.best2_csr(q, db, cand, rowptr, init)), init


def _stereo_case(h, w, nf, idx):
    left, right = synth.stereo_pair(h, w, idx)
    pl, pr = port.PortExtractor(nf), port.PortExtractor(nf)
    _, kl, dl, _ = pl.extract(left)
    _, kr, dr, _ = pr.extract(right)
    return left, right, pl, pr, kl, dl, kr, dr


def test_stereo_matches_oracle_kitti_shape():
    bf, b = np.float32(718.856 * 0.53716), np.float32(0.53716)      # config/Stereo/KITTI00-02.yaml
    left, right, pl, pr, kl, dl, kr, dr = _stereo_case(376, 1241, 2000, 0)
    ur0, dp0, br0, sad0, kept = port.stereo(pl, pr, kl, dl, kr, dr, bf, b)
    gl, gr = ORBextractor(2000), ORBextractor(2000)
    gl(left)
    gr(right)
    ur1, dp1, br1, sad1 = compute_stereo_matches(gl, gr, float(bf), float(b))
    assert (ur0 >= 0).sum() > 0.3 * len(ur0)                        # the synthetic pair really exercises the matcher
    assert np.array_equal(br0, br1) and np.array_equal(sad0, sad1)
    assert np.array_equal(ur0, ur1) and np.array_equal(dp0, dp1)


def test_stereo_batch_and_edge_cases():
    import torch
    bf, b = np.float32(380.0), np.float32(0.5)
    pairs = [synth.stereo_pair(240, 400, i, dmax=40) for i in range(3)]
    # pair 2: a right image without any texture -> no right keypoints -> no matches (empty vDistIdx guard)
    pairs[2] = (pairs[2][0], np.full_like(pairs[2][1], 100))
    L = np.stack([p[0] for p in pairs])
    R = np.stack([p[1] for p in pairs])
    gl, gr = ORBextractor(600, 1.2, 6, max_batch=3), ORBextractor(600, 1.2, 6, max_batch=3)
    gl.extract_batch_device(torch.from_numpy(L).cuda(), 3, 400, 240)
    gr.extract_batch_device(torch.from_numpy(R).cuda(), 3, 400, 240)
    stereo_match_batch(gl, gr, 3, float(bf), float(b))
    ur, dp = stereo_fetch(gl, 3)
    counts, _, _ = gl.fetch(3, with_data=False)
    for f in range(3):
        pl, pr = port.PortExtractor(600, 1.2, 6), port.PortExtractor(600, 1.2, 6)
        _, kl, dl, _ = pl.extract(L[f])
        _, kr, dr, _ = pr.extract(R[f])
        ur0, dp0, _, _, _ = port.stereo(pl, pr, kl, dl, kr, dr, bf, b)
        n = counts[f, 0]
        assert n == len(ur0)
        assert np.array_equal(ur[f, :n], ur0) and np.array_equal(dp[f, :n], dp0), f
    assert (ur[2, :counts[2, 0]] == -1).all()


def test_distinctive_descriptors_match_oracle(matcher):
    """MapPoint::ComputeDistinctiveDescriptors ("next" row), batched over map points"""
    rng = np.random.default_rng(21)
    sizes = list(rng.integers(0, 40, 300)) + [1, 2, 3, 64, 97, 200, 0]
    rowptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    db, _ = synth.descriptor_db(len(sizes), 2, seed=8)
    desc = np.repeat(db, sizes, axis=0)
    flip = rng.integers(0, 256, (len(desc), 12))
    bits = np.unpackbits(desc, axis=1)
    np.bitwise_xor.at(bits, (np.repeat(np.arange(len(desc)), 12), flip.ravel()), 1)
    desc = np.packbits(bits, axis=1)
    assert np.array_equal(matcher.distinctive(desc, rowptr), port.distinctive(desc, rowptr))


def test_knn2_full_size_properties(matcher):
    """BASELINE config 4 at full size (200 000 queries x 2 000 000 database rows, one shard): size-independent properties
    of the result plus an exact oracle scan of a sample of queries."""
    import torch
    nq, nd = 200_000, 2_000_000
    db, q = synth.descriptor_db(nd, nq, seed=77)
    d_db, d_q = torch.from_numpy(db).cuda(), torch.from_numpy(q).cuda()
    idx = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
    dst = torch.empty((nq, 2), dtype=torch.int32, device="cuda")
    matcher.knn2_device(d_q, nq, d_db, nd, idx, dst)
    torch.cuda.synchronize()
    idx, dst = idx.cpu().numpy(), dst.cpu().numpy()
    assert (idx >= 0).all() and (idx < nd).all() and (idx[:, 0] != idx[:, 1]).all()
    assert (dst[:, 0] <= dst[:, 1]).all() and (dst >= 0).all() and (dst <= 256).all()
    # the reported distances are the distances of the reported rows
    for col in (0, 1):
        d = np.unpackbits(db[idx[::997, col]] ^ q[::997], axis=1).sum(1)
        assert np.array_equal(d, dst[::997, col])
    # planted queries (first half: a database row with <= 40 flipped bits) find a row at least that close; ties on
    # duplicated rows go to the lower index
    assert (dst[: nq // 2, 0] <= 40).all()
    assert (dst[nq // 2:, 0] > 40).mean() > 0.99            # random queries have no close row (256-bit uniform: ~85 at best)
    eq = dst[:, 0] == dst[:, 1]
    assert (idx[eq, 0] < idx[eq, 1]).all()
    # exact scan of a sample on the CPU oracle
    import os
    pick = np.r_[np.arange(0, nq // 2, 80), np.arange(nq // 2, nq, 100)]          # 2 250 queries: planted and random halves
    i0, d0 = port.knn2(q[pick], db, nthreads=max(1, os.cpu_count() or 1))
    assert np.array_equal(idx[pick], i0) and np.array_equal(dst[pick], d0)


def test_stereo_full_size_batch_properties():
    """BASELINE config 2 shape (1241x376, 2000 features per eye) as a batch: determinism, value ranges, and the
    sequential per-frame call gives the same numbers as the batch."""
    import torch
    B = 16
    bf, b = float(np.float32(718.856 * 0.53716)), float(np.float32(0.53716))
    pairs = [synth.stereo_pair(376, 1241, i) for i in range(B)]
    L = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda()
    R = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
    gl, gr = ORBextractor(2000, max_batch=B), ORBextractor(2000, max_batch=B)
    out = []
    for _ in range(2):
        gl.extract_batch_device(L, B, 1241, 376)
        gr.extract_batch_device(R, B, 1241, 376)
        stereo_match_batch(gl, gr, B, bf, b)
        out.append(stereo_fetch(gl, B))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    ur, dp = out[0]
    counts, kps, _ = gl.fetch(B)
    matched = 0
    for f in range(B):
        n = counts[f, 0]
        u, d, x = ur[f, :n], dp[f, :n], kps[f, :n]["x"]
        ok = u >= 0
        matched += ok.sum()
        assert ((u == -1) == (d == -1)).all()
        assert (d[ok] > 0).all() and (x[ok] - u[ok] >= 0).all() and (x[ok] - u[ok] < np.float32(bf / b)).all()     # disparity in [0, maxD)
        assert np.allclose(d[ok], np.float32(bf) / np.maximum(x[ok] - u[ok], np.float32(0.01)), rtol=1e-6)
    assert matched > 0.3 * counts[:, 0].sum()
    f = 5                                                     # one pair against the oracle
    pl, pr = port.PortExtractor(2000), port.PortExtractor(2000)
    _, kl, dl, _ = pl.extract(pairs[f][0])
    _, kr, dr, _ = pr.extract(pairs[f][1])
    ur0, dp0, _, _, _ = port.stereo(pl, pr, kl, dl, kr, dr, np.float32(bf), np.float32(b))
    n = counts[f, 0]
    assert n == len(ur0) and np.array_equal(ur[f, :n], ur0) and np.array_equal(dp[f, :n], dp0)


def test_rotation_check_matches_oracle(matcher):
    """rotation histogram + ComputeThreeMaxima (ORBmatcher.cc:345-352, :405-423, :2012-2053), several match sets per call"""
    rng = np.random.default_rng(12)
    sets = []
    for n, spread in ((900, 8.0), (300, 60.0), (57, 200.0), (1, 1.0), (0, 1.0), (2000, 0.5)):
        base = rng.uniform(0, 360, n).astype(np.float32)
        delta = rng.normal(rng.uniform(0, 360), spread, n)
        other = np.mod(base - delta, 360).astype(np.float32)
        sets.append((base, other))
    # exact bin boundaries (multiples of 15 and 30 degrees), equal angles, and a tie between the fullest bins
    a = np.float32(np.arange(0, 360, 7.5))
    sets.append((a, np.zeros_like(a)))
    sets.append((np.float32([10, 10, 50, 50, 200, 200, 200, 200]), np.float32([10, 10, 10, 10, 10, 10, 50, 50])))
    got = matcher.rotation_check(sets)
    for (a, b), (keep, ind) in zip(sets, got):
        want_keep, want_ind = port.rotation_check(a, b)
        assert ind == want_ind and np.array_equal(keep, want_keep)


def test_best2_csr_rejects_bad_candidate_lists():
    """host-side validation (the arrays are host-resident): a candidate index outside the train set or a non-monotonic rowptr is
    an argument error, not an out-of-bounds device read"""
    import ctypes as C
    lib = capi.load()
    m = ORBmatcher()
    rng = np.random.default_rng(0)
    q = rng.integers(0, 256, (4, 32), dtype=np.uint8)
    tr = rng.integers(0, 256, (10, 32), dtype=np.uint8)
    out = np.zeros((4, 4), np.int32)
    good_row = np.array([0, 2, 2, 5, 6], np.int32)
    for cand, rowptr in [(np.array([0, 1, 2, 3, 10, 5], np.int32), good_row),        # 10 >= ntrain
                         (np.array([0, 1, -1, 3, 4, 5], np.int32), good_row),
                         (np.array([0, 1, 2, 3, 4, 5], np.int32), np.array([0, 3, 2, 5, 6], np.int32)),
                         (np.array([0, 1, 2, 3, 4, 5], np.int32), np.array([1, 2, 2, 5, 6], np.int32))]:
        rc = lib.orbb_best2_csr(m._m, capi.ptr(q), 4, capi.ptr(tr), 10, capi.ptr(cand), capi.ptr(rowptr), 256, capi.ptr(out))
        assert rc == capi.ORBB_ERR_ARG
    cand = np.array([0, 1, 2, 3, 9, 5], np.int32)
    assert lib.orbb_best2_csr(m._m, capi.ptr(q), 4, capi.ptr(tr), 10, capi.ptr(cand), capi.ptr(good_row), 256, capi.ptr(out)) == 0
    assert np.array_equal(out, port.best2_csr(q, tr, cand, good_row, 256))


def test_fisheye_stereo_tail_matching_equals_bfmatcher_with_ratio(matcher):
    """Frame::ComputeStereoFishEyeMatches (Frame.cc:1126-1166), descriptor part: both eyes extracted with their lapping areas (key
    points inside it are written from the back, so rows monoIndex.. of each descriptor matrix are the overlap), brute-force 2-NN between
    the two tails, Lowe's 0.7 test -- against cv2.BFMatcher(NORM_HAMMING).knnMatch + the same test on the oracle's extraction"""
    import torch
    from oracle import orb_ref
    left, right = synth.stereo_pair(376, 620, 9, dmax=30)
    lap_l, lap_r = (250, 619), (0, 370)                       # Frame.cc:1059-1060: the overlapping columns of the two cameras
    gl, gr = ORBextractor(800, 1.2, 8), ORBextractor(800, 1.2, 8)
    ml, kl, dl = gl(left, None, lap_l)
    mr, kr, dr = gr(right, None, lap_r)
    rc, k0, d0, m0 = port.PortExtractor(800, 1.2, 8).extract(left, lap_l)
    rc, k1, d1, m1 = port.PortExtractor(800, 1.2, 8).extract(right, lap_r)
    assert (ml, mr) == (m0, m1) and np.array_equal(dl, d0) and np.array_equal(dr, d1)
    assert 0 < ml < len(kl) and 0 < mr < len(kr)
    tl, tr = dl[ml:], dr[mr:]
    idx, dist = matcher.knn2(tl, tr)
    d_i, d_d = torch.from_numpy(idx).cuda(), torch.from_numpy(dist).cuda()
    keep = torch.zeros(len(tl), dtype=torch.uint8, device="cuda")
    matcher.ratio_test_device(d_i, d_d, len(tl), 0.7, keep)
    torch.cuda.synchronize()
    i0, dd0 = orb_ref.bf_knn2(tl, tr)                          # cv2.BFMatcher
    want = (i0[:, 1] >= 0) & (dd0[:, 0].astype(np.float32).astype(np.float64) < dd0[:, 1].astype(np.float32).astype(np.float64) * 0.7)
    assert np.array_equal(idx, i0) and np.array_equal(dist, dd0)
    assert np.array_equal(keep.cpu().numpy().astype(bool), want) and 5 < want.sum() < len(tl)

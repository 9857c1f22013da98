#!/usr/bin/env python3
"""Stage-by-stage parity report of the CUDA extractor against the oracle port (run on the GPU box)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import port                                    # noqa: E402  (checker only)
from orb_slam3_ros_b200 import synth                       # noqa: E402
from orb_slam3_ros_b200.extractor import ORBextractor, compute_stereo_matches      # noqa: E402
from orb_slam3_ros_b200.matcher import ORBmatcher          # noqa: E402


def diag(h, w, nf, nl, lap, idx=0):
    img = synth.frame(h, w, idx)
    pe = port.PortExtractor(nf, 1.2, nl, 20, 7)
    rc, k0, d0, m0 = pe.extract(img, lap)
    ge = ORBextractor(nf, 1.2, nl, 20, 7)
    t = time.time()
    m1, k1, d1 = ge(img, None, lap)
    print(f"== {w}x{h} nf={nf} nl={nl} lap={lap}: port n={len(k0)} mono={m0} | gpu n={len(k1)} mono={m1} ({time.time()-t:.3f}s)")
    ok = True
    for l in range(nl):
        a, b = pe.level(l, bordered=True), ge.debug_level(0, l, bordered=True)
        pyr = int((a != b).sum()) if a.shape == b.shape else -1
        ra, rb = pe.raw_keys(l), ge.debug_raw_keys(0, l)
        raw_eq = ra.shape == rb.shape and np.array_equal(ra, rb)
        sa = pe.selected(l)
        sb = ge.debug_selected(0, l)
        sa3 = np.stack([sa["x"], sa["y"], sa["response"]], 1) if len(sa) else np.zeros((0, 3), np.float32)
        sel_eq = sa3.shape == sb.shape and np.array_equal(sa3, sb)
        bl = -2
        if len(sa):
            x, y = pe.level(l, blurred=True), ge.debug_level(0, l, blurred=True)
            bl = int((x != y).sum()) if x.shape == y.shape else -1
        print(f"  L{l}: pyr_mismatch={pyr} raw {len(ra)}/{len(rb)} eq={raw_eq} sel {len(sa)}/{len(sb)} eq={sel_eq} blur_mismatch={bl}")
        ok &= pyr == 0 and raw_eq and sel_eq and bl in (0, -2)
    if len(k0) == len(k1):
        same = all(np.array_equal(k0[f], k1[f]) for f in ("x", "y", "size", "response", "octave"))
        dang = np.abs(k0["angle"] - k1["angle"]).max() if len(k0) else 0
        bits = int(np.unpackbits(d0 ^ d1).sum())
        print(f"  final: kp fields equal={same} max|dangle|={dang:g} desc bit diffs={bits}/{d0.size*8} mono {m0}/{m1}")
        ok &= same and dang <= 1e-3 and bits <= 1e-4 * d0.size * 8 and m0 == m1
    else:
        ok = False
    print("  RESULT:", "OK" if ok else "MISMATCH")
    return ok


def diag_stereo():
    h, w, nf = 376, 1241, 2000
    left, right = synth.stereo_pair(h, w, 0)
    pl, pr = port.PortExtractor(nf), port.PortExtractor(nf)
    _, kl, dl, _ = pl.extract(left)
    _, kr, dr, _ = pr.extract(right)
    bf, b = 718.856 * 0.53716, 0.53716
    ur0, dp0, br0, sad0, kept = port.stereo(pl, pr, kl, dl, kr, dr, np.float32(bf), np.float32(b))
    gl, gr = ORBextractor(nf), ORBextractor(nf)
    gl(left); gr(right)
    ur1, dp1, br1, sad1 = compute_stereo_matches(gl, gr, bf, b)
    ok = len(ur0) == len(ur1) and np.array_equal(ur0, ur1) and np.array_equal(dp0, dp1) and np.array_equal(br0, br1) and np.array_equal(sad0, sad1)
    print(f"== stereo: nL={len(ur0)}/{len(ur1)} matched port={(ur0>=0).sum()} gpu={(ur1>=0).sum()} "
          f"uR eq={np.array_equal(ur0, ur1)} depth eq={np.array_equal(dp0, dp1)} bestR eq={np.array_equal(br0, br1)} sad eq={np.array_equal(sad0, sad1)}")
    print("  RESULT:", "OK" if ok else "MISMATCH")
    return ok


def diag_knn():
    db, q = synth.descriptor_db(50000, 3000, seed=5, dup_every=97)
    m = ORBmatcher()
    t = time.time()
    i1, d1 = m.knn2(q, db)
    dt = time.time() - t
    i0, d0 = port.knn2(q, db, nthreads=8)
    ok = np.array_equal(i0, i1) and np.array_equal(d0, d1)
    print(f"== knn2 3000x50000: idx eq={np.array_equal(i0, i1)} dist eq={np.array_equal(d0, d1)} ({dt:.3f}s)")
    rng = np.random.default_rng(1)
    nq = 500
    lens = rng.integers(0, 90, nq)
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    cand = rng.integers(0, len(db), rowptr[-1]).astype(np.int32)
    o0 = port.best2_csr(q[:nq], db, cand, rowptr, 256)
    o1 = m.best2_csr(q[:nq], db, cand, rowptr, 256)
    # second index is only defined up to the list position; both sides report the candidate index at that position
    ok2 = np.array_equal(o0, o1)
    print(f"== best2_csr: eq={ok2}")
    print("  RESULT:", "OK" if ok and ok2 else "MISMATCH")
    return ok and ok2


if __name__ == "__main__":
    res = [diag(480, 752, 1000, 8, (0, 1000)), diag(376, 1241, 2000, 8, (0, 0)), diag(480, 640, 1000, 8, (0, 0), 3),
           diag(134, 210, 100, 3, (0, 0), 4), diag_stereo(), diag_knn()]
    print("ALL OK" if all(res) else "SOME MISMATCH")
    sys.exit(0 if all(res) else 1)

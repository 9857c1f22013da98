#!/usr/bin/env python3
"""Randomised parity sweep on a GPU box: CUDA extractor (through the C ABI) vs the oracle port -- and the reference's own
ORBextractor.cc when oracle/_ref is present -- over random image sizes, level counts, scale factors, thresholds, feature budgets,
lapping areas and image statistics.  usage: gpu_sweep.py [seed] [cases] [--ref]   (about 0.1-0.3 s of host time per case: the oracle runs on one core)"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import port, ref                                       # noqa: E402  (checker)
from orb_slam3_ros_b200 import capi, synth                         # noqa: E402
from orb_slam3_ros_b200.extractor import ORBextractor              # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cases = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rng = np.random.default_rng(seed)
have_ref = ref.available() and "--ref" in sys.argv      # (the port is swept against the reference source on the CPU: tests/test_reference_source.py)
bad = skipped = 0
start = int(sys.argv[sys.argv.index("--start") + 1]) if "--start" in sys.argv else 0
for it in range(cases):
    big = "--big" in sys.argv
    h, w = (int(rng.integers(70, 700)), int(rng.integers(70, 1100))) if big else (int(rng.integers(70, 500)), int(rng.integers(70, 760)))
    nl = int(rng.integers(1, 11))
    sf = float(rng.choice([1.2, 1.2, 1.2, 1.1, 1.3, 1.5, 2.0, 1.25]))
    nf = int(rng.integers(1, 4000 if big else 1600))
    ini = int(rng.integers(5, 80))
    mn = int(rng.integers(1, ini + 1))
    lap = (int(rng.integers(0, w)), int(rng.integers(0, w)))
    kind = int(rng.integers(0, 4))
    if kind == 0:
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    elif kind == 1:
        img = np.clip(synth.frame(h, w, it).astype(np.int32) // 4 + 100, 0, 255).astype(np.uint8)      # low contrast: minTh fallback cells
    else:
        img = synth.frame(h, w, 1000 + it)
    if it < start:
        continue
    print(it, (h, w, nl, sf, nf, ini, mn, lap, kind), flush=True) if "-v" in sys.argv else None
    rc, k0, d0, m0 = port.PortExtractor(nf, sf, nl, ini, mn).extract(img, lap)
    ge = ORBextractor(nf, sf, nl, ini, mn)
    try:
        m1, k1, d1 = ge(img, None, lap)
    except capi.OrbbError as e:
        if rc != 0 and e.code == capi.ORBB_ERR_UNSUPPORTED:
            skipped += 1                                           # both reject the geometry (undefined behaviour in the reference)
            continue
        print("CUDA ERROR", e, (h, w, nl, sf, nf, ini, mn, lap, kind), "port rc", rc)
        bad += 1
        continue
    finally:
        ge.close()
    if rc != 0:
        print("PORT REJECTS, CUDA ACCEPTS", rc, (h, w, nl, sf, nf, ini, mn, lap, kind))
        bad += 1
        continue
    ok = len(k0) == len(k1) and m0 == m1 and all(np.array_equal(k0[f], k1[f]) for f in ("x", "y", "size", "response", "octave"))
    if ok and len(k0):
        ok = np.abs(k0["angle"] - k1["angle"]).max() <= 1e-3 and np.unpackbits(d0 ^ d1).sum() <= 1e-4 * d0.size * 8
    if ok and have_ref:
        rc2, kr, dr, mr = ref.RefExtractor(nf, sf, nl, ini, mn).extract(img, lap)
        ok = rc2 == 0 and len(kr) == len(k1) and mr == m1 and all(np.array_equal(kr[f], k1[f]) for f in ("x", "y", "size", "response", "octave"))
    if not ok:
        bad += 1
        print("MISMATCH", (h, w, nl, sf, nf, ini, mn, lap, kind), len(k0), len(k1), m0, m1, flush=True)
print(f"sweep seed {seed}: {cases} cases, {skipped} rejected by both, {bad} bad, reference source checked: {have_ref}")
sys.exit(1 if bad else 0)

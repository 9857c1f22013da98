#!/usr/bin/env python3
"""Generate tests/golden/*.npz with the cv2-backed oracle (oracle/orb_ref.py).  Run in the build container; the
fixtures are committed so that the GPU box (no /root/reference, possibly another OpenCV) checks against frozen answers."""
import sys
import zlib
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import orb_ref                      # noqa: E402
from orb_slam3_ros_b200 import synth            # noqa: E402

OUT = ROOT / "tests" / "golden"


def extract_case(name, img, nf, nl, ini, mn, lap):
    e = orb_ref.RefExtractor(nf, 1.2, nl, ini, mn)
    rc, k, d, mono = e.extract(img, lap)
    assert rc == 0
    crc = np.array([zlib.crc32(np.ascontiguousarray(e.pyramid[l]).tobytes()) for l in range(nl)], np.uint32)
    bcrc = np.array([zlib.crc32(np.ascontiguousarray(e.blurred[l]).tobytes()) if e.blurred[l] is not None else 0 for l in range(nl)], np.uint32)
    raw_n = np.array([len(r) for r in e.raw], np.int32)
    sel = np.concatenate([np.stack([s["x"], s["y"], s["response"], s["octave"].astype(np.float32)], 1) for s in e.selected])
    np.savez_compressed(OUT / f"{name}.npz", image=img, params=np.array([nf, nl, ini, mn, lap[0], lap[1]], np.int32), kps=k, desc=d,
                        mono=np.int32(mono), pyr_crc=crc, blur_crc=bcrc, raw_n=raw_n, sel=sel, cv2_version=np.array(cv2.__version__))
    print(name, img.shape, "n =", len(k), "mono =", mono, "raw =", raw_n.tolist())


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    extract_case("mono_320x240", synth.frame(240, 320, 11), 300, 4, 20, 7, (0, 1000))
    extract_case("wide_400x200", synth.frame(200, 400, 12), 500, 5, 20, 7, (0, 0))
    rng = np.random.default_rng(3)
    extract_case("noise_176x144", rng.integers(0, 256, (144, 176), dtype=np.uint8), 200, 3, 20, 7, (60, 110))
    flat = synth.frame(200, 260, 13)
    flat[:, 130:] = 127                     # half the cells have no corner at either threshold
    flat[40:90, 140:200] += (np.arange(60) % 9).astype(np.uint8)   # faint texture: only the minTh fallback fires
    extract_case("fallback_260x200", flat, 250, 4, 20, 7, (0, 0))
    db, q = synth.descriptor_db(4000, 300, seed=21, dup_every=53)
    idx, dist = orb_ref.bf_knn2(q, db)
    np.savez_compressed(OUT / "knn_300x4000.npz", db=db, q=q, idx=idx, dist=dist, cv2_version=np.array(cv2.__version__))
    print("knn", idx.shape)


if __name__ == "__main__":
    main()

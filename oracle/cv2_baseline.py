"""ORACLE / CPU BASELINE (test infrastructure, NOT product code): the reference's extractor with OpenCV's OWN primitives.

The `--impl reference` arm of bench.py times the reference's ORBextractor.cc against the scalar OpenCV stand-ins of oracle/cvshim
(this image has no OpenCV C++ headers to build against); the reference's real build calls OpenCV's SIMD cv::resize / cv::FAST /
cv::GaussianBlur.  This module gives the honest number for that: the control flow of ORBextractor::operator() (orb_slam3/src/
ORBextractor.cc:1086-1168: ComputePyramid :1170-1195, the per-cell FAST loop :787-872, DistributeOctTree, IC_Angle, blur,
descriptors) with python-cv2's resize / copyMakeBorder / FastFeatureDetector / GaussianBlur doing the pixel work and the g++ -O3
glue of oracle/orb_port.cpp for the rest (DistributeOctTree, IC_Angle + fastAtan2, steered BRIEF) -- the same glue the port uses,
so the results are bit-identical to the port's (tests/test_oracle_extract.py::test_cv2_baseline_equals_port).  About 720 cv2 calls
per 752x480 frame carry a few microseconds of Python overhead each; that is part of the number and stated next to it.
"""
import math
import os
import time

import cv2
import numpy as np

from . import port

EDGE = 19


class Cv2Extractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self._p = port.PortExtractor(nfeatures, scale_factor, nlevels, ini_th, min_th)     # tables only (:409-469)
        self.nlevels = nlevels
        self.scale, self.inv_scale = self._p.scale_factors, self._p.inv_scale_factors
        self.per_level, self.umax = self._p.features_per_level, self._p.umax
        self._ini = cv2.FastFeatureDetector_create(ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        self._min = cv2.FastFeatureDetector_create(min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        self._cells = {}
        self.t_prim = 0.0        # seconds spent inside cv2 primitives and the C++ glue (no Python list handling): what a C++ build would pay

    def _grid(self, rows, cols):
        """the cell rectangles of one level (:789-822), cached per level size"""
        key = (rows, cols)
        if key not in self._cells:
            min_b, max_bx, max_by = EDGE - 3, cols - EDGE + 3, rows - EDGE + 3
            width, height = np.float32(max_bx - min_b), np.float32(max_by - min_b)
            n_cols, n_rows = int(width / np.float32(35)), int(height / np.float32(35))
            cells = []
            if n_cols > 0 and n_rows > 0:
                w_cell = int(math.ceil(np.float32(width / np.float32(n_cols))))
                h_cell = int(math.ceil(np.float32(height / np.float32(n_rows))))
                for i in range(n_rows):
                    ini_y = min_b + i * h_cell
                    if ini_y >= max_by - 3:
                        continue
                    max_y = min(ini_y + h_cell + 6, max_by)
                    for j in range(n_cols):
                        ini_x = min_b + j * w_cell
                        if ini_x >= max_bx - 6:
                            continue
                        cells.append((ini_y, max_y, ini_x, min(ini_x + w_cell + 6, max_bx), j * w_cell, i * h_cell))
            self._cells[key] = (cells, min_b, max_bx, max_by)
        return self._cells[key]

    def extract(self, image, lapping=(0, 0)):
        """-> (rc, kps[KP_DTYPE], desc[n,32], monoIndex) like port.PortExtractor.extract"""
        if image is None or image.size == 0:
            return -1, np.zeros(0, port.KP_DTYPE), np.zeros((0, 32), np.uint8), 0
        h, w = image.shape
        pyr, kps_l, desc_l = [], [], []
        clk = time.perf_counter
        t0 = clk()
        for l in range(self.nlevels):                                                     # ComputePyramid :1170-1195
            if l == 0:
                cur = image
            else:
                s = np.float32(self.inv_scale[l])
                size = (int(np.rint(np.float32(w) * s)), int(np.rint(np.float32(h) * s)))
                cur = cv2.resize(pyr[l - 1][EDGE:-EDGE, EDGE:-EDGE], size, interpolation=cv2.INTER_LINEAR)
            pyr.append(cv2.copyMakeBorder(cur, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101))
        tp = clk() - t0
        for l in range(self.nlevels):                                                     # ComputeKeyPointsOctTree :787-896
            img = pyr[l][EDGE:-EDGE, EDGE:-EDGE]
            cells, min_b, max_bx, max_by = self._grid(*img.shape)
            xs, ys, rs = [], [], []
            for (y0, y1, x0, x1, ox, oy) in cells:
                roi = img[y0:y1, x0:x1]
                t0 = clk()
                k = self._ini.detect(roi, None)
                if not k:
                    k = self._min.detect(roi, None)
                tp += clk() - t0
                for kp in k:
                    xs.append(kp.pt[0] + ox); ys.append(kp.pt[1] + oy); rs.append(kp.response)
            raw = np.stack([np.asarray(xs, np.float32), np.asarray(ys, np.float32), np.asarray(rs, np.float32)], 1) if xs else np.zeros((0, 3), np.float32)
            t0 = clk()
            picked = port.distribute(raw, min_b, max_bx, min_b, max_by, int(self.per_level[l])) if len(raw) else np.zeros(0, np.int32)
            tp += clk() - t0
            k = np.zeros(len(picked), port.KP_DTYPE)
            if len(picked):
                k["x"] = raw[picked, 0] + min_b; k["y"] = raw[picked, 1] + min_b; k["response"] = raw[picked, 2]
                k["octave"] = l; k["size"] = int(np.float32(31) * np.float32(self.scale[l]))
                xy = np.stack([k["x"], k["y"]], 1)
                t0 = clk()
                k["angle"] = port.ic_angles(img, xy, self.umax)                           # :76-103
                tp += clk() - t0
            kps_l.append(k)
        for l in range(self.nlevels):                                                     # blur + descriptors :1121-1135
            k = kps_l[l]
            if len(k) == 0:
                desc_l.append(np.zeros((0, 32), np.uint8))
                continue
            xya = np.stack([k["x"], k["y"], k["angle"]], 1)
            t0 = clk()
            work = cv2.GaussianBlur(pyr[l][EDGE:-EDGE, EDGE:-EDGE].copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)      # (:1132 clones the level)
            desc_l.append(port.descriptors(work, xya))
            tp += clk() - t0
        n = sum(len(k) for k in kps_l)                                                    # output assembly :1137-1167
        out_k, out_d = np.zeros(n, port.KP_DTYPE), np.zeros((n, 32), np.uint8)
        k = np.concatenate(kps_l) if n else out_k
        d = np.concatenate(desc_l) if n else out_d
        s = np.asarray(self.scale, np.float32)[k["octave"]] if n else np.zeros(0, np.float32)
        scaled = k["octave"] != 0
        k["x"] = np.where(scaled, k["x"] * s, k["x"]); k["y"] = np.where(scaled, k["y"] * s, k["y"])
        lap = (k["x"] >= lapping[0]) & (k["x"] <= lapping[1])
        mono = int((~lap).sum())
        out_k[:mono], out_d[:mono] = k[~lap], d[~lap]
        out_k[mono:], out_d[mono:] = k[lap][::-1], d[lap][::-1]                           # lapping key points are written from the back
        self.t_prim += tp
        return 0, out_k, out_d, mono


def _worker(args):
    frames, params, lapping = args
    cv2.setNumThreads(1)
    ex = Cv2Extractor(*params)
    t0 = time.perf_counter()
    for f in frames:
        ex.extract(f, lapping)
    return time.perf_counter() - t0, ex.t_prim


def rate(frames, params=(1000, 1.2, 8, 20, 7), lapping=(0, 1000), processes=1):
    """frames/s of the cv2-primitive extractor over `frames` ([n,h,w] uint8) on `processes` host processes (one frame stream each,
    cv2 single-threaded inside a process, like one extractor per thread in the reference); single process: in this process"""
    return rates(frames, params, lapping, processes)[0]


def rates(frames, params=(1000, 1.2, 8, 20, 7), lapping=(0, 1000), processes=1):
    """-> (frames/s as run, frames/s counting only the time inside the cv2 primitives and the C++ glue -- the rate a C++ build of
    the same calls would reach with no Python between them; both over `processes` concurrent workers).  The workers are fresh
    interpreters (subprocess, not fork: the caller may hold a CUDA context and helper threads); each times its own share after a
    warm-up frame, so interpreter start-up is not part of the number."""
    if processes <= 1:
        dt, tp = _worker((frames, params, lapping))
        return len(frames) / dt, len(frames) / tp
    import subprocess
    import sys
    import tempfile
    frames = np.ascontiguousarray(frames)
    processes = min(processes, len(frames))
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "frames.npy")
        np.save(path, frames)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        cmd = [sys.executable, "-m", "oracle.cv2_baseline", path, str(processes)] + [repr(p) for p in params] + [str(lapping[0]), str(lapping[1])]
        procs = [subprocess.Popen(cmd + [str(i)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, cwd=root, text=True) for i in range(processes)]
        res = []
        for pr in procs:
            out, _ = pr.communicate(timeout=600)
            n, dt, tp = out.split()[-3:]
            res.append((int(n), float(dt), float(tp)))
    # the workers run concurrently: the pool's rate is the sum of the per-worker rates
    return sum(n / max(dt, 1e-9) for n, dt, _ in res), sum(n / max(tp, 1e-9) for n, _, tp in res)


if __name__ == "__main__":      # worker: frames.npy processes nfeatures scale nlevels ini min lap0 lap1 index
    import sys
    a = sys.argv[1:]
    fr = np.load(a[0], mmap_mode="r")
    nproc, idx = int(a[1]), int(a[9])
    prm = (int(a[2]), float(a[3]), int(a[4]), int(a[5]), int(a[6]))
    lap = (int(a[7]), int(a[8]))
    mine = np.ascontiguousarray(fr[idx::nproc])
    cv2.setNumThreads(1)
    Cv2Extractor(*prm).extract(mine[0], lap)      # warm-up (library loading, first-call allocations)
    dt, tp = _worker((mine, prm, lap))
    print(len(mine), dt, tp)

"""ORACLE (test infrastructure, NOT product code): the reference's extractor control flow run through the
REAL OpenCV primitives (python cv2), so that the OpenCV-free port (orb_port.cpp) and the CUDA kernels can be
pinned against OpenCV itself.

Follows /root/reference/orb_slam3/src/ORBextractor.cc line by line:
  ComputePyramid           :1170-1195  -> cv2.resize(INTER_LINEAR) + cv2.copyMakeBorder(REFLECT_101)
  ComputeKeyPointsOctTree  :781-896    -> cv2.FastFeatureDetector per 35-px cell, ini/min threshold fallback
  DistributeOctTree        :555-779    -> oracle/orb_port.cpp port_distribute (real std::list + std::sort)
  IC_Angle                 :76-103     -> integer moments in numpy + cv2.fastAtan2
  operator()               :1086-1168  -> cv2.GaussianBlur(7x7, sigma 2, REFLECT_101) on a copy of the level,
                                          computeOrbDescriptor via port_descriptors (glibc cosf/sinf, un-fused)
and Frame.cc:1126-1151 (BFMatcher kNN-2) for the brute-force matcher.

The reference repository holds no golden vectors for this path; this module pins the OpenCV side (the build behind
python cv2, cv2.__version__ is recorded in every golden file), oracle/ref.py (the reference's own ORBextractor.cc,
compiled unmodified) pins the ORB-SLAM3 side of the extractor.  The matcher rows stay "parity unpinned" against the
reference repository (its matcher sources cannot be compiled here).
"""
import math

import cv2
import numpy as np

from . import port

cv2.setNumThreads(1)

EDGE_THRESHOLD = 19
PATCH_SIZE = 31
HALF_PATCH_SIZE = 15


class RefExtractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        # scale tables / features per level / umax come from the port (pure float/double bookkeeping,
        # ORBextractor.cc:409-469); tests/test_oracle_tables.py freezes them against SURVEY.md §8.
        self._p = port.PortExtractor(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        self.scale = self._p.scale_factors
        self.inv_scale = self._p.inv_scale_factors
        self.features_per_level = self._p.features_per_level
        self.umax = self._p.umax
        self._fast_ini = cv2.FastFeatureDetector_create(ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        self._fast_min = cv2.FastFeatureDetector_create(min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        self.pyramid = []       # bordered buffers
        self.raw = []
        self.selected = []
        self.blurred = []

    def level_roi(self, l):
        b = self.pyramid[l]
        return b[EDGE_THRESHOLD:-EDGE_THRESHOLD, EDGE_THRESHOLD:-EDGE_THRESHOLD]

    def compute_pyramid(self, image):
        self.pyramid = []
        h, w = image.shape
        for level in range(self.nlevels):
            s = np.float32(self.inv_scale[level])
            lw = port.lib().port_cv_round(float(np.float32(w) * s))
            lh = port.lib().port_cv_round(float(np.float32(h) * s))
            if level != 0:
                prev = np.ascontiguousarray(self.level_roi(level - 1))
                cur = cv2.resize(prev, (lw, lh), interpolation=cv2.INTER_LINEAR)
            else:
                cur = image
            self.pyramid.append(cv2.copyMakeBorder(cur, EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD,
                                                   cv2.BORDER_REFLECT_101))

    def compute_keypoints(self):
        self.raw, self.selected = [], []
        W = np.float32(35)
        for level in range(self.nlevels):
            img = self.level_roi(level)
            rows, cols = img.shape
            min_bx = min_by = EDGE_THRESHOLD - 3
            max_bx, max_by = cols - EDGE_THRESHOLD + 3, rows - EDGE_THRESHOLD + 3
            width, height = np.float32(max_bx - min_bx), np.float32(max_by - min_by)
            n_cols, n_rows = int(width / W), int(height / W)
            w_cell = int(math.ceil(np.float32(width / np.float32(n_cols))))
            h_cell = int(math.ceil(np.float32(height / np.float32(n_rows))))
            keys = []
            for i in range(n_rows):
                ini_y = min_by + i * h_cell
                max_y = ini_y + h_cell + 6
                if ini_y >= max_by - 3:
                    continue
                max_y = min(max_y, max_by)
                for j in range(n_cols):
                    ini_x = min_bx + j * w_cell
                    max_x = ini_x + w_cell + 6
                    if ini_x >= max_bx - 6:
                        continue
                    max_x = min(max_x, max_bx)
                    cell = np.ascontiguousarray(img[ini_y:max_y, ini_x:max_x])
                    kps = self._fast_ini.detect(cell, None)
                    if len(kps) == 0:
                        kps = self._fast_min.detect(cell, None)
                    for kp in kps:
                        keys.append((kp.pt[0] + j * w_cell, kp.pt[1] + i * h_cell, kp.response))
            raw = np.array(keys, np.float32).reshape(-1, 3)
            self.raw.append(raw)
            picked = port.distribute(raw, min_bx, max_bx, min_by, max_by, int(self.features_per_level[level]))
            sel = np.zeros(len(picked), port.KP_DTYPE)
            sel["x"] = raw[picked, 0] + min_bx
            sel["y"] = raw[picked, 1] + min_by
            sel["response"] = raw[picked, 2]
            sel["octave"] = level
            sel["size"] = int(np.float32(PATCH_SIZE) * np.float32(self.scale[level]))
            sel["angle"] = -1
            self.selected.append(sel)
        for level in range(self.nlevels):
            sel = self.selected[level]
            if len(sel):
                sel["angle"] = self.ic_angles(level, sel["x"], sel["y"])

    def ic_angles(self, level, xs, ys):
        """IC_Angle (:76-103) over the bordered buffer so that the full 31x31 disc is always addressable."""
        buf = self.pyramid[level].astype(np.int64)
        m01 = np.zeros(len(xs), np.int64)
        m10 = np.zeros(len(xs), np.int64)
        cx = np.rint(xs).astype(np.int64) + EDGE_THRESHOLD
        cy = np.rint(ys).astype(np.int64) + EDGE_THRESHOLD
        for v in range(-HALF_PATCH_SIZE, HALF_PATCH_SIZE + 1):
            d = int(self.umax[abs(v)])
            for u in range(-d, d + 1):
                val = buf[cy + v, cx + u]
                m10 += u * val
                m01 += v * val
        ang = cv2.fastAtan2  # scalar API: fastAtan2(y, x)
        return np.array([ang(float(np.float32(a)), float(np.float32(b))) for a, b in zip(m01, m10)], np.float32)

    def extract(self, image, lapping=(0, 0)):
        if image is None or image.size == 0:
            return -1, np.zeros(0, port.KP_DTYPE), np.zeros((0, 32), np.uint8), 0
        self.compute_pyramid(image)
        self.compute_keypoints()
        n = sum(len(s) for s in self.selected)
        out_k = np.zeros(n, port.KP_DTYPE)
        out_d = np.zeros((n, 32), np.uint8)
        mono, stereo = 0, n - 1
        self.blurred = [None] * self.nlevels
        for level in range(self.nlevels):
            sel = self.selected[level]
            if len(sel) == 0:
                continue
            work = np.ascontiguousarray(self.level_roi(level)).copy()
            work = cv2.GaussianBlur(work, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
            self.blurred[level] = work
            xya = np.stack([sel["x"], sel["y"], sel["angle"]], axis=1)
            desc = port.descriptors(work, xya)
            s = np.float32(self.scale[level])
            for i in range(len(sel)):
                kp = sel[i].copy()
                if level != 0:
                    kp["x"] = np.float32(kp["x"]) * s
                    kp["y"] = np.float32(kp["y"]) * s
                if lapping[0] <= kp["x"] <= lapping[1]:
                    pos = stereo
                    stereo -= 1
                else:
                    pos = mono
                    mono += 1
                out_k[pos] = kp
                out_d[pos] = desc[i]
        return 0, out_k, out_d, mono


def bf_knn2(query, train):
    """cv::BFMatcher(NORM_HAMMING).knnMatch(query, train, 2) (Frame.cc:1144) -> idx[nq,2], dist[nq,2];
    missing neighbours are (-1, INT_MAX)."""
    bf = cv2.BFMatcher(cv2.NORM_HAMMING)
    idx = np.full((len(query), 2), -1, np.int32)
    dist = np.full((len(query), 2), np.iinfo(np.int32).max, np.int32)
    if len(query) == 0 or len(train) == 0:
        return idx, dist
    for qi, ms in enumerate(bf.knnMatch(np.ascontiguousarray(query), np.ascontiguousarray(train), k=2)):
        for k, m in enumerate(ms):
            idx[qi, k] = m.trainIdx
            dist[qi, k] = int(m.distance)
    return idx, dist

"""ORACLE (test infrastructure, NOT product code): ctypes binding of oracle/orb_port.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  See orb_port.cpp's header for the parity-pinning statement.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4")])
assert KP_DTYPE.itemsize == 24


def build(force=False):
    """(Re)build liborbport.so with g++ when missing or older than its source."""
    so = _BUILD / "liborbport.so"
    src = _HERE / "orb_port.cpp"
    stale = (not so.exists()) or so.stat().st_mtime < src.stat().st_mtime
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", str(_HERE)] + (["-B"] if force else []))
    return so


_lib = None
_lib_fma = None
u8p = C.POINTER(C.c_uint8)


def _sig(lib):
    lib.port_create.restype = C.c_void_p
    lib.port_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    lib.port_destroy.argtypes = [C.c_void_p]
    lib.port_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    lib.port_extract.restype = C.c_int
    lib.port_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p,
                                 C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.port_extract_batch.restype = C.c_int
    lib.port_extract_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.port_level.restype = C.c_int
    lib.port_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                               C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
    lib.port_raw_count.restype = C.c_int
    lib.port_raw_count.argtypes = [C.c_void_p, C.c_int]
    lib.port_raw_keys.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.port_sel_count.restype = C.c_int
    lib.port_sel_count.argtypes = [C.c_void_p, C.c_int]
    lib.port_sel_keys.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.port_resize_linear.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
    lib.port_fast9.restype = C.c_int
    lib.port_fast9.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
    lib.port_gaussian7.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]
    lib.port_fast_atan2.restype = C.c_float
    lib.port_fast_atan2.argtypes = [C.c_float, C.c_float]
    lib.port_fast_atan2_array.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    lib.port_cv_round.restype = C.c_int
    lib.port_cv_round.argtypes = [C.c_float]
    lib.port_distribute.restype = C.c_int
    lib.port_distribute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.port_sort_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.port_ic_angle.restype = C.c_float
    lib.port_ic_angle.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
    lib.port_ic_angles.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.port_descriptors.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
    lib.port_hamming.restype = C.c_int
    lib.port_hamming.argtypes = [C.c_void_p, C.c_void_p]
    lib.port_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int]
    lib.port_best2_csr.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.port_bow_transform.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
    lib.port_search_area_best2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.port_distinctive.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.port_gray.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
    lib.port_stereo.restype = C.c_int
    lib.port_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def lib(fma=False):
    global _lib, _lib_fma
    if fma:
        if _lib_fma is None:
            build()
            _lib_fma = _sig(C.CDLL(str(_BUILD / "liborbport_fma.so")))
        return _lib_fma
    if _lib is None:
        _lib = _sig(C.CDLL(str(build())))
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8c(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


class PortExtractor:
    """Mirror of ORB_SLAM3::ORBextractor (ORBextractor.h:43-109) on the C++ port."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, fma=False):
        self._l = lib(fma)
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = self._l.port_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        sc = [np.zeros(nlevels, np.float32) for _ in range(4)]
        nf = np.zeros(nlevels, np.int32)
        um = np.zeros(16, np.int32)
        self._l.port_tables(self._h, _ptr(sc[0]), _ptr(sc[1]), _ptr(sc[2]), _ptr(sc[3]), _ptr(nf), _ptr(um))
        self.scale_factors, self.inv_scale_factors, self.level_sigma2, self.inv_level_sigma2 = sc
        self.features_per_level, self.umax = nf, um

    def __del__(self):
        try:
            self._l.port_destroy(self._h)
        except Exception:
            pass

    def extract(self, img, lapping=(0, 0)):
        """-> (rc, keypoints[KP_DTYPE], descriptors[n,32] u8, mono_index); rc=-1 on an empty image."""
        if img is None or img.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8), 0
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        h, w = img.shape
        cap = self.nfeatures + 8 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        rc = self._l.port_extract(self._h, _ptr(img), w, h, img.strides[0], int(lapping[0]), int(lapping[1]), _ptr(kps),
                                  _ptr(desc), cap, C.byref(n), C.byref(mono))
        return rc, kps[:n.value].copy(), desc[:n.value].copy(), mono.value

    def level(self, level, blurred=False, bordered=False):
        p, w, h, s = C.c_void_p(), C.c_int(), C.c_int(), C.c_size_t()
        rc = self._l.port_level(self._h, level, int(blurred), int(bordered), C.byref(p), C.byref(w), C.byref(h), C.byref(s))
        if rc:
            return None
        ww, hh = w.value, h.value
        if bordered and not blurred:
            ww, hh = ww + 38, hh + 38
        buf = (C.c_uint8 * (s.value * hh)).from_address(p.value)
        a = np.frombuffer(buf, np.uint8).reshape(hh, s.value)[:, :ww]
        return a.copy()

    def raw_keys(self, level):
        n = self._l.port_raw_count(self._h, level)
        a = np.zeros((n, 3), np.float32)
        if n:
            self._l.port_raw_keys(self._h, level, _ptr(a))
        return a

    def selected(self, level):
        n = self._l.port_sel_count(self._h, level)
        a = np.zeros(n, KP_DTYPE)
        if n:
            self._l.port_sel_keys(self._h, level, _ptr(a))
        return a


def extract_batch(images, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, lapping=(0, 0), nthreads=1,
                  with_data=True):
    """All-core CPU extraction of a [n,h,w] uint8 stack -> counts[n,2], kps[n,cap], desc[n,cap,32]."""
    images = np.ascontiguousarray(images, np.uint8)
    n, h, w = images.shape
    cap = nfeatures + 8 * nlevels + 64
    counts = np.zeros((n, 2), np.int32)
    kps = np.zeros((n, cap), KP_DTYPE) if with_data else None
    desc = np.zeros((n, cap, 32), np.uint8) if with_data else None
    rc = lib().port_extract_batch(nfeatures, scale_factor, nlevels, ini_th, min_th, _ptr(images), n, w, h, images.strides[1],
                                  images.strides[0], int(lapping[0]), int(lapping[1]), _ptr(kps) if with_data else None,
                                  _ptr(desc) if with_data else None, cap, _ptr(counts), nthreads)
    if rc:
        raise RuntimeError(f"port_extract_batch rc={rc}")
    return counts, kps, desc


def resize_linear(src, dw, dh):
    src = _u8c(src)
    dst = np.zeros((dh, dw), np.uint8)
    lib().port_resize_linear(_ptr(src), src.shape[1], src.shape[0], src.strides[0], _ptr(dst), dw, dh, dst.strides[0])
    return dst


def fast9(img, threshold):
    img = _u8c(img)
    cap = max(16, img.size // 2)
    out = np.zeros((cap, 3), np.float32)
    n = lib().port_fast9(_ptr(img), img.shape[1], img.shape[0], img.strides[0], threshold, _ptr(out), cap)
    return out[:n].copy()


def gaussian7(img):
    img = _u8c(img)
    dst = np.zeros_like(img)
    lib().port_gaussian7(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(dst), dst.strides[0])
    return dst


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(y)
    lib().port_fast_atan2_array(_ptr(y), _ptr(x), _ptr(out), y.size)
    return out


def distribute(xyr, min_x, max_x, min_y, max_y, n):
    xyr = np.ascontiguousarray(xyr, np.float32).reshape(-1, 3)
    cap = n + 64
    out = np.zeros(cap, np.int32)
    rc = lib().port_distribute(_ptr(xyr), len(xyr), min_x, max_x, min_y, max_y, n, _ptr(out), cap)
    if rc < 0:
        raise RuntimeError(f"port_distribute rc={rc}")
    return out[:rc].copy()


def sort_nodes(sizes, x0s):
    sizes = np.ascontiguousarray(sizes, np.int32)
    x0s = np.ascontiguousarray(x0s, np.int32)
    perm = np.zeros(len(sizes), np.int32)
    lib().port_sort_nodes(_ptr(sizes), _ptr(x0s), len(sizes), _ptr(perm))
    return perm


def ic_angle(img, x, y, umax):
    img = _u8c(img)
    um = np.ascontiguousarray(umax, np.int32)
    return float(lib().port_ic_angle(_ptr(img), img.strides[0], int(x), int(y), _ptr(um)))


def ic_angles(img, xy, umax):
    """IC_Angle of n key points of one level: img = the level (its parent buffer must extend 15 px beyond every key point), xy float32 [n,2]"""
    um = np.ascontiguousarray(umax, np.int32)
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.zeros(len(xy), np.float32)
    lib().port_ic_angles(C.c_void_p(img.ctypes.data), img.strides[0], _ptr(xy), len(xy), _ptr(um), _ptr(out))
    return out


def descriptors(blurred, xya, fma=False):
    blurred = _u8c(blurred)
    xya = np.ascontiguousarray(xya, np.float32).reshape(-1, 3)
    d = np.zeros((len(xya), 32), np.uint8)
    lib(fma).port_descriptors(_ptr(blurred), blurred.strides[0], _ptr(xya), len(xya), _ptr(d))
    return d


def hamming(a, b):
    a = _u8c(a)
    b = _u8c(b)
    return lib().port_hamming(_ptr(a), _ptr(b))


def knn2(q, db, nthreads=1):
    q = _u8c(q)
    db = _u8c(db)
    idx = np.zeros((len(q), 2), np.int32)
    dist = np.zeros((len(q), 2), np.int32)
    lib().port_knn2(_ptr(q), len(q), _ptr(db), len(db), _ptr(idx), _ptr(dist), nthreads)
    return idx, dist


def best2_csr(q, train, cand, rowptr, init=256):
    q = _u8c(q)
    train = _u8c(train)
    cand = np.ascontiguousarray(cand, np.int32)
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    out = np.zeros((len(q), 4), np.int32)
    lib().port_best2_csr(_ptr(q), len(q), _ptr(train), _ptr(cand), _ptr(rowptr), init, _ptr(out))
    return out


def bow_transform(vocab, desc, levelsup=4, norm=1):
    """DBoW2 transform of one descriptor set.  vocab = dict(child_begin, child_count, child_list, node_desc, node_weight,
    node_word, depth).  -> (bow_id, bow_val, fv_node, fv_start, fv_feat, n_valid)"""
    desc = _u8c(desc).reshape(-1, 32)
    n = len(desc)
    v = {k: np.ascontiguousarray(a) for k, a in vocab.items() if k != "depth"}
    bow_id, bow_val = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
    fv_node, fv_start, fv_feat = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
    counts = np.zeros(3, np.int32)
    lib().port_bow_transform(len(v["child_begin"]), _ptr(v["child_begin"]), _ptr(v["child_count"]), _ptr(v["child_list"]),
                             _ptr(v["node_desc"]), _ptr(v["node_weight"]), _ptr(v["node_word"]), int(vocab["depth"]), _ptr(desc), n,
                             levelsup, norm, _ptr(bow_id), _ptr(bow_val), _ptr(fv_node), _ptr(fv_start), _ptr(fv_feat), _ptr(counts))
    return bow_id[:counts[0]], bow_val[:counts[0]], fv_node[:counts[1]], fv_start[:counts[1]], fv_feat[:counts[2]], int(counts[2])


def search_area_best2(kps_xy, octaves, train, grid4, queries, qlev, qdesc, skip=None, u_right=None, init=256):
    """grid lookup (Frame::GetFeaturesInArea) + best/second scan (ORBmatcher::SearchByProjection) per query"""
    kps_xy = np.ascontiguousarray(kps_xy, np.float32).reshape(-1, 2)
    octaves = np.ascontiguousarray(octaves, np.int32)
    train = _u8c(train)
    grid4 = np.ascontiguousarray(grid4, np.float32)
    queries = np.ascontiguousarray(queries, np.float32).reshape(-1, 4)
    qlev = np.ascontiguousarray(qlev, np.int32).reshape(-1, 2)
    qdesc = _u8c(qdesc)
    out = np.zeros((len(queries), 4), np.int32)
    sk = None if skip is None else np.ascontiguousarray(skip, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    lib().port_search_area_best2(_ptr(kps_xy), _ptr(octaves), _ptr(train), len(kps_xy), _ptr(grid4), _ptr(queries), _ptr(qlev),
                                 _ptr(qdesc), len(queries), None if sk is None else _ptr(sk), None if ur is None else _ptr(ur),
                                 init, _ptr(out))
    return out


def distinctive(desc, rowptr):
    """MapPoint::ComputeDistinctiveDescriptors per CSR group -> best index within each group (-1 for empty groups)"""
    desc = _u8c(desc)
    rowptr = np.ascontiguousarray(rowptr, np.int32)
    best = np.zeros(len(rowptr) - 1, np.int32)
    lib().port_distinctive(_ptr(desc), _ptr(rowptr), len(best), _ptr(best))
    return best


def gray(img, rgb=True):
    """cv::cvtColor(..., COLOR_{RGB,BGR,RGBA,BGRA}2GRAY) on an [h,w,3|4] uint8 image"""
    img = _u8c(img)
    h, w, c = img.shape
    out = np.zeros((h, w), np.uint8)
    lib().port_gray(_ptr(img), w, h, img.strides[0], c, int(rgb), _ptr(out), out.strides[0])
    return out


def stereo(ext_l, ext_r, kps_l, desc_l, kps_r, desc_r, bf, b):
    """Frame::ComputeStereoMatches on the two port extractors' current pyramids.
    -> (uRight[nL] f32, depth[nL] f32, bestR[nL] i32, sad[nL] i32, n_kept)"""
    kl = np.ascontiguousarray(kps_l, KP_DTYPE)
    kr = np.ascontiguousarray(kps_r, KP_DTYPE)
    dl, dr = _u8c(desc_l), _u8c(desc_r)
    n = len(kl)
    ur = np.zeros(n, np.float32)
    dp = np.zeros(n, np.float32)
    br = np.zeros(n, np.int32)
    sad = np.zeros(n, np.int32)
    kept = lib().port_stereo(ext_l._h, ext_r._h, _ptr(kl), _ptr(dl), n, _ptr(kr), _ptr(dr), len(kr), bf, b, _ptr(ur),
                             _ptr(dp), _ptr(br), _ptr(sad))
    return ur, dp, br, sad, kept


# ---- input side ("next" rows): stereo rectification and keypoint undistortion -------------------------------------------
def remap_linear(src, map_x, map_y):
    """cv::remap(src, dst, map_x, map_y, cv::INTER_LINEAR) for uint8 single-channel images, CV_32FC1 maps, constant (0)
    border (reference System.cc:239-240).  Restatement of OpenCV's RemapInvoker + remapBilinear (imgproc/imgwarp.cpp):
    maps to fixed point with cvRound(m * 32); weights BilinearTab_i (shorts, (0,0) entry = {32767,0,0,1} after the
    table's saturation + sum fix-up); (sum + 2^14) >> 15.  Pinned to cv2.remap by tests/test_oracle_primitives.py."""
    src = np.ascontiguousarray(src, np.uint8)
    H, W = src.shape
    sx = np.rint(np.asarray(map_x, np.float32) * np.float32(32)).astype(np.int64)
    sy = np.rint(np.asarray(map_y, np.float32) * np.float32(32)).astype(np.int64)
    ix, iy = np.clip(sx >> 5, -32768, 32767), np.clip(sy >> 5, -32768, 32767)
    fx, fy = sx & 31, sy & 31
    w = [(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32]
    zero = (fx == 0) & (fy == 0)
    w[0] = np.where(zero, 32767, w[0])
    w[3] = np.where(zero, 1, w[3])

    def tap(y, x):
        ok = (x >= 0) & (x < W) & (y >= 0) & (y < H)
        return np.where(ok, src[np.clip(y, 0, H - 1), np.clip(x, 0, W - 1)].astype(np.int64), 0)

    acc = tap(iy, ix) * w[0] + tap(iy, ix + 1) * w[1] + tap(iy + 1, ix) * w[2] + tap(iy + 1, ix + 1) * w[3]
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


def undistort_points(xy, K4, dist, new_K4=None):
    """cv::undistortPoints(xy, out, K, dist, noArray(), newK) as Frame::UndistortKeyPoints calls it (Frame.cc:747-780):
    cvUndistortPointsInternal with the default criteria (5 iterations), double precision, un-fused; float32 result.
    K4 = (fx, fy, cx, cy); dist = k1 k2 p1 p2 [k3 ...].  Pinned to cv2.undistortPoints by tests/test_oracle_primitives.py."""
    xy = np.asarray(xy, np.float32).reshape(-1, 2)
    dist = np.asarray(dist, np.float32).ravel()
    if len(dist) == 0 or dist[0] == 0.0:                      # Frame.cc:749-753
        return xy.copy()
    new_K4 = K4 if new_K4 is None else new_K4
    fx, fy, cx, cy = [np.float64(np.float32(v)) for v in K4]
    nfx, nfy, ncx, ncy = [np.float64(np.float32(v)) for v in new_K4]
    k = np.zeros(12, np.float64)
    k[:len(dist)] = dist.astype(np.float64)
    u, v = xy[:, 0].astype(np.float64), xy[:, 1].astype(np.float64)
    ifx, ify = np.float64(1.) / fx, np.float64(1.) / fy
    x, y = (u - cx) * ifx, (v - cy) * ify
    x0, y0 = x.copy(), y.copy()
    live = np.ones(len(x), bool)
    for _ in range(5):
        r2 = x * x + y * y
        icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2)
        neg = live & (icdist < 0)
        dx = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2
        dy = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2
        nx, ny = (x0 - dx) * icdist, (y0 - dy) * icdist
        upd = live & ~neg
        x, y = np.where(upd, nx, x), np.where(upd, ny, y)
        x, y = np.where(neg, (u - cx) * ifx, x), np.where(neg, (v - cy) * ify, y)
        live &= ~neg
    xx, yy, ww = nfx * x + 0. * y + ncx, 0. * x + nfy * y + ncy, 1. / (0. * x + 0. * y + 1.)
    return np.stack([(xx * ww).astype(np.float32), (yy * ww).astype(np.float32)], 1)


def three_maxima(hist):
    """ORBmatcher::ComputeThreeMaxima (reference ORBmatcher.cc:2012-2053) on the bin sizes of a rotation histogram"""
    max1 = max2 = max3 = 0
    ind1 = ind2 = ind3 = -1
    for i in range(len(hist)):
        s = int(hist[i])
        if s > max1:
            max3, max2, max1 = max2, max1, s
            ind3, ind2, ind1 = ind2, ind1, i
        elif s > max2:
            max3, max2 = max2, s
            ind3, ind2 = ind2, i
        elif s > max3:
            max3, ind3 = s, i
    if np.float32(max2) < np.float32(0.1) * np.float32(max1):
        ind2 = ind3 = -1
    elif np.float32(max3) < np.float32(0.1) * np.float32(max1):
        ind3 = -1
    return ind1, ind2, ind3


def rotation_check(angle_a, angle_b):
    """the rotation-consistency filter of the match scans (reference ORBmatcher.cc:236, :345-352, :405-423 and
    ComputeThreeMaxima :2012-2053) for one set of matches -> (keep[n] bool, (ind1, ind2, ind3))"""
    a = np.asarray(angle_a, np.float32)
    b = np.asarray(angle_b, np.float32)
    factor = np.float32(1.0) / np.float32(30)
    rot = a - b
    rot = np.where(rot < 0, rot + np.float32(360.0), rot).astype(np.float32)
    x = (rot * factor).astype(np.float32)
    bins = np.floor(x.astype(np.float64) + 0.5).astype(np.int64)      # round(): half away from zero, x >= 0 here
    bins[bins == 30] = 0
    hist = np.bincount(bins, minlength=30)
    ind1, ind2, ind3 = three_maxima(hist)
    keep = (bins == ind1) | (bins == ind2) | (bins == ind3)
    return keep, (ind1, ind2, ind3)


def rgbd_stereo(kps_xy, depth, K4, dist, bf):
    """Frame::ComputeStereoFromRGBD (reference Frame.cc:984-1005): depth = float32 [h, w]; kps_xy = the (distorted) keypoints;
    the undistorted x comes from undistort_points (Frame::UndistortKeyPoints) -> (uRight[n], depth[n]) float32, -1 = none"""
    xy = np.asarray(kps_xy, np.float32).reshape(-1, 2)
    depth = np.asarray(depth, np.float32)
    xu = undistort_points(xy, K4, dist)[:, 0]
    d = depth[xy[:, 1].astype(np.int64), xy[:, 0].astype(np.int64)]       # Mat::at<float>(v, u) with float arguments: int conversion
    ok = d > 0
    ur = np.full(len(xy), -1, np.float32)
    dp = np.full(len(xy), -1, np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (np.float32(bf) / d).astype(np.float32)
    ur[ok] = (xu[ok] - q[ok]).astype(np.float32)
    dp[ok] = d[ok]
    return ur, dp

"""ORACLE (test infrastructure, NOT product code): ctypes binding of oracle/_ref/liborbref.so -- the reference's OWN
ORB_SLAM3::ORBextractor (orb_slam3/src/ORBextractor.cc) and its vendored DBoW2 vocabulary (orb_slam3/Thirdparty/DBoW2),
compiled unmodified from /root/reference by `make -C oracle ref` against the OpenCV stand-in of oracle/cvshim/, plus the
reference's own definitions of the stereo / matcher / grid functions of Frame.cc and ORBmatcher.cc (cut out at build time and
compiled inside stand-in classes, oracle/ref_cut_tu.cpp).

Only tests/, __graft_entry__ and bench.py's cpu_baseline / --impl reference legs may import this module.  The library is
built in the development container (where /root/reference exists) and travels to the GPU box as a built file; when it
is absent `available()` is False and the callers fall back to the port (tests skip).
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from .port import KP_DTYPE

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_ref" / "liborbref.so"
REFERENCE_SRC = Path("/root/reference/orb_slam3")
_lib = None


def build(force=False):
    """Compile the reference extractor when its sources are present (never on the GPU box); returns the .so path or None."""
    if (REFERENCE_SRC / "src" / "ORBextractor.cc").exists():
        subprocess.check_call(["make", "-s", "-C", str(_HERE), "ref"] + (["-B"] if force else []))
    return _SO if _SO.exists() else None


def available():
    return _SO.exists() or build() is not None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/liborbref.so is missing and /root/reference is not here to build it")
        l = C.CDLL(str(_SO))
        l.ref_create.restype = C.c_void_p
        l.ref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        l.ref_destroy.argtypes = [C.c_void_p]
        l.ref_tables.argtypes = [C.c_void_p] * 7
        l.ref_extract.restype = C.c_int
        l.ref_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]
        l.ref_level.restype = C.c_int
        l.ref_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_size_t)]
        l.ref_extract_batch.restype = C.c_int
        l.ref_extract_batch.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_size_t,
                                        C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        l.refbow_load_text.restype = C.c_void_p
        l.refbow_load_text.argtypes = [C.c_char_p]
        l.refbow_destroy.argtypes = [C.c_void_p]
        l.refbow_size.restype = C.c_int
        l.refbow_size.argtypes = [C.c_void_p]
        l.refbow_distance.restype = C.c_int
        l.refbow_distance.argtypes = [C.c_void_p, C.c_void_p]
        l.refbow_transform.restype = C.c_int
        l.refbow_transform.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 6
        l.refcut_constants.argtypes = [C.POINTER(C.c_int)] * 3
        l.refcut_descriptor_distance.restype = C.c_int
        l.refcut_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        l.refcut_three_maxima.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        l.refcut_stereo.restype = C.c_int
        l.refcut_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                    C.c_void_p, C.c_void_p]
        l.refcut_rgbd.restype = C.c_int
        l.refcut_rgbd.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_float, C.c_void_p, C.c_void_p]
        l.refcut_search_area_best2.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        l.refcut_search_by_projection.restype = C.c_int
        l.refcut_search_by_projection.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p]
        l.refcut_search_by_bow.restype = C.c_int
        l.refcut_search_by_bow.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 3 + \
            [C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
        l.refcut_distinctive.restype = C.c_int
        l.refcut_distinctive.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _lib = l
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class RefExtractor:
    """ORB_SLAM3::ORBextractor itself (ORBextractor.h:43-109); same Python surface as oracle.port.PortExtractor."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self._l = lib()
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = self._l.ref_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        sc = [np.zeros(nlevels, np.float32) for _ in range(4)]
        nf = np.zeros(nlevels, np.int32)
        um = np.zeros(16, np.int32)
        self._l.ref_tables(self._h, _ptr(sc[0]), _ptr(sc[1]), _ptr(sc[2]), _ptr(sc[3]), _ptr(nf), _ptr(um))
        self.scale_factors, self.inv_scale_factors, self.level_sigma2, self.inv_level_sigma2 = sc
        self.features_per_level, self.umax = nf, um

    def __del__(self):
        try:
            self._l.ref_destroy(self._h)
        except Exception:
            pass

    def extract(self, img, lapping=(0, 0)):
        """-> (rc, keypoints[KP_DTYPE], descriptors[n,32] u8, mono_index); rc=-1 on an empty image (ORBextractor.cc:1090)."""
        if img is None or img.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8), 0
        assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
        h, w = img.shape
        cap = self.nfeatures + 8 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        rc = self._l.ref_extract(self._h, _ptr(img), w, h, img.strides[0], int(lapping[0]), int(lapping[1]), _ptr(kps), _ptr(desc), cap,
                                 C.byref(n), C.byref(mono))
        return rc, kps[:n.value].copy(), desc[:n.value].copy(), mono.value

    def level(self, level, bordered=False):
        """mvImagePyramid[level] (bordered: with its 19-pixel apron)."""
        p, w, h, s = C.c_void_p(), C.c_int(), C.c_int(), C.c_size_t()
        if self._l.ref_level(self._h, level, int(bordered), C.byref(p), C.byref(w), C.byref(h), C.byref(s)):
            return None
        ww, hh = (w.value + 38, h.value + 38) if bordered else (w.value, h.value)
        buf = (C.c_uint8 * (s.value * hh)).from_address(p.value)
        return np.frombuffer(buf, np.uint8).reshape(hh, s.value)[:, :ww].copy()


def extract_batch(images, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, lapping=(0, 0), nthreads=1, with_data=True):
    """All-core run of the reference extractor over a [n,h,w] uint8 stack -> counts[n,2], kps[n,cap], desc[n,cap,32]."""
    images = np.ascontiguousarray(images, np.uint8)
    n, h, w = images.shape
    cap = nfeatures + 8 * nlevels + 64
    counts = np.zeros((n, 2), np.int32)
    kps = np.zeros((n, cap), KP_DTYPE) if with_data else None
    desc = np.zeros((n, cap, 32), np.uint8) if with_data else None
    rc = lib().ref_extract_batch(nfeatures, scale_factor, nlevels, ini_th, min_th, _ptr(images), n, w, h, images.strides[1], images.strides[0],
                                 int(lapping[0]), int(lapping[1]), _ptr(kps) if with_data else None, _ptr(desc) if with_data else None, cap,
                                 _ptr(counts), nthreads)
    if rc:
        raise RuntimeError(f"ref_extract_batch rc={rc}")
    return counts, kps, desc


def write_vocabulary_text(vocab, path):
    """A flat vocabulary (orb_slam3_ros_b200.bow.synthetic_vocabulary layout, children numbered after their parents) in the text
    format of ORBvoc.txt that DBoW2's loadFromTextFile reads (TemplatedVocabulary.h:1350-1434): header `k L scoring weighting`
    (L1_NORM = 0, TF_IDF = 0), then one line per node in id order: `parent isLeaf d0 .. d31 weight`."""
    cb, cc, cl = vocab["child_begin"], vocab["child_count"], vocab["child_list"]
    n = len(cb)
    parent = np.zeros(n, np.int64)
    for p in range(n):
        for c in cl[cb[p]:cb[p] + cc[p]]:
            assert c > p
            parent[c] = p
    lines = [f"{int(cc.max())} {int(vocab['depth'])} 0 0"]
    for i in range(1, n):
        leaf = int(cc[i] == 0)
        lines.append(f"{parent[i]} {leaf} " + " ".join(str(int(b)) for b in vocab["node_desc"][i]) + f" {float(vocab['node_weight'][i])!r}")
    with open(path, "w") as f:
        f.write("\n".join(lines))      # no trailing newline: the loader's `while(!f.eof())` would read one more, empty, node line


class RefVocabulary:
    """The reference's ORBVocabulary (vendored DBoW2 TemplatedVocabulary<FORB::TDescriptor, FORB>), loaded from a text file."""

    def __init__(self, path):
        self._l = lib()
        self._v = self._l.refbow_load_text(str(path).encode())
        if not self._v:
            raise RuntimeError(f"DBoW2 loadFromTextFile failed on {path}")
        self.words = self._l.refbow_size(self._v)

    def __del__(self):
        try:
            self._l.refbow_destroy(self._v)
        except Exception:
            pass

    def transform(self, desc, levelsup=4):
        """Frame::ComputeBoW (Frame.cc:738-745) -> (bow_id, bow_val, fv_node, fv_start, fv_feat, n_valid), all in std::map order"""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        bow_id, bow_val = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
        fv_node, fv_start, fv_feat = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
        counts = np.zeros(3, np.int32)
        self._l.refbow_transform(self._v, _ptr(desc), n, levelsup, _ptr(bow_id), _ptr(bow_val), _ptr(fv_node), _ptr(fv_start), _ptr(fv_feat),
                                 _ptr(counts))
        return bow_id[:counts[0]], bow_val[:counts[0]], fv_node[:counts[1]], fv_start[:counts[1]], fv_feat[:counts[2]], int(counts[2])


def descriptor_distance(a, b):
    """DBoW2::FORB::distance (FORB.cpp:79-97) -- the same bit trick as ORBmatcher::DescriptorDistance (ORBmatcher.cc:2058-2074)"""
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().refbow_distance(_ptr(a), _ptr(b))


# ---- the reference's own matcher / stereo / grid functions (oracle/ref_cut_tu.cpp: definitions cut out of Frame.cc and ORBmatcher.cc
# ---- at build time and compiled inside stand-in classes) ---------------------------------------------------------------------
def matcher_constants():
    """(TH_LOW, TH_HIGH, HISTO_LENGTH) as ORBmatcher.cc:35-37 defines them"""
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    lib().refcut_constants(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def matcher_descriptor_distance(a, b):
    """ORBmatcher::DescriptorDistance (ORBmatcher.cc:2058-2074)"""
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().refcut_descriptor_distance(_ptr(a), _ptr(b))


def three_maxima(counts):
    """ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:2012-2053) on a histogram given by its bin sizes -> (ind1, ind2, ind3)"""
    counts = np.ascontiguousarray(counts, np.int32)
    out = np.zeros(3, np.int32)
    lib().refcut_three_maxima(_ptr(counts), len(counts), _ptr(out))
    return tuple(int(v) for v in out)


def stereo(ext_l, ext_r, kps_l, desc_l, kps_r, desc_r, bf, b):
    """Frame::ComputeStereoMatches (Frame.cc:811-981) on two RefExtractors that have just extracted the two eyes -> (mvuRight, mvDepth)"""
    kl = np.ascontiguousarray(kps_l, KP_DTYPE)
    kr = np.ascontiguousarray(kps_r, KP_DTYPE)
    dl, dr = np.ascontiguousarray(desc_l, np.uint8), np.ascontiguousarray(desc_r, np.uint8)
    ur, dp = np.zeros(len(kl), np.float32), np.zeros(len(kl), np.float32)
    lib().refcut_stereo(ext_l._h, ext_r._h, _ptr(kl), _ptr(dl), len(kl), _ptr(kr), _ptr(dr), len(kr), bf, b, _ptr(ur), _ptr(dp))
    return ur, dp


def rgbd_stereo(kps_xy, x_undistorted, depth, bf):
    """Frame::ComputeStereoFromRGBD (Frame.cc:984-1005) -> (mvuRight, mvDepth)"""
    xy = np.ascontiguousarray(kps_xy, np.float32).reshape(-1, 2)
    xu = np.ascontiguousarray(x_undistorted, np.float32)
    depth = np.ascontiguousarray(depth, np.float32)
    ur, dp = np.zeros(len(xy), np.float32), np.zeros(len(xy), np.float32)
    lib().refcut_rgbd(_ptr(xy), _ptr(xu), len(xy), _ptr(depth), depth.shape[1], depth.shape[0], depth.strides[0] // 4, bf, _ptr(ur), _ptr(dp))
    return ur, dp


def search_area_best2(kps_xy, octaves, train, grid4, queries, qlev, qdesc, skip=None, u_right=None, init=256):
    """Frame::AssignFeaturesToGrid + GetFeaturesInArea (reference text) followed by the best / second scan; same arguments and
    result as oracle.port.search_area_best2"""
    kps_xy = np.ascontiguousarray(kps_xy, np.float32).reshape(-1, 2)
    octaves = np.ascontiguousarray(octaves, np.int32)
    train = np.ascontiguousarray(train, np.uint8)
    grid4 = np.ascontiguousarray(grid4, np.float32)
    queries = np.ascontiguousarray(queries, np.float32).reshape(-1, 4)
    qlev = np.ascontiguousarray(qlev, np.int32).reshape(-1, 2)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    out = np.zeros((len(queries), 4), np.int32)
    sk = None if skip is None else np.ascontiguousarray(skip, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    lib().refcut_search_area_best2(_ptr(kps_xy), _ptr(octaves), _ptr(train), len(kps_xy), _ptr(grid4), _ptr(queries), _ptr(qlev), _ptr(qdesc),
                                   len(queries), None if sk is None else _ptr(sk), None if ur is None else _ptr(ur), init, _ptr(out))
    return out


def distinctive(kf_desc, left_right, bad=None):
    """MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:329-403) for one map point: kf_desc [nkf, 2, 32] = two descriptor rows per
    observing key frame, left_right [nkf, 2] = the observation's (left, right) row indices (0 / 1, -1 = none), bad [nkf] = key frames
    flagged bad -> the chosen descriptor [32] or None (early return)"""
    kf_desc = np.ascontiguousarray(kf_desc, np.uint8).reshape(-1, 2, 32)
    lr = np.ascontiguousarray(left_right, np.int32).reshape(-1, 2)
    b = None if bad is None else np.ascontiguousarray(bad, np.uint8)
    out = np.zeros(32, np.uint8)
    ok = lib().refcut_distinctive(_ptr(kf_desc), _ptr(lr), None if b is None else _ptr(b), len(kf_desc), _ptr(out))
    return out if ok else None


def search_by_projection(kps_xy, octaves, train, grid4, scale_factors, proj, level, mp_desc, in_view=None, u_right=None, has_point=None,
                         nnratio=0.8, th=1.0):
    """ORBmatcher(nnratio).SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:43-213, reference text) on a frame with Nleft == -1:
    proj [nmp, 4] = {mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos}, level = mnTrackScaleLevel, has_point = key points that
    already hold a map point with observations -> (nmatches, match_of[n]: map point assigned to each key point by this call, -1 none)"""
    kps_xy = np.ascontiguousarray(kps_xy, np.float32).reshape(-1, 2)
    octaves = np.ascontiguousarray(octaves, np.int32)
    train = np.ascontiguousarray(train, np.uint8)
    grid4 = np.ascontiguousarray(grid4, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    proj = np.ascontiguousarray(proj, np.float32).reshape(-1, 4)
    level = np.ascontiguousarray(level, np.int32)
    mp_desc = np.ascontiguousarray(mp_desc, np.uint8)
    n, nmp = len(kps_xy), len(proj)
    iv = np.ones(nmp, np.uint8) if in_view is None else np.ascontiguousarray(in_view, np.uint8)
    ur = None if u_right is None else np.ascontiguousarray(u_right, np.float32)
    hp = None if has_point is None else np.ascontiguousarray(has_point, np.uint8)
    match_of = np.full(n, -1, np.int32)
    nm = lib().refcut_search_by_projection(_ptr(kps_xy), _ptr(octaves), _ptr(train), n, _ptr(grid4), None if ur is None else _ptr(ur),
                                           None if hp is None else _ptr(hp), _ptr(sf), len(sf), _ptr(proj), _ptr(level), _ptr(mp_desc), _ptr(iv),
                                           nmp, nnratio, th, _ptr(match_of))
    return nm, match_of


def search_by_projection_fisheye(f, mp, nnratio=0.8, th=1.0):
    """ORBmatcher(nnratio).SearchByProjection(F, vpMapPoints, th) (ORBmatcher.cc:43-213, reference text) on a fisheye-stereo frame (Nleft != -1).
      f:  dict(kps_l [nL,2], oct_l, kps_r [nR,2], oct_r, desc [nL+nR,32], fp (mnMinX, mnMaxX, mnMinY, mnMaxY, gridWInv, gridHInv), l2r [nL], r2l [nR],
          has_point [nL+nR], scale_factors)
      mp: dict(proj_l [m,3] (x, y, viewCos), level_l, in_view_l, proj_r [m,3], level_r (-1 none), in_view_r, desc [m,32])
    -> (nmatches, match_of[nL+nR])"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    kl, ol, kr, orr, d, fp = f32(f["kps_l"]).reshape(-1, 2), i32(f["oct_l"]), f32(f["kps_r"]).reshape(-1, 2), i32(f["oct_r"]), u8(f["desc"]), f32(f["fp"])
    l2r, r2l, hp, sf = i32(f["l2r"]), i32(f["r2l"]), u8(f["has_point"]), f32(f["scale_factors"])
    pl, ll, il, pr, lr, ir, md = f32(mp["proj_l"]).reshape(-1, 3), i32(mp["level_l"]), u8(mp["in_view_l"]), f32(mp["proj_r"]).reshape(-1, 3), \
        i32(mp["level_r"]), u8(mp["in_view_r"]), u8(mp["desc"])
    match_of = np.full(len(kl) + len(kr), -1, np.int32)
    fn = lib().refcut_search_by_projection_fisheye
    fn.restype = C.c_int
    fn.argtypes = FISHEYE_ARGTYPES
    nm = fn(_ptr(kl), _ptr(ol), len(kl), _ptr(kr), _ptr(orr), len(kr), _ptr(d), _ptr(fp), _ptr(l2r), _ptr(r2l), _ptr(hp), _ptr(sf), len(sf), _ptr(pl), _ptr(ll),
            _ptr(il), _ptr(pr), _ptr(lr), _ptr(ir), _ptr(md), len(pl), nnratio, th, _ptr(match_of))
    return nm, match_of


FISHEYE_ARGTYPES = [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_float, C.c_float, C.c_void_p]


def search_by_projection_motion(cur, last, th, mono, nnratio=0.9, check_orientation=True):
    """ORBmatcher(nnratio, checkOri).SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1676-1887, reference text;
    the call of Tracking::TrackWithMotionModel) for frames with Nleft == -1.
      cur:  dict(kps_xy [n,2], octaves, angles, desc [n,32], u_right (or None), state [n] (0 no map point / 1 one with observations /
            2 one without), fp = (mnMinX, mnMaxX, mnMinY, mnMaxY, gridWInv, gridHInv, mbf, mb), scale_factors, Tcw [12] = R row-major + t,
            cam4 = (fx, fy, cx, cy))
      last: dict(octaves, angles, state [m] (0 / 1 / 2 as above), outlier [m], pos [m,3], desc [m,32], Tlw [12])
    -> (nmatches, match_of[n]: last-frame feature whose map point the key point received in this call, -1 otherwise)"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    k, o, a, d = f32(cur["kps_xy"]).reshape(-1, 2), i32(cur["octaves"]), f32(cur["angles"]), u8(cur["desc"])
    n = len(k)
    ur = None if cur.get("u_right") is None else f32(cur["u_right"])
    cs, fp, sf, tcw, cam = u8(cur["state"]), f32(cur["fp"]), f32(cur["scale_factors"]), f32(cur["Tcw"]), f32(cur["cam4"])
    lo, la, ls, lout = i32(last["octaves"]), f32(last["angles"]), u8(last["state"]), u8(last["outlier"])
    lp, ld, tlw = f32(last["pos"]).reshape(-1, 3), u8(last["desc"]), f32(last["Tlw"])
    match_of = np.full(n, -1, np.int32)
    fn = lib().refcut_search_by_projection_motion
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 7 + \
                  [C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p]
    nm = fn(_ptr(k), _ptr(o), _ptr(a), _ptr(d), n, _ptr(fp), None if ur is None else _ptr(ur), _ptr(cs), _ptr(sf), len(sf), _ptr(tcw), _ptr(cam),
            len(lo), _ptr(lo), _ptr(la), _ptr(ls), _ptr(lout), _ptr(lp), _ptr(ld), _ptr(tlw), th, int(mono), nnratio, int(check_orientation),
            _ptr(match_of))
    return nm, match_of


def search_by_projection_motion_fisheye(cur, last, th, mono, nnratio=0.9, check_orientation=True):
    """ORBmatcher(nnratio, checkOri).SearchByProjection(CurrentFrame, LastFrame, th, bMono) (ORBmatcher.cc:1676-1887, reference text) on fisheye-stereo
    frames.  cur: dict(kps_l, oct_l, ang_l, kps_r, oct_r, ang_r, desc [nL+nR,32], state [nL+nR], fp (.., mbf, mb), scale_factors, Tcw [12], Trl [12],
    cam4); last: dict(n_left, octaves, angles, state, outlier, pos, desc, Tlw)  -> (nmatches, match_of[nL+nR])"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    kl, ol, al, kr, orr, ar = f32(cur["kps_l"]).reshape(-1, 2), i32(cur["oct_l"]), f32(cur["ang_l"]), f32(cur["kps_r"]).reshape(-1, 2), i32(cur["oct_r"]), f32(cur["ang_r"])
    d, fp, cs, sf, tcw, trl, cam = u8(cur["desc"]), f32(cur["fp"]), u8(cur["state"]), f32(cur["scale_factors"]), f32(cur["Tcw"]), f32(cur["Trl"]), f32(cur["cam4"])
    lo, la, ls, lout, lp, ld, tlw = i32(last["octaves"]), f32(last["angles"]), u8(last["state"]), u8(last["outlier"]), f32(last["pos"]).reshape(-1, 3), \
        u8(last["desc"]), f32(last["Tlw"])
    match_of = np.full(len(kl) + len(kr), -1, np.int32)
    fn = lib().refcut_search_by_projection_motion_fisheye
    fn.restype = C.c_int
    fn.argtypes = MOTION_FISHEYE_ARGTYPES
    nm = fn(_ptr(kl), _ptr(ol), _ptr(al), len(kl), _ptr(kr), _ptr(orr), _ptr(ar), len(kr), _ptr(d), _ptr(fp), _ptr(cs), _ptr(sf), len(sf), _ptr(tcw), _ptr(trl),
            _ptr(cam), len(lo), int(last["n_left"]), _ptr(lo), _ptr(la), _ptr(ls), _ptr(lout), _ptr(lp), _ptr(ld), _ptr(tlw), th, int(mono), nnratio,
            int(check_orientation), _ptr(match_of))
    return nm, match_of


MOTION_FISHEYE_ARGTYPES = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_int] + \
    [C.c_void_p] * 7 + [C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p]


def search_by_projection_reloc(cur, kf, th, orb_dist, nnratio=0.9, check_orientation=True):
    """ORBmatcher(nnratio, checkOri).SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (ORBmatcher.cc:1889-2010, reference
    text; the refinement calls of Tracking::Relocalization, Tracking.cc:3765 / :3779).
      cur: dict(kps_xy, octaves, angles, desc, holds [n] (key point already holds a map point), fp = (mnMinX, mnMaxX, mnMinY, mnMaxY, gridWInv,
           gridHInv, 0, 0, mnScaleLevels, mfLogScaleFactor), scale_factors, Tcw [12], cam4)
      kf:  dict(angles, state [m] (0 no map point / 1 good / 2 bad / 3 already found), pos [m,3], desc [m,32], min_dist [m], max_dist [m])
    -> (nmatches, match_of[n]: key-frame feature whose map point the key point received, -1 otherwise)"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    k, o, a, d = f32(cur["kps_xy"]).reshape(-1, 2), i32(cur["octaves"]), f32(cur["angles"]), u8(cur["desc"])
    holds, fp, sf, tcw, cam = u8(cur["holds"]), f32(cur["fp"]), f32(cur["scale_factors"]), f32(cur["Tcw"]), f32(cur["cam4"])
    ka, ks, kp, kd, kmin, kmax = f32(kf["angles"]), u8(kf["state"]), f32(kf["pos"]).reshape(-1, 3), u8(kf["desc"]), f32(kf["min_dist"]), f32(kf["max_dist"])
    match_of = np.full(len(k), -1, np.int32)
    fn = lib().refcut_search_by_projection_reloc
    fn.restype = C.c_int
    fn.argtypes = RELOC_ARGTYPES
    nm = fn(_ptr(k), _ptr(o), _ptr(a), _ptr(d), len(k), _ptr(fp), _ptr(holds), _ptr(sf), len(sf), _ptr(tcw), _ptr(cam), len(ka), _ptr(ka), _ptr(ks),
            _ptr(kp), _ptr(kd), _ptr(kmin), _ptr(kmax), th, int(orb_dist), nnratio, int(check_orientation), _ptr(match_of))
    return nm, match_of


RELOC_ARGTYPES = [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 6 + \
    [C.c_float, C.c_int, C.c_float, C.c_int, C.c_void_p]


def search_by_projection_sim3(kf, pts, sim3, th, ratio_hamming, with_kfs=False):
    """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (ORBmatcher.cc:427-530, reference text; LoopClosing.cc:1795 /
    :1982) with the reference's KeyFrame::GetFeaturesInArea / IsInImage and MapPoint::PredictScale(float, KeyFrame*).
      kf:  dict(kps_xy, octaves, desc, held [n] (matched already), fp = (mnMinX, mnMaxX, mnMinY, mnMaxY, gridWInv, gridHInv, 0, 0, mnScaleLevels,
           mfLogScaleFactor), scale_factors, cam4)
      pts: dict(state [m] (1 good / 2 bad), pos [m,3], normal [m,3], desc [m,32], min_dist [m], max_dist [m]);  sim3 [13] = R row-major, t, s
    -> (nmatches, match_of[n]: map point received by each key point in this call, -1 otherwise); with_kfs: the overload with vpPointsKFs /
    vpMatchedKF (ORBmatcher.cc:532-646, LoopClosing.cc:1773; point j comes from key frame j % 7) -> (nmatches, match_of, match_kf)"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    k, o, d, held, fp, sf, cam = f32(kf["kps_xy"]).reshape(-1, 2), i32(kf["octaves"]), u8(kf["desc"]), u8(kf["held"]), f32(kf["fp"]), f32(kf["scale_factors"]), f32(kf["cam4"])
    ps, pp, pn, pd, pmin, pmax = u8(pts["state"]), f32(pts["pos"]).reshape(-1, 3), f32(pts["normal"]).reshape(-1, 3), u8(pts["desc"]), f32(pts["min_dist"]), f32(pts["max_dist"])
    s3 = f32(sim3)
    match_of = np.full(len(k), -1, np.int32)
    args = (_ptr(k), _ptr(o), _ptr(d), len(k), _ptr(fp), _ptr(held), _ptr(sf), len(sf), _ptr(s3), _ptr(cam), len(ps), _ptr(ps), _ptr(pp), _ptr(pn), _ptr(pd),
            _ptr(pmin), _ptr(pmax), int(th), float(ratio_hamming), _ptr(match_of))
    if with_kfs:
        match_kf = np.full(len(k), -1, np.int32)
        fn = lib().refcut_search_by_projection_sim3_kfs
        fn.restype = C.c_int
        fn.argtypes = SIM3_ARGTYPES + [C.c_void_p]
        return fn(*args, _ptr(match_kf)), match_of, match_kf
    fn = lib().refcut_search_by_projection_sim3
    fn.restype = C.c_int
    fn.argtypes = SIM3_ARGTYPES
    return fn(*args), match_of


def fuse_sim3(kf, pts, sim3, th):
    """ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (ORBmatcher.cc:1340-1455, reference text; LoopClosing::SearchAndFuse).  Arguments as
    search_by_projection_sim3, kf["held"]: 0 none / 1 a good map point / 2 a bad one
    -> (nFused, replace_of[m]: key point whose map point replaces point j, added_at[m]: key point that received point j as an observation)"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    k, o, d, held, fp, sf, cam = f32(kf["kps_xy"]).reshape(-1, 2), i32(kf["octaves"]), u8(kf["desc"]), u8(kf["held"]), f32(kf["fp"]), f32(kf["scale_factors"]), f32(kf["cam4"])
    ps, pp, pn, pd, pmin, pmax = u8(pts["state"]), f32(pts["pos"]).reshape(-1, 3), f32(pts["normal"]).reshape(-1, 3), u8(pts["desc"]), f32(pts["min_dist"]), f32(pts["max_dist"])
    s3 = f32(sim3)
    rep, add = np.full(len(ps), -1, np.int32), np.full(len(ps), -1, np.int32)
    fn = lib().refcut_fuse_sim3
    fn.restype = C.c_int
    fn.argtypes = FUSE_SIM3_ARGTYPES
    nf = fn(_ptr(k), _ptr(o), _ptr(d), len(k), _ptr(fp), _ptr(held), _ptr(sf), len(sf), _ptr(s3), _ptr(cam), len(ps), _ptr(ps), _ptr(pp), _ptr(pn), _ptr(pd),
            _ptr(pmin), _ptr(pmax), float(th), _ptr(rep), _ptr(add))
    return nf, rep, add


def search_for_triangulation(k1, k2, common, only_stereo=False, coarse=False, check_orientation=False):
    """ORBmatcher(0.6, checkOri).SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (ORBmatcher.cc:906-1146, reference text;
    LocalMapping::CreateNewMapPoints).  k1 / k2: dict(kps_xy, octaves, angles, desc, has_point, u_right (or None), fv, Tcw [12]); common:
    dict(sigma2, scale_factors, cam4)  -> (nmatches, pairs [nmatches, 2])"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    keep, args = [], []
    for k in (k1, k2):
        arrs = [f32(k["kps_xy"]).reshape(-1, 2), i32(k["octaves"]), f32(k["angles"]), u8(k["desc"]), u8(k["has_point"]),
                None if k.get("u_right") is None else f32(k["u_right"])]
        fv = [i32(a) for a in k["fv"]]
        t = f32(k["Tcw"])
        keep += arrs + fv + [t]
        args += [None if a is None else _ptr(a) for a in arrs] + [len(arrs[0])] + [_ptr(a) for a in fv] + [len(fv[0]), len(fv[2]), _ptr(t)]
    sg, sf, cam = f32(common["sigma2"]), f32(common["scale_factors"]), f32(common["cam4"])
    pairs = np.full((len(keep[0]), 2), -1, np.int32)
    fn = lib().refcut_search_for_triangulation
    fn.restype = C.c_int
    fn.argtypes = TRI_ARGTYPES
    nm = fn(*args, _ptr(sg), _ptr(sf), len(sf), _ptr(cam), int(only_stereo), int(coarse), int(check_orientation), _ptr(pairs))
    return nm, pairs[:max(nm, 0)]


_TRI_KF = [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_void_p]
TRI_ARGTYPES = _TRI_KF + _TRI_KF + [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]


def fuse_kf(kf, pts, th):
    """ORBmatcher::Fuse(pKF, vpMapPoints, th) (ORBmatcher.cc:1148-1338, reference text; LocalMapping::SearchInNeighbors) on a key frame with
    NLeft == -1.  kf: dict(kps_xy, octaves, desc, held [n] (0 none / 1 good / 2 bad), held_obs [n], u_right (or None), inv_sigma2, fp (.., mbf at
    [6], mnScaleLevels, mfLogScaleFactor), scale_factors, Tcw [12], cam4); pts: dict(state [m] (0 null entry / 1 good / 2 bad), obs [m], pos, normal,
    desc, min_dist, max_dist)  -> (nFused, kp_holds[n], own_bad[n], pt_bad[m])"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    k, o, d, held, hobs = f32(kf["kps_xy"]).reshape(-1, 2), i32(kf["octaves"]), u8(kf["desc"]), u8(kf["held"]), i32(kf["held_obs"])
    ur = None if kf.get("u_right") is None else f32(kf["u_right"])
    isg, fp, sf, tcw, cam = f32(kf["inv_sigma2"]), f32(kf["fp"]), f32(kf["scale_factors"]), f32(kf["Tcw"]), f32(kf["cam4"])
    ps, pobs, pp, pn, pd, pmin, pmax = u8(pts["state"]), i32(pts["obs"]), f32(pts["pos"]).reshape(-1, 3), f32(pts["normal"]).reshape(-1, 3), u8(pts["desc"]), \
        f32(pts["min_dist"]), f32(pts["max_dist"])
    holds, own_bad, pt_bad = np.full(len(k), -1, np.int32), np.zeros(len(k), np.uint8), np.zeros(len(ps), np.uint8)
    fn = lib().refcut_fuse_kf
    fn.restype = C.c_int
    fn.argtypes = FUSE_KF_ARGTYPES
    nf = fn(_ptr(k), _ptr(o), _ptr(d), len(k), _ptr(fp), _ptr(held), _ptr(hobs), None if ur is None else _ptr(ur), _ptr(isg), _ptr(sf), len(sf), _ptr(tcw),
            _ptr(cam), len(ps), _ptr(ps), _ptr(pobs), _ptr(pp), _ptr(pn), _ptr(pd), _ptr(pmin), _ptr(pmax), float(th), _ptr(holds), _ptr(own_bad), _ptr(pt_bad))
    return nf, holds, own_bad, pt_bad


FUSE_KF_ARGTYPES = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 7 + [C.c_float] + [C.c_void_p] * 3
FUSE_SIM3_ARGTYPES = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 6 + [C.c_float, C.c_void_p, C.c_void_p]
SIM3_ARGTYPES = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int, C.c_float, C.c_void_p]


def search_by_bow_kf(angle1, desc1, state1, fv1, angle2, desc2, state2, fv2, nnratio=0.8, check_orientation=True):
    """ORBmatcher(nnratio, checkOri).SearchByBoW(pKF1, pKF2, vpMatches12) (ORBmatcher.cc:765-905, reference text; LoopClosing.cc:1680) for two
    monocular key frames; state: 0 no map point / 1 good / 2 bad; fv = (node, start, feat) arrays
    -> (nmatches, match_of[n1]: feature of key frame 2 whose map point feature i of key frame 1 was matched to, -1 none)"""
    a1, a2 = np.ascontiguousarray(angle1, np.float32), np.ascontiguousarray(angle2, np.float32)
    d1, d2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    s1, s2 = np.ascontiguousarray(state1, np.uint8), np.ascontiguousarray(state2, np.uint8)
    n1_, st1, f1_ = (np.ascontiguousarray(a, np.int32) for a in fv1)
    n2_, st2, f2_ = (np.ascontiguousarray(a, np.int32) for a in fv2)
    match_of = np.full(len(a1), -1, np.int32)
    fn = lib().refcut_search_by_bow_kf
    fn.restype = C.c_int
    fn.argtypes = BOW_KF_ARGTYPES
    nm = fn(_ptr(a1), _ptr(d1), _ptr(s1), len(a1), _ptr(n1_), _ptr(st1), _ptr(f1_), len(n1_), len(f1_), _ptr(a2), _ptr(d2), _ptr(s2), len(a2), _ptr(n2_),
            _ptr(st2), _ptr(f2_), len(n2_), len(f2_), nnratio, int(check_orientation), _ptr(match_of))
    return nm, match_of


BOW_KF_ARGTYPES = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + \
    [C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]


def search_for_initialization(f1, f2, prev, window=100, nnratio=0.9, check_orientation=True):
    """ORBmatcher(nnratio, checkOri).SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (ORBmatcher.cc:648-766,
    reference text; the call of Tracking::MonocularInitialization, Tracking.cc:2527).
      f1: dict(octaves, angles, desc [n1,32]);  f2: dict(kps_xy [n2,2], octaves, angles, desc [n2,32], fp = (mnMinX, mnMaxX, mnMinY, mnMaxY,
      gridWInv, gridHInv));  prev [n1,2] = vbPrevMatched
    -> (nmatches, matches12[n1], prev after the call [n1,2])"""
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    u8 = lambda a: np.ascontiguousarray(a, np.uint8)
    i32 = lambda a: np.ascontiguousarray(a, np.int32)
    o1, a1, d1 = i32(f1["octaves"]), f32(f1["angles"]), u8(f1["desc"])
    k2, o2, a2, d2, fp = f32(f2["kps_xy"]).reshape(-1, 2), i32(f2["octaves"]), f32(f2["angles"]), u8(f2["desc"]), f32(f2["fp"])
    pv = f32(prev).reshape(-1, 2).copy()
    m12 = np.full(len(o1), -1, np.int32)
    fn = lib().refcut_search_for_initialization
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 2 + [C.c_int, C.c_float, C.c_int, C.c_void_p]
    nm = fn(_ptr(o1), _ptr(a1), _ptr(d1), len(o1), _ptr(k2), _ptr(o2), _ptr(a2), _ptr(d2), len(o2), _ptr(fp), _ptr(pv), int(window), nnratio,
            int(check_orientation), _ptr(m12))
    return nm, m12, pv


def search_by_bow(kf_angle, kf_desc, kf_has_point, kf_fv, f_angle, f_desc, f_fv, nnratio=0.7, check_orientation=True):
    """ORBmatcher(nnratio, checkOri).SearchByBoW(pKF, F, vpMapPointMatches) (ORBmatcher.cc:223-421, reference text) for a monocular
    pair; kf_fv / f_fv = (node, start, feat) arrays of the two feature vectors (RefVocabulary.transform()[2:5] or the port's)
    -> (nmatches, match_of[nF]: key-frame feature whose map point was matched to each frame feature, -1 none)"""
    ka, fa = np.ascontiguousarray(kf_angle, np.float32), np.ascontiguousarray(f_angle, np.float32)
    kd, fd = np.ascontiguousarray(kf_desc, np.uint8), np.ascontiguousarray(f_desc, np.uint8)
    hp = np.ascontiguousarray(kf_has_point, np.uint8)
    kn, ks, kf_ = (np.ascontiguousarray(a, np.int32) for a in kf_fv)
    fn, fs, ff = (np.ascontiguousarray(a, np.int32) for a in f_fv)
    match_of = np.full(len(fa), -1, np.int32)
    nm = lib().refcut_search_by_bow(_ptr(ka), _ptr(kd), _ptr(hp), len(ka), _ptr(kn), _ptr(ks), _ptr(kf_), len(kn), len(kf_), _ptr(fa), _ptr(fd), len(fa),
                                    _ptr(fn), _ptr(fs), _ptr(ff), len(fn), len(ff), nnratio, int(check_orientation), _ptr(match_of))
    return nm, match_of

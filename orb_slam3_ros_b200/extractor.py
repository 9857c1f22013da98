"""Python mirror of ORB_SLAM3::ORBextractor (reference orb_slam3/include/ORBextractor.h:43-109) on liborbb200.so.

Used by the parity tests and bench.py; the C++ adapter with the same interface for the SLAM code itself lives in
orb_slam3_ros_b200/host/ORBextractor.{h,cc}.  Every method goes through the C ABI -- there is no CPU path here.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import KP_DTYPE, OrbbError


class ORBextractor:
    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th_fast=20, min_th_fast=7, device=0, max_batch=1):
        self._lib = capi.load()
        self._h = C.c_void_p()
        prm = capi.Params(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, device, max_batch)
        capi.check(self._lib.orbb_create(C.byref(prm), C.byref(self._h)))
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.device = device
        t = [np.zeros(nlevels, np.float32) for _ in range(4)]
        nf = np.zeros(nlevels, np.int32)
        capi.check(self._lib.orbb_get_tables(self._h, *[capi.ptr(a) for a in t], capi.ptr(nf)), self._h)
        self._scale, self._inv_scale, self._sigma2, self._inv_sigma2 = t
        self.features_per_level = nf

    def close(self):
        if getattr(self, "_h", None):
            self._free_staging()
            self._lib.orbb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- ORBextractor.h:61-82 getters ----
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return float(self._scale[1]) if self.nlevels > 1 else 1.0

    def GetScaleFactors(self):
        return self._scale.copy()

    def GetInverseScaleFactors(self):
        return self._inv_scale.copy()

    def GetScaleSigmaSquares(self):
        return self._sigma2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._inv_sigma2.copy()

    @property
    def max_keypoints(self):
        return self._lib.orbb_max_keypoints(self._h)

    # ---- operator() (ORBextractor.cc:1086-1168) ----
    def __call__(self, image, mask=None, lapping=(0, 0)):
        """-> (mono_index, keypoints[KP_DTYPE], descriptors[n,32]); returns (-1, empty, empty) on an empty image
        exactly like the reference.  `mask` is ignored (as in the reference)."""
        if image is None or image.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise TypeError("image must be CV_8UC1 (2-D uint8)")          # assert at ORBextractor.cc:1094
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        h, w = image.shape
        cap, kps, desc = self._staging()
        n, mono = C.c_int(0), C.c_int(0)
        rc = self._lib.orbb_extract(self._h, capi.ptr(image), w, h, image.strides[0], int(lapping[0]), int(lapping[1]),
                                    capi.ptr(kps), capi.ptr(desc), cap, C.byref(n), C.byref(mono))
        if rc == capi.ORBB_ERR_CAPACITY:      # the plan for this image size allows more keypoints than the estimate
            cap, kps, desc = self._staging()
            rc = self._lib.orbb_extract(self._h, capi.ptr(image), w, h, image.strides[0], int(lapping[0]), int(lapping[1]),
                                        capi.ptr(kps), capi.ptr(desc), cap, C.byref(n), C.byref(mono))
        capi.check(rc, self._h)
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def _staging(self):
        """result staging in pinned host memory (orbb_host_alloc), as the C++ adapter keeps it: the call's device-to-host copies
        are then asynchronous instead of going through the driver's pageable path"""
        cap = self.max_keypoints
        if cap > getattr(self, "_stage_cap", 0):
            self._free_staging()
            self._stage_ptrs = [self._lib.orbb_host_alloc(cap * KP_DTYPE.itemsize), self._lib.orbb_host_alloc(cap * 32)]
            if not all(self._stage_ptrs):
                raise MemoryError("orbb_host_alloc failed")
            self._stage_kps = np.frombuffer((C.c_uint8 * (cap * KP_DTYPE.itemsize)).from_address(self._stage_ptrs[0]), KP_DTYPE)
            self._stage_desc = np.frombuffer((C.c_uint8 * (cap * 32)).from_address(self._stage_ptrs[1]), np.uint8).reshape(cap, 32)
            self._stage_cap = cap
        return self._stage_cap, self._stage_kps, self._stage_desc

    def _free_staging(self):
        for p in getattr(self, "_stage_ptrs", []):
            if p:
                self._lib.orbb_host_free(p)
        self._stage_ptrs, self._stage_cap = [], 0
        self._stage_kps = self._stage_desc = None

    def extract_color(self, image, rgb=True, lapping=(0, 0)):
        """colour frame [h,w,3|4] uint8: cv::cvtColor(..2GRAY) on the device (Tracking.cc:1498-1525), then operator()"""
        if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] not in (3, 4):
            raise TypeError("image must be [h, w, 3|4] uint8")
        image = np.ascontiguousarray(image)
        h, w, c = image.shape
        cap = self.max_keypoints
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n, mono = C.c_int(0), C.c_int(0)
        capi.check(self._lib.orbb_extract_color(self._h, capi.ptr(image), w, h, image.strides[0], c, int(rgb), int(lapping[0]),
                                                int(lapping[1]), capi.ptr(kps), capi.ptr(desc), cap, C.byref(n), C.byref(mono)), self._h)
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_resized(self, image, new_size, lapping=(0, 0)):
        """cv::resize(image, new_size=(width, height)) on the device (System.cc:241-244), then operator()"""
        image = np.ascontiguousarray(image, np.uint8)
        h, w = image.shape
        n, mono = C.c_int(0), C.c_int(0)
        for _ in range(2):                      # a second pass if the plan for this size allows more keypoints than the estimate
            cap = self.max_keypoints
            kps = np.zeros(cap, KP_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            rc = self._lib.orbb_extract_resized(self._h, capi.ptr(image), w, h, image.strides[0], int(new_size[0]), int(new_size[1]),
                                                int(lapping[0]), int(lapping[1]), capi.ptr(kps), capi.ptr(desc), cap, C.byref(n), C.byref(mono))
            if rc != capi.ORBB_ERR_CAPACITY:
                break
        capi.check(rc, self._h)
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch_resized_device(self, dev_ptr, nframes, width, height, new_size, row_stride=None, frame_stride=None, lapping=(0, 0)):
        row_stride = row_stride or width
        frame_stride = frame_stride or row_stride * height
        capi.check(self._lib.orbb_extract_batch_resized(self._h, capi.ptr(dev_ptr), nframes, width, height, row_stride, frame_stride,
                                                        int(new_size[0]), int(new_size[1]), int(lapping[0]), int(lapping[1])), self._h)

    def image_pyramid(self, level, with_border=False):
        """mvImagePyramid[level] of the last frame (ORBextractor.h:84), as a numpy copy."""
        p, w, h, s = C.c_void_p(), C.c_int(), C.c_int(), C.c_size_t()
        capi.check(self._lib.orbb_pyramid_level(self._h, level, C.byref(p), C.byref(w), C.byref(h), C.byref(s)), self._h)
        b = 19 if with_border else 0
        base = p.value - b * s.value - b
        rows = h.value + 2 * b
        buf = (C.c_uint8 * (s.value * rows)).from_address(base)
        return np.frombuffer(buf, np.uint8).reshape(rows, s.value)[:, :w.value + 2 * b].copy()

    # ---- batched paths ----
    def extract_batch_device(self, dev_ptr, nframes, width, height, row_stride=None, frame_stride=None, lapping=(0, 0)):
        """frames already resident in device memory (int address or torch CUDA tensor); asynchronous."""
        row_stride = row_stride or width
        frame_stride = frame_stride or row_stride * height
        capi.check(self._lib.orbb_extract_batch(self._h, capi.ptr(dev_ptr), nframes, width, height, row_stride, frame_stride,
                                                int(lapping[0]), int(lapping[1])), self._h)

    def sync(self):
        capi.check(self._lib.orbb_sync(self._h), self._h)

    def fetch(self, nframes, with_data=True):
        """-> counts[nframes,2] (n, mono_index), kps[nframes,cap], desc[nframes,cap,32] of the last batch"""
        cap = self.max_keypoints
        counts = np.zeros((nframes, 2), np.int32)
        kps = np.zeros((nframes, cap), KP_DTYPE) if with_data else None
        desc = np.zeros((nframes, cap, 32), np.uint8) if with_data else None
        capi.check(self._lib.orbb_batch_fetch(self._h, nframes, capi.ptr(kps), capi.ptr(desc), cap, capi.ptr(counts)), self._h)
        return counts, kps, desc

    def extract_batch_host(self, images, lapping=(0, 0), out=None):
        """images: uint8 [n, h, w] in host memory (pinned = asynchronous copy).  H2D + kernels + D2H, synchronous.
        `out` = (kps, desc, counts) preallocated host arrays to reuse (e.g. pinned)."""
        n, h, w = images.shape
        cap = self.max_keypoints
        if out is None:
            out = (np.zeros((n, cap), KP_DTYPE), np.zeros((n, cap, 32), np.uint8), np.zeros((n, 2), np.int32))
        kps, desc, counts = out
        rc = self._lib.orbb_extract_batch_host(self._h, capi.ptr(images), n, w, h, images.strides[1], images.strides[0],
                                               int(lapping[0]), int(lapping[1]), capi.ptr(kps), capi.ptr(desc), kps.shape[1],
                                               capi.ptr(counts))
        capi.check(rc, self._h)
        return counts, kps, desc

    def submit_batch_host(self, images, lapping=(0, 0), out=None):
        """asynchronous half of extract_batch_host: enqueue upload + kernels + download; call wait_batch_host() later.
        images / out must stay alive (and should be pinned) until then."""
        n, h, w = images.shape
        kps, desc, counts = out
        self._pending = (images, out)
        capi.check(self._lib.orbb_extract_batch_host_submit(self._h, capi.ptr(images), n, w, h, images.strides[1], images.strides[0],
                                                            int(lapping[0]), int(lapping[1]), capi.ptr(kps), capi.ptr(desc),
                                                            kps.shape[1]), self._h)

    def wait_batch_host(self):
        images, (kps, desc, counts) = self._pending
        capi.check(self._lib.orbb_extract_batch_host_wait(self._h, capi.ptr(counts)), self._h)
        self._pending = None
        return counts, kps, desc

    def set_profiling(self, on=True):
        capi.check(self._lib.orbb_set_profiling(self._h, int(on)), self._h)

    def stage_times(self):
        ms = np.zeros(16, np.float32)
        n = self._lib.orbb_stage_times(self._h, capi.ptr(ms), 16)
        return {self._lib.orbb_stage_name(i).decode(): float(ms[i]) for i in range(n)}

    @property
    def launch_count(self):
        return int(self._lib.orbb_launch_count(self._h))

    @property
    def stream(self):
        """cudaStream_t of the handle (wrap with torch.cuda.ExternalStream to record events on it)"""
        return self._lib.orbb_stream(self._h)

    # ---- stage taps (parity tests) ----
    def debug_level(self, frame, level, blurred=False, bordered=False):
        w, h = C.c_int(), C.c_int()
        capi.check(self._lib.orbb_debug_level(self._h, frame, level, int(blurred), int(bordered), None, 0, C.byref(w), C.byref(h)), self._h)
        out = np.zeros((h.value, w.value), np.uint8)
        capi.check(self._lib.orbb_debug_level(self._h, frame, level, int(blurred), int(bordered), capi.ptr(out), out.size,
                                              C.byref(w), C.byref(h)), self._h)
        return out

    def debug_raw_keys(self, frame, level):
        n = C.c_int()
        capi.check(self._lib.orbb_debug_raw_keys(self._h, frame, level, None, 0, C.byref(n)), self._h)
        out = np.zeros((max(n.value, 1), 3), np.float32)
        capi.check(self._lib.orbb_debug_raw_keys(self._h, frame, level, capi.ptr(out), len(out), C.byref(n)), self._h)
        return out[:n.value]

    def debug_selected(self, frame, level):
        n = C.c_int()
        capi.check(self._lib.orbb_debug_selected(self._h, frame, level, None, 0, C.byref(n)), self._h)
        out = np.zeros((max(n.value, 1), 3), np.float32)
        capi.check(self._lib.orbb_debug_selected(self._h, frame, level, capi.ptr(out), len(out), C.byref(n)), self._h)
        return out[:n.value]


def compute_stereo_matches(ext_left, ext_right, bf, b, frame=0):
    """Frame::ComputeStereoMatches (Frame.cc:811-981) for frame `frame` of the two extractors' last batches.
    -> (uRight[nL], depth[nL], bestR[nL], sad[nL])"""
    lib = capi.load()
    cap = ext_left.max_keypoints
    ur = np.zeros(cap, np.float32)
    dp = np.zeros(cap, np.float32)
    br = np.zeros(cap, np.int32)
    sad = np.zeros(cap, np.int32)
    n = C.c_int()
    capi.check(lib.orbb_stereo_match(ext_left._h, ext_right._h, frame, bf, b, capi.ptr(ur), capi.ptr(dp), capi.ptr(br),
                                     capi.ptr(sad), cap, C.byref(n)), ext_left._h)
    return ur[:n.value], dp[:n.value], br[:n.value], sad[:n.value]


def stereo_match_batch(ext_left, ext_right, nframes, bf, b):
    capi.check(capi.load().orbb_stereo_match_batch(ext_left._h, ext_right._h, nframes, bf, b), ext_left._h)


def rgbd_stereo_batch(ext, depth_dev, nframes, width, height, K4, dist, bf, depth_is_u16=False, depth_factor=1.0, row_stride=None,
                      frame_stride=None):
    """Frame::ComputeStereoFromRGBD (Frame.cc:984-1005) for the extractor's last batch; depth maps resident on the device
    (torch CUDA tensor / address, float32 or uint16); results via stereo_fetch(ext, nframes)"""
    px = 2 if depth_is_u16 else 4
    row_stride = row_stride or width * px
    frame_stride = frame_stride or row_stride * height
    k = np.ascontiguousarray(K4, np.float32)
    d = np.ascontiguousarray(dist, np.float32).ravel()
    capi.check(capi.load().orbb_rgbd_stereo_batch(ext._h, capi.ptr(depth_dev), int(depth_is_u16), float(depth_factor), row_stride, frame_stride,
                                                  nframes, capi.ptr(k), capi.ptr(d), len(d), float(bf)), ext._h)


def stereo_fetch(ext_left, nframes):
    cap = ext_left.max_keypoints
    ur = np.zeros((nframes, cap), np.float32)
    dp = np.zeros((nframes, cap), np.float32)
    capi.check(capi.load().orbb_stereo_fetch(ext_left._h, nframes, capi.ptr(ur), capi.ptr(dp), cap), ext_left._h)
    return ur, dp

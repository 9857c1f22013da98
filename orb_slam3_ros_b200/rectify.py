"""Input-side rows of the scope table on liborbb200.so: the stereo rectification System::TrackStereo applies before
tracking (cv::remap with the maps of cv::initUndistortRectifyMap, reference orb_slam3/src/System.cc:233-240,
Settings.cc:506-509)."""
import ctypes as C

import numpy as np

from . import capi
from .capi import KP_DTYPE


class Rectifier:
    """cv::remap(src, dst, map_x, map_y, cv::INTER_LINEAR) for uint8 frames; maps are float32 [dst_h, dst_w]."""

    def __init__(self, map_x, map_y, src_shape, device=0):
        self._lib = capi.load()
        mx = np.ascontiguousarray(map_x, np.float32)
        my = np.ascontiguousarray(map_y, np.float32)
        if mx.shape != my.shape or mx.ndim != 2:
            raise ValueError("map_x / map_y must be float32 [dst_h, dst_w]")
        self.dst_shape = mx.shape
        self.src_shape = (int(src_shape[0]), int(src_shape[1]))
        self._r = C.c_void_p()
        capi.check(self._lib.orbb_rectifier_create(device, capi.ptr(mx), capi.ptr(my), mx.shape[1], mx.shape[1], mx.shape[0],
                                                   self.src_shape[1], self.src_shape[0], C.byref(self._r)))

    def close(self):
        if getattr(self, "_r", None):
            self._lib.orbb_rectifier_destroy(self._r)
            self._r = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def remap(self, image):
        image = np.ascontiguousarray(image, np.uint8)
        if image.shape != self.src_shape:
            raise ValueError(f"image must be {self.src_shape}")
        out = np.zeros(self.dst_shape, np.uint8)
        capi.check(self._lib.orbb_remap(self._r, capi.ptr(image), image.strides[0], capi.ptr(out), out.strides[0]))
        return out

    def extract(self, extractor, image, lapping=(0, 0)):
        """rectify + ORBextractor::operator() with the rectified image staying on the device"""
        image = np.ascontiguousarray(image, np.uint8)
        if image.shape != self.src_shape:
            raise ValueError(f"image must be {self.src_shape}")
        n, mono = C.c_int(0), C.c_int(0)
        for _ in range(2):                      # a second pass if the plan for this size allows more keypoints than the estimate
            cap = extractor.max_keypoints
            kps = np.zeros(cap, KP_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            rc = self._lib.orbb_extract_rectified(extractor._h, self._r, capi.ptr(image), image.strides[0], int(lapping[0]), int(lapping[1]),
                                                  capi.ptr(kps), capi.ptr(desc), cap, C.byref(n), C.byref(mono))
            if rc != capi.ORBB_ERR_CAPACITY:
                break
        capi.check(rc, extractor._h)
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch_device(self, extractor, dev_ptr, nframes, row_stride=None, frame_stride=None, lapping=(0, 0)):
        """raw frames resident in device memory -> rectified + extracted (asynchronous; results via extractor.fetch)"""
        row_stride = row_stride or self.src_shape[1]
        frame_stride = frame_stride or row_stride * self.src_shape[0]
        capi.check(self._lib.orbb_extract_batch_rectified(extractor._h, self._r, capi.ptr(dev_ptr), nframes, row_stride, frame_stride,
                                                          int(lapping[0]), int(lapping[1])), extractor._h)

#!/usr/bin/env python3
"""Where the single-frame call's time goes: orbb_extract with pageable / pinned / no output buffers (wall clock, one B200)."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from orb_slam3_ros_b200 import capi, synth                         # noqa: E402
from orb_slam3_ros_b200.extractor import KP_DTYPE, ORBextractor    # noqa: E402

ext = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=1)
lib, h = ext._lib, ext._h
img_pageable = synth.frame(480, 752, 1)
img_pinned_t = torch.from_numpy(img_pageable).pin_memory()
img_pinned = img_pinned_t.numpy()
cap = ext.max_keypoints
kp_page, d_page = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8)
kp_pin_t, d_pin_t = torch.zeros(cap * 24, dtype=torch.uint8).pin_memory(), torch.zeros(cap * 32, dtype=torch.uint8).pin_memory()
kp_pin, d_pin = kp_pin_t.numpy(), d_pin_t.numpy()
n, mono = C.c_int(0), C.c_int(0)


def run(img, kps, desc, reps=200):
    args = (h, capi.ptr(img), 752, 480, img.strides[0], 0, 1000, capi.ptr(kps) if kps is not None else None,
            capi.ptr(desc) if desc is not None else None, cap, C.byref(n), C.byref(mono))
    for _ in range(10):
        capi.check(lib.orbb_extract(*args), h)
    t0 = time.perf_counter()
    for _ in range(reps):
        lib.orbb_extract(*args)
    return (time.perf_counter() - t0) / reps * 1e3


print("python __call__ (pageable in/out, allocs): %.4f ms" % (lambda: (ext(img_pageable, None, (0, 1000)), [ext(img_pageable, None, (0, 1000)) for _ in range(10)],
      (lambda t0: ([ext(img_pageable, None, (0, 1000)) for _ in range(200)], (time.perf_counter() - t0) / 200 * 1e3)[1])(time.perf_counter()))[2])())
print("C call, pageable image, pageable outputs: %.4f ms" % run(img_pageable, kp_page, d_page))
print("C call, pinned image,   pageable outputs: %.4f ms" % run(img_pinned, kp_page, d_page))
print("C call, pinned image,   pinned outputs:   %.4f ms" % run(img_pinned, kp_pin, d_pin))
print("C call, pinned image,   no outputs:       %.4f ms" % run(img_pinned, None, None))
print("keypoints:", n.value)

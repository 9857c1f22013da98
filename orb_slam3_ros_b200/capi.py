"""ctypes binding of liborbb200.so (include/orbb200.h).  No fallback: if the CUDA library is missing this raises."""
import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["ORBB_LIB"]) if os.environ.get("ORBB_LIB") else PKG / "liborbb200.so"      # (ORBB_LIB: A/B runs against another build)

ORBB_OK = 0
ORBB_ERR_EMPTY = -1
ORBB_ERR_UNSUPPORTED = -2
ORBB_ERR_CAPACITY = -3
ORBB_ERR_ARG = -4
ORBB_ERR_CUDA = -5
ORBB_ERR_INTERNAL = -6
ORBB_MAX_LEVELS = 16

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4")])


class OrbbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liborbb200 error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("device", C.c_int32), ("max_batch", C.c_int32)]


class FrameView(C.Structure):
    """orbb_frame_view (include/orbb200.h): the frame side of the matcher scans, host- or device-resident"""
    _fields_ = [("kps_xy", C.c_void_p), ("kps_stride", C.c_size_t), ("octaves", C.c_void_p), ("oct_stride", C.c_size_t),
                ("desc", C.c_void_p), ("u_right", C.c_void_p), ("n", C.c_int32), ("on_device", C.c_int32)]


# every symbol include/orbb200.h declares: (name, restype, argtypes)
_VP, _I, _F, _SZ, _LL, _D = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong, C.c_double
_PI = C.POINTER(C.c_int)
SYMBOLS = [
    ("orbb_create", _I, [C.POINTER(Params), C.POINTER(_VP)]),
    ("orbb_destroy", None, [_VP]),
    ("orbb_last_error", C.c_char_p, [_VP]),
    ("orbb_version", C.c_char_p, []),
    ("orbb_get_tables", _I, [_VP, _VP, _VP, _VP, _VP, _VP]),
    ("orbb_max_keypoints", _I, [_VP]),
    ("orbb_extract", _I, [_VP, _VP, _I, _I, _SZ, _I, _I, _VP, _VP, _I, _PI, _PI]),
    ("orbb_pyramid_level", _I, [_VP, _I, C.POINTER(_VP), _PI, _PI, C.POINTER(_SZ)]),
    ("orbb_extract_batch", _I, [_VP, _VP, _I, _I, _I, _SZ, _SZ, _I, _I]),
    ("orbb_extract_batch_host", _I, [_VP, _VP, _I, _I, _I, _SZ, _SZ, _I, _I, _VP, _VP, _I, _VP]),
    ("orbb_extract_batch_host_submit", _I, [_VP, _VP, _I, _I, _I, _SZ, _SZ, _I, _I, _VP, _VP, _I]),
    ("orbb_extract_batch_host_wait", _I, [_VP, _VP]),
    ("orbb_extract_color", _I, [_VP, _VP, _I, _I, _SZ, _I, _I, _I, _I, _VP, _VP, _I, _PI, _PI]),
    ("orbb_extract_batch_color", _I, [_VP, _VP, _I, _I, _I, _SZ, _SZ, _I, _I, _I, _I]),
    ("orbb_sync", _I, [_VP]),
    ("orbb_stream", _VP, [_VP]),
    ("orbb_batch_fetch", _I, [_VP, _I, _VP, _VP, _I, _VP]),
    ("orbb_batch_device_ptrs", _I, [_VP, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP)]),
    ("orbb_launch_count", _LL, [_VP]),
    ("orbb_stage_times", _I, [_VP, _VP, _I]),
    ("orbb_stage_name", C.c_char_p, [_I]),
    ("orbb_set_profiling", _I, [_VP, _I]),
    ("orbb_debug_level", _I, [_VP, _I, _I, _I, _I, _VP, _SZ, _PI, _PI]),
    ("orbb_debug_raw_keys", _I, [_VP, _I, _I, _VP, _I, _PI]),
    ("orbb_debug_selected", _I, [_VP, _I, _I, _VP, _I, _PI]),
    ("orbb_stereo_match", _I, [_VP, _VP, _I, _F, _F, _VP, _VP, _VP, _VP, _I, _PI]),
    ("orbb_stereo_match_batch", _I, [_VP, _VP, _I, _F, _F]),
    ("orbb_stereo_fetch", _I, [_VP, _I, _VP, _VP, _I]),
    ("orbb_rgbd_stereo_batch", _I, [_VP, _VP, _I, _F, _SZ, _SZ, _I, _VP, _VP, _I, _F]),
    ("orbb_hamming_distance", _I, [_VP, _VP]),
    ("orbb_matcher_create", _I, [_I, C.POINTER(_VP)]),
    ("orbb_matcher_destroy", None, [_VP]),
    ("orbb_matcher_last_error", C.c_char_p, [_VP]),
    ("orbb_matcher_launch_count", _LL, [_VP]),
    ("orbb_matcher_stream", _VP, [_VP]),
    ("orbb_knn2", _I, [_VP, _VP, _I, _VP, C.c_int64, _VP, _VP]),
    ("orbb_knn2_dev", _I, [_VP, _VP, _I, _VP, C.c_int64, C.c_int32, _VP, _VP]),
    ("orbb_knn2_merge_dev", _I, [_VP, _VP, _VP, _I, _I, _VP, _VP]),
    ("orbb_knn2_sharded", _I, [_VP, _VP, _VP, _I, _VP, C.c_int64, C.c_int32, _VP, _VP]),
    ("orbb_nccl_version", _I, []),
    ("orbb_nccl_unique_id", _I, [_VP]),
    ("orbb_nccl_comm_create", _I, [_I, _I, _I, _VP, C.POINTER(_VP)]),
    ("orbb_nccl_comm_destroy", None, [_VP]),
    ("orbb_ratio_test_dev", _I, [_VP, _VP, _VP, _I, _D, _VP]),
    ("orbb_best2_csr", _I, [_VP, _VP, _I, _VP, _I, _VP, _VP, _I, _VP]),
    ("orbb_search_area_best2", _I, [_VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _I, _VP, _VP, _I, _VP]),
    ("orbb_frame_upload", _I, [_VP, _I, _VP, _VP]),
    ("orbb_search_area_topk", _I, [_VP, _VP, _VP, _VP, _VP, _VP, _I, _VP, _I, _I, _VP]),
    ("orbb_best2_csr_dev", _I, [_VP, _VP, _I, _VP, _I, _VP, _VP, _I, _VP]),
    ("orbb_distinctive_csr", _I, [_VP, _VP, _I, _VP, _I, _VP]),
    ("orbb_vocab_create", _I, [_I, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP, _I, C.POINTER(_VP)]),
    ("orbb_vocab_destroy", None, [_VP]),
    ("orbb_vocab_last_error", C.c_char_p, [_VP]),
    ("orbb_vocab_launch_count", _LL, [_VP]),
    ("orbb_bow_transform", _I, [_VP, _VP, _VP, _I, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP]),
    ("orbb_extract_batch_resized", _I, [_VP, _VP, _I, _I, _I, _SZ, _SZ, _I, _I, _I, _I]),
    ("orbb_extract_resized", _I, [_VP, _VP, _I, _I, _SZ, _I, _I, _I, _I, _VP, _VP, _I, _VP, _VP]),
    ("orbb_rectifier_create", _I, [_I, _VP, _VP, _SZ, _I, _I, _I, _I, C.POINTER(_VP)]),
    ("orbb_rectifier_destroy", None, [_VP]),
    ("orbb_remap", _I, [_VP, _VP, _SZ, _VP, _SZ]),
    ("orbb_extract_batch_rectified", _I, [_VP, _VP, _VP, _I, _SZ, _SZ, _I, _I]),
    ("orbb_extract_rectified", _I, [_VP, _VP, _VP, _SZ, _I, _I, _VP, _VP, _I, _VP, _VP]),
    ("orbb_rotation_check_csr", _I, [_VP, _VP, _VP, _I, _VP, _I, _VP, _VP]),
    ("orbb_undistort_points", _I, [_VP, _VP, _I, _VP, _VP, _I, _VP, _VP]),
    ("orbb_host_alloc", _VP, [_SZ]),
    ("orbb_host_free", None, [_VP]),
]

_lib = None


def load():
    """Load liborbb200.so; raises if it has not been built (python -m orb_slam3_ros_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise OrbbError(ORBB_ERR_CUDA, f"{LIB_PATH} is missing -- build it with __graft_entry__.build(); there is no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """data pointer of a numpy array / int address / torch tensor"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def check(rc, handle=None, matcher=False):
    if rc != ORBB_OK:
        lib = load()
        msg = (lib.orbb_matcher_last_error(handle) if matcher else lib.orbb_last_error(handle)) or b""
        raise OrbbError(rc, msg.decode(errors="replace"))

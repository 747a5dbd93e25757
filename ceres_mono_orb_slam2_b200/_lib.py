"""ctypes binding of libcmos_b200.so (the C ABI declared in include/cmos_b200.h).

There is no fallback: if the CUDA library is missing or fails to load, importing a compute entry point
raises.  Build it with `make -C ceres_mono_orb_slam2_b200/csrc` (or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CMOS_B200_LIB") or os.path.join(_HERE, "libcmos_b200.so")   # override: A/B builds while profiling

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28   # == sizeof(cv::KeyPoint) == sizeof(cmos_keypoint)

CMOS_MAX_LEVELS = 16


class CmosError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"cmos_b200 status {status}: {text}")
        self.status = status


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("max_width", C.c_int32),
                ("max_height", C.c_int32), ("max_batch", C.c_int32), ("device", C.c_int32)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not built: run `make -C {os.path.join(_HERE, 'csrc')}`; "
                              "there is no CPU fallback for the hot paths")
        _lib = C.CDLL(LIB_PATH)
        _lib.cmos_last_error.restype = C.c_char_p
        _lib.cmos_version.restype = C.c_char_p
    return _lib


def check(status: int) -> None:
    if status != 0:
        raise CmosError(status, lib().cmos_last_error().decode(errors="replace"))


def ptr(a) -> C.c_void_p:
    """numpy array, torch tensor (via data_ptr) or int -> void*"""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))

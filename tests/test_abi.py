"""CPU tests of the drop-in boundary: the shared library loads, exports every function include/cmos_b200.h
declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from ceres_mono_orb_slam2_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "cmos_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmos_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in cmos_b200.h but not exported: {missing}"
    assert b"sm_100a" in L.cmos_version()


def test_keypoint_layout_is_cv_keypoint():
    assert _lib.KP_DTYPE.itemsize == 28
    assert [_lib.KP_DTYPE.fields[n][1] for n in ("x", "y", "size", "angle", "response", "octave", "class_id")] == \
        [0, 4, 8, 12, 16, 20, 24]


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = _lib.lib()
    p = _lib.OrbParams(1000, 1.2, 8, 20, 7, 640, 480, 1, 0)
    h = C.c_void_p()
    rc = L.cmos_orb_create(C.byref(p), C.byref(h))
    assert rc == -2, "cmos_orb_create must fail with CMOS_ERR_CUDA when there is no GPU"
    assert b"no CPU fallback" in L.cmos_last_error() or b"CUDA" in L.cmos_last_error()


def test_product_package_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "ceres_mono_orb_slam2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "oracle/", "pyoracle"):
                    assert needle not in src, f"{f} references the oracle ({needle})"

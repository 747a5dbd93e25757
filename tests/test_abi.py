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


def test_last_frame_record_layouts_and_packers():
    """Host side of cmos_track_submit_points / cmos_track_submit_map: the numpy record layouts equal the C structs of
    include/cmos_b200.h (64 and 12 bytes, field offsets), and the two packers visit the same keypoints in the same order —
    every keypoint with flag bit 0 inside the frame's count, frame after frame, in increasing keypoint index."""
    import numpy as np
    from ceres_mono_orb_slam2_b200.tracking import (ASSOC_DTYPE, LAST_POINT_DTYPE, pack_last_points, pack_map_associations)
    assert LAST_POINT_DTYPE.itemsize == 64
    assert [LAST_POINT_DTYPE.fields[n][1] for n in ("descriptor", "xw", "angle", "index", "octave", "flags")] == [0, 32, 56, 60, 62, 63]
    assert ASSOC_DTYPE.itemsize == 12
    assert [ASSOC_DTYPE.fields[n][1] for n in ("slot", "angle", "index", "octave", "flags")] == [0, 4, 8, 10, 11]
    header = open(os.path.join(ROOT, "include", "cmos_b200.h")).read()
    assert "} cmos_track_assoc;        /* 12 bytes */" in header and "} cmos_last_point;         /* 64 bytes */" in header
    rng = np.random.default_rng(5)
    B, S = 4, 50
    kps = np.zeros((B, S), _lib.KP_DTYPE)
    kps["angle"] = rng.uniform(0, 360, (B, S)).astype(np.float32); kps["octave"] = rng.integers(0, 8, (B, S))
    counts = np.array([50, 0, 17, 33], np.int32)
    flags = rng.integers(0, 4, (B, S)).astype(np.uint8)
    flags[3] &= 2                                            # a frame whose keypoints carry no usable map point
    xw = rng.normal(size=(B, S, 3)); desc = rng.integers(0, 256, (B, S, 32)).astype(np.uint8)
    slots = rng.permutation(B * S).astype(np.int32).reshape(B, S)
    pts, pstart = pack_last_points(kps, counts, flags, xw, desc)
    assoc, astart = pack_map_associations(kps, counts, flags, slots)
    assert np.array_equal(pstart, astart) and pstart[0] == 0 and len(pts) == len(assoc) == pstart[-1]
    assert pstart[2] == pstart[1] and pstart[4] == pstart[3]  # empty frame, frame without usable points
    for f in range(B):
        idx = np.nonzero(flags[f, :counts[f]] & 1)[0]
        r = pts[pstart[f]:pstart[f + 1]]; a = assoc[astart[f]:astart[f + 1]]
        assert np.array_equal(r["index"], idx) and np.array_equal(a["index"], idx)
        assert np.array_equal(r["xw"], xw[f, idx]) and np.array_equal(r["descriptor"], desc[f, idx])
        assert np.array_equal(a["slot"], slots[f, idx])
        for name in ("angle", "octave", "flags"):
            assert np.array_equal(r[name], a[name])
        assert np.array_equal(a["angle"], kps["angle"][f, idx]) and np.all(a["flags"] & 1)

"""Host-side mirror of the batched MapPoint maintenance (cmos_map_* in include/cmos_b200.h):
MapPoint::ComputeDistinctiveDescriptors and MapPoint::UpdateNormalAndDepth (src/MapPoint.cc:256-315,335-378) for
many map points per call.  All compute is in libcmos_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


class MapParams(C.Structure):
    _fields_ = [("max_points", C.c_int32), ("max_observations", C.c_int32), ("max_keyframes", C.c_int32), ("device", C.c_int32)]


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


class MapPointOps:
    def __init__(self, max_points: int = 100000, max_observations: int = 1000000, max_keyframes: int = 4096, device: int = 0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        p = MapParams(max_points, max_observations, max_keyframes, device)
        check(self._L.cmos_map_create(C.byref(p), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.cmos_map_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def ComputeDistinctiveDescriptors(self, obs_start, descriptors):
        """-> (best_index [P], descriptor [P,32])"""
        st = _c(obs_start, np.int32); d = _c(descriptors, np.uint8)
        P = len(st) - 1
        best = np.full(max(P, 1), -1, np.int32); out = np.zeros((max(P, 1), 32), np.uint8)
        check(self._L.cmos_map_distinctive_descriptors(self._h, P, ptr(st), ptr(d), ptr(best), ptr(out)))
        return best[:P], out[:P]

    def UpdateNormalAndDepth(self, obs_start, obs_keyframe, camera_centers, world_pos, ref_keyframe, ref_level,
                             scale_factors, normal, min_distance, max_distance):
        st = _c(obs_start, np.int32); ok = _c(obs_keyframe, np.int32); ow = _c(camera_centers, np.float64)
        x = _c(world_pos, np.float64); rk = _c(ref_keyframe, np.int32); rl = _c(ref_level, np.int32)
        sf = _c(scale_factors, np.float32)
        nr = _c(normal, np.float64).copy(); mn = _c(min_distance, np.float32).copy(); mx = _c(max_distance, np.float32).copy()
        check(self._L.cmos_map_update_normal_and_depth(self._h, len(st) - 1, ptr(st), ptr(ok), len(ow), ptr(ow), ptr(x), ptr(rk),
                                                       ptr(rl), ptr(sf), len(sf), ptr(nr), ptr(mn), ptr(mx)))
        return nr, mn, mx

    def launch_count(self) -> int:
        n = C.c_int32()
        check(self._L.cmos_map_last_launch_count(self._h, C.byref(n)))
        return n.value

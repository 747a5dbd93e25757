"""Host-side mirror of the reference's ORBmatcher (include/ORBmatcher.h:36-102) over the C ABI, working on
flattened frame views instead of Frame*/MapPoint* graphs.  All compute is in libcmos_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import CMOS_MAX_LEVELS, KP_DTYPE, check, ptr

GRID_COLS, GRID_ROWS = 64, 48


class Camera(C.Structure):
    """cmos_camera: the Frame statics a search reads (Frame.h:158-189)."""
    _fields_ = [("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float),
                ("grid_element_width_inv", C.c_float), ("grid_element_height_inv", C.c_float),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float),
                ("nlevels", C.c_int32), ("scale_factors", C.c_float * CMOS_MAX_LEVELS),
                ("log_scale_factor", C.c_float)]

    @classmethod
    def create(cls, width, height, K, scale_factors, scale_factor=1.2):
        cam = cls()
        sf = np.ascontiguousarray(scale_factors, np.float32)
        check(_lib.lib().cmos_camera_init(C.byref(cam), width, height, C.c_float(K[0]), C.c_float(K[1]),
                                          C.c_float(K[2]), C.c_float(K[3]), ptr(sf), len(sf),
                                          C.c_float(scale_factor)))
        return cam

    @classmethod
    def create_distorted(cls, width, height, K, dist_coef, scale_factors, scale_factor=1.2, device=0):
        """Camera with lens distortion: bounds = undistorted image corners (Frame::ComputeImageBounds)."""
        cam = cls()
        sf = np.ascontiguousarray(scale_factors, np.float32)
        dc = np.ascontiguousarray(dist_coef, np.float32)
        check(_lib.lib().cmos_camera_init_distorted(C.byref(cam), width, height, C.c_float(K[0]), C.c_float(K[1]),
                                                    C.c_float(K[2]), C.c_float(K[3]), ptr(dc), len(dc), ptr(sf), len(sf),
                                                    C.c_float(scale_factor), device))
        return cam

    def bounds6(self):
        return np.array([self.min_x, self.max_x, self.min_y, self.max_y, self.grid_element_width_inv,
                         self.grid_element_height_inv], np.float32)

    def K4(self):
        return np.array([self.fx, self.fy, self.cx, self.cy], np.float32)


class MatchParams(C.Structure):
    _fields_ = [("max_batch", C.c_int32), ("max_keypoints", C.c_int32), ("max_points", C.c_int32),
                ("device", C.c_int32)]


class ORBmatcher:
    TH_HIGH = 100
    TH_LOW = 50
    HISTO_LENGTH = 30

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, max_batch: int = 1, max_keypoints: int = 2048,
                 max_points: int = 4096, device: int = 0):
        self._L = _lib.lib()
        self.nnratio = float(nnratio)
        self.check_ori = bool(checkOri)
        self._h = C.c_void_p()
        p = MatchParams(max_batch, max_keypoints, max_points, device)
        check(self._L.cmos_match_create(C.byref(p), C.byref(self._h)))
        self.max_batch, self.max_keypoints, self.max_points = max_batch, max_keypoints, max_points
        self.n_frames = 0
        self.stride = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.cmos_match_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- Frame::AssignFeaturesToGrid for the current frames ----
    def set_frames(self, cam: Camera, keypoints, descriptors, counts, n_frames: int, stride: int,
                   on_device: bool = False, stream: int = 0):
        self._keep = (keypoints, descriptors, counts)
        check(self._L.cmos_match_set_frames(self._h, C.byref(cam), ptr(keypoints), ptr(descriptors), ptr(counts),
                                            n_frames, stride, int(on_device), C.c_void_p(stream)))
        self.n_frames, self.stride = n_frames, stride

    # ---- Frame::UndistortKeyPoints (Frame.cc:329-355) for a batch ----
    def UndistortKeyPoints(self, K4, dist_coef, keypoints, counts, out=None, on_device: bool = False, stream: int = 0):
        """keypoints [B, stride] (KP_DTYPE) -> undistorted keypoints, same layout."""
        k4 = np.ascontiguousarray(K4, np.float32); dc = np.ascontiguousarray(dist_coef, np.float32)
        if on_device:
            n_frames, stride = out[1]
            dst = out[0]
        else:
            keypoints = np.ascontiguousarray(keypoints); counts = np.ascontiguousarray(counts, np.int32)
            n_frames, stride = keypoints.shape
            dst = np.zeros_like(keypoints) if out is None else out
        check(self._L.cmos_match_undistort_keypoints(self._h, ptr(k4), ptr(dc), len(dc), ptr(keypoints), ptr(counts),
                                                     n_frames, stride, ptr(dst), int(on_device), C.c_void_p(stream)))
        return dst

    def debug_grid(self, frame: int):
        gs = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32)
        gi = np.zeros(self.stride, np.int32)
        check(self._L.cmos_match_debug_grid(self._h, frame, ptr(gs), ptr(gi)))
        return gs, gi[:gs[-1]]

    # ---- SearchByProjection(CurrentFrame, LastFrame, th), ORBmatcher.h:49-53 ----
    def SearchByProjectionFrame(self, Tcw, last_keypoints, last_counts, last_flags, last_xw, last_descriptors,
                                last_stride: int, th: float, claimed=None, out=None, on_device: bool = False,
                                stream: int = 0):
        if out is None:
            match = np.full((self.n_frames, self.stride), -1, np.int32)
            nmatches = np.zeros(self.n_frames, np.int32)
        else:
            match, nmatches = out
        check(self._L.cmos_match_search_by_projection_frame(
            self._h, ptr(Tcw), ptr(last_keypoints), ptr(last_counts), ptr(last_flags), ptr(last_xw),
            ptr(last_descriptors), last_stride, C.c_float(th), int(self.check_ori), ptr(claimed), ptr(match),
            ptr(nmatches), int(on_device), C.c_void_p(stream)))
        return match, nmatches

    # ---- SearchByProjection(F, vpMapPoints, th), ORBmatcher.h:43-47 ----
    def SearchByProjectionPoints(self, n_points, in_view, level, view_cos, proj_xy, descriptors, has_obs,
                                 point_stride: int, th: float = 3.0, claimed=None, out=None,
                                 on_device: bool = False, stream: int = 0):
        if out is None:
            assign = np.full((self.n_frames, self.stride), -1, np.int32)
            nmatches = np.zeros(self.n_frames, np.int32)
        else:
            assign, nmatches = out
        check(self._L.cmos_match_search_by_projection_points(
            self._h, ptr(n_points), ptr(in_view), ptr(level), ptr(view_cos), ptr(proj_xy), ptr(descriptors),
            ptr(has_obs), point_stride, C.c_float(th), C.c_float(self.nnratio), ptr(claimed), ptr(assign),
            ptr(nmatches), int(on_device), C.c_void_p(stream)))
        return assign, nmatches

    # ---- Frame::isInFrustum for a batch of points ----
    def IsInFrustum(self, cam: Camera, pose15, view_cos_limit, n_points, xw, normal, min_distance, max_distance,
                    point_stride: int, n_frames: int):
        in_view = np.zeros((n_frames, point_stride), np.uint8)
        proj = np.zeros((n_frames, point_stride, 2), np.float32)
        level = np.zeros((n_frames, point_stride), np.int32)
        vcos = np.zeros((n_frames, point_stride), np.float32)
        check(self._L.cmos_match_is_in_frustum(self._h, C.byref(cam), ptr(pose15), C.c_float(view_cos_limit),
                                               ptr(n_points), ptr(xw), ptr(normal), ptr(min_distance),
                                               ptr(max_distance), point_stride, n_frames, ptr(in_view), ptr(proj),
                                               ptr(level), ptr(vcos), 0, C.c_void_p(0)))
        return in_view, proj, level, vcos

    def set_profiling(self, enable: bool = True):
        check(self._L.cmos_match_set_profiling(self._h, int(enable)))

    def stage_times(self):
        """-> dict kernel name -> (accumulated ms, calls)."""
        ms = (C.c_double * 4)()
        calls = (C.c_int64 * 4)()
        check(self._L.cmos_match_stage_times(self._h, ms, calls))
        names = ["grid", "search_frame", "search_points", "in_frustum"]
        return {names[i]: (ms[i], calls[i]) for i in range(4)}

    def launch_count(self) -> int:
        n = C.c_int32()
        check(self._L.cmos_match_last_launch_count(self._h, C.byref(n)))
        return n.value

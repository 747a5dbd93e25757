"""Host-side mirror of the remaining ORBmatcher searches (include/ORBmatcher.h:50-96) over the C ABI
(cmos_kfmatch_* in include/cmos_b200.h): relocalisation / loop-closing projections, SearchByBoW x2,
SearchForInitialization, SearchForTriangulation, Fuse x2, SearchBySim3.  Flattened views instead of
Frame*/KeyFrame*/MapPoint* graphs; all compute is in libcmos_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, ptr
from .orb_matcher import Camera


class KfMatchParams(C.Structure):
    _fields_ = [("max_keypoints", C.c_int32), ("max_points", C.c_int32), ("max_nodes", C.c_int32), ("device", C.c_int32)]


class FeatureVector(C.Structure):
    """cmos_feature_vector: flattened DBoW2::FeatureVector."""
    _fields_ = [("n_nodes", C.c_int32), ("node_ids", C.c_void_p), ("start", C.c_void_p), ("features", C.c_void_p)]

    @classmethod
    def create(cls, node_ids, start, features):
        fv = cls()
        fv._keep = (np.ascontiguousarray(node_ids, np.int32), np.ascontiguousarray(start, np.int32),
                    np.ascontiguousarray(np.concatenate([np.asarray(features, np.int32), [0]]), np.int32))
        fv.n_nodes = len(fv._keep[0])
        fv.node_ids = fv._keep[0].ctypes.data; fv.start = fv._keep[1].ctypes.data; fv.features = fv._keep[2].ctypes.data
        return fv


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


class KeyFrameMatcher:
    TH_HIGH = 100
    TH_LOW = 50

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, max_keypoints: int = 4096, max_points: int = 8192,
                 max_nodes: int = 4096, device: int = 0):
        self._L = _lib.lib()
        self.nnratio, self.check_ori = float(nnratio), bool(checkOri)
        self._h = C.c_void_p()
        p = KfMatchParams(max_keypoints, max_points, max_nodes, device)
        check(self._L.cmos_kfmatch_create(C.byref(p), C.byref(self._h)))
        self.n = [0, 0]

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.cmos_kfmatch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_view(self, slot: int, cam: Camera, keypoints, descriptors, is_keyframe: bool):
        k = _c(keypoints, KP_DTYPE); d = _c(descriptors, np.uint8)
        check(self._L.cmos_kfmatch_set_view(self._h, slot, C.byref(cam), int(is_keyframe), ptr(k), ptr(d), len(k)))
        self.n[slot] = len(k)

    # SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist), ORBmatcher.h:52-54
    def SearchByProjectionKeyFrame(self, Tcw, kf_valid, kf_xw, kf_min_d, kf_max_d, kf_desc, kf_angle, th, ORBdist,
                                   cur_has_point):
        n = self.n[0]
        hp = _c(cur_has_point, np.uint8).copy(); match = np.full(max(n, 1), -1, np.int32); nm = C.c_int32()
        a = (_c(Tcw, np.float64), _c(kf_valid, np.uint8), _c(kf_xw, np.float64), _c(kf_min_d, np.float32),
             _c(kf_max_d, np.float32), _c(kf_desc, np.uint8), _c(kf_angle, np.float32))
        check(self._L.cmos_kfmatch_search_by_projection_reloc(self._h, ptr(a[0]), len(a[1]), ptr(a[1]), ptr(a[2]), ptr(a[3]),
                                                              ptr(a[4]), ptr(a[5]), ptr(a[6]), C.c_float(th), int(ORBdist),
                                                              int(self.check_ori), ptr(hp), ptr(match), C.byref(nm)))
        return match[:n], nm.value, hp

    # SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th), ORBmatcher.h:58-60
    def SearchByProjectionSim3(self, Scw, pt_skip, xw, normal, min_d, max_d, pt_desc, th, matched):
        n = self.n[0]
        m = _c(matched, np.uint8).copy(); assign = np.full(max(n, 1), -1, np.int32); nm = C.c_int32()
        a = (_c(Scw, np.float64), _c(pt_skip, np.uint8), _c(xw, np.float64), _c(normal, np.float64), _c(min_d, np.float32),
             _c(max_d, np.float32), _c(pt_desc, np.uint8))
        check(self._L.cmos_kfmatch_search_by_projection_sim3(self._h, ptr(a[0]), len(a[1]), ptr(a[1]), ptr(a[2]), ptr(a[3]),
                                                             ptr(a[4]), ptr(a[5]), ptr(a[6]), int(th), ptr(m), ptr(assign),
                                                             C.byref(nm)))
        return assign[:n], nm.value, m

    # Fuse(KeyFrame*, vpMapPoints, th) / Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint), ORBmatcher.h:90-96
    def Fuse(self, pose, pt_skip, xw, normal, min_d, max_d, pt_desc, th, inv_level_sigma2=None, sim3=False):
        n = len(pt_skip)
        bi = np.full(max(n, 1), -1, np.int32); bd = np.full(max(n, 1), 256, np.int32); nf = C.c_int32()
        a = (_c(pose, np.float64), _c(pt_skip, np.uint8), _c(xw, np.float64), _c(normal, np.float64), _c(min_d, np.float32),
             _c(max_d, np.float32), _c(pt_desc, np.uint8))
        iv = None if inv_level_sigma2 is None else _c(inv_level_sigma2, np.float32)
        check(self._L.cmos_kfmatch_fuse(self._h, int(sim3), ptr(a[0]), ptr(iv), n, ptr(a[1]), ptr(a[2]), ptr(a[3]), ptr(a[4]),
                                        ptr(a[5]), ptr(a[6]), C.c_float(th), ptr(bi), ptr(bd), C.byref(nf)))
        return bi[:n], bd[:n], nf.value

    # SearchBySim3, ORBmatcher.h:84-87
    def SearchBySim3(self, pose1, pose2, s12, R12, t12, side1, side2, th):
        n1 = self.n[0]
        m12 = np.full(max(n1, 1), -1, np.int32); nf = C.c_int32()
        a = []
        for s in (side1, side2):
            a += [_c(s[0], np.uint8), _c(s[1], np.uint8), _c(s[2], np.float64), _c(s[3], np.float32), _c(s[4], np.float32),
                  _c(s[5], np.uint8)]
        p = (_c(pose1, np.float64), _c(pose2, np.float64), _c(R12, np.float64), _c(t12, np.float64))
        check(self._L.cmos_kfmatch_search_by_sim3(self._h, ptr(p[0]), ptr(p[1]), C.c_float(s12), ptr(p[2]), ptr(p[3]),
                                                  *[ptr(x) for x in a], C.c_float(th), ptr(m12), C.byref(nf)))
        return m12[:n1], nf.value

    # SearchByBoW(KeyFrame*, Frame&, ...) (mode 0) / SearchByBoW(KeyFrame*, KeyFrame*, ...) (mode 1), ORBmatcher.h:65-68
    def SearchByBoW(self, mode, valid1, fv1: FeatureVector, valid2, fv2: FeatureVector):
        n_out = self.n[1] if mode == 0 else self.n[0]
        match = np.full(max(n_out, 1), -1, np.int32); nm = C.c_int32()
        v1 = _c(valid1, np.uint8); v2 = None if valid2 is None else _c(valid2, np.uint8)
        check(self._L.cmos_kfmatch_search_by_bow(self._h, int(mode), ptr(v1), C.byref(fv1), ptr(v2), C.byref(fv2),
                                                 C.c_float(self.nnratio), int(self.check_ori), ptr(match), C.byref(nm)))
        return match[:n_out], nm.value

    # SearchForTriangulation, ORBmatcher.h:77-80
    def SearchForTriangulation(self, has1, fv1: FeatureVector, has2, fv2: FeatureVector, F12, Cw, R2w, t2w, level_sigma2_2):
        n1 = self.n[0]
        m12 = np.full(max(n1, 1), -1, np.int32); nm = C.c_int32()
        a = (_c(has1, np.uint8), _c(has2, np.uint8), _c(F12, np.float64), _c(Cw, np.float64), _c(R2w, np.float64),
             _c(t2w, np.float64), _c(level_sigma2_2, np.float32))
        check(self._L.cmos_kfmatch_search_for_triangulation(self._h, ptr(a[0]), C.byref(fv1), ptr(a[1]), C.byref(fv2),
                                                            ptr(a[2]), ptr(a[3]), ptr(a[4]), ptr(a[5]), ptr(a[6]),
                                                            int(self.check_ori), ptr(m12), C.byref(nm)))
        return m12[:n1], nm.value

    # SearchForInitialization, ORBmatcher.h:71-74
    def SearchForInitialization(self, prev_matched, windowSize: int = 10):
        n1 = self.n[0]
        prev = _c(prev_matched, np.float32).copy()
        m12 = np.full(max(n1, 1), -1, np.int32); nm = C.c_int32()
        check(self._L.cmos_kfmatch_search_for_initialization(self._h, ptr(prev), int(windowSize), C.c_float(self.nnratio),
                                                             int(self.check_ori), ptr(m12), C.byref(nm)))
        return m12[:n1], nm.value, prev

    def launch_count(self) -> int:
        n = C.c_int32()
        check(self._L.cmos_kfmatch_last_launch_count(self._h, C.byref(n)))
        return n.value

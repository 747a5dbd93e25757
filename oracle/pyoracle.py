"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/orb_oracle.cpp header).

ctypes front for oracle/liboracle.so plus a second, independent ORB oracle that calls cv2 (OpenCV 4.13,
Python wheel) at exactly the entry points where the reference calls the C++ API
(cv::resize / copyMakeBorder / FAST / GaussianBlur / fastAtan2, ORBextractor.cc:809,814,1086,1120-1127,103).
The cv2 oracle pins the C++ restatement; neither is ever imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".cpp")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orb_oracle_create.restype = C.c_void_p
        _LIB.orb_oracle_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _LIB.orb_oracle_destroy.argtypes = [C.c_void_p]
        _LIB.ba_oracle_solve.restype = C.c_int
        _LIB.ba_oracle_pose_optimization.restype = C.c_int
        for name in ("orb_oracle_tables", "orb_oracle_extract", "orb_oracle_result", "orb_oracle_level_size",
                     "orb_oracle_level_image", "orb_oracle_level_blurred", "orb_oracle_level_candidates",
                     "orb_oracle_level_keypoints"):
            getattr(_LIB, name).argtypes = None
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OrbOracle:
    """C++ restatement of ORBextractor (oracle/orb_oracle.cpp)."""

    def __init__(self, nfeatures=2000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nlevels = nlevels
        self.h = C.c_void_p(self.L.orb_oracle_create(nfeatures, C.c_float(scale), nlevels, ini_th, min_th))
        sf, isf, s2, is2 = (np.zeros(nlevels, np.float32) for _ in range(4))
        quota = np.zeros(nlevels, np.int32); umax = np.zeros(16, np.int32)
        self.L.orb_oracle_tables(self.h, _p(sf), _p(isf), _p(s2), _p(is2), _p(quota), _p(umax))
        self.scale_factors, self.inv_scale_factors, self.sigma2, self.inv_sigma2 = sf, isf, s2, is2
        self.quota, self.umax = quota, umax

    def __del__(self):
        try:
            self.L.orb_oracle_destroy(self.h)
        except Exception:
            pass

    def extract(self, img: np.ndarray):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        n = self.L.orb_oracle_extract(self.h, _p(img), w, h, w)
        kps = np.zeros(n, KP_DTYPE); desc = np.zeros((n, 32), np.uint8)
        self.L.orb_oracle_result(self.h, _p(kps), _p(desc))
        return kps, desc

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.orb_oracle_level_size(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_image(self, l):
        w, h = self.level_size(l)
        out = np.zeros((h + 38, w + 38), np.uint8)
        self.L.orb_oracle_level_image(self.h, l, _p(out))
        return out

    def level_blurred(self, l):
        w, h = self.level_size(l)
        out = np.zeros((h, w), np.uint8)
        self.L.orb_oracle_level_blurred(self.h, l, _p(out))
        return out

    def level_candidates(self, l):
        n = self.L.orb_oracle_level_candidates(self.h, l, None)
        out = np.zeros(n, KP_DTYPE)
        self.L.orb_oracle_level_candidates(self.h, l, _p(out))
        return out

    def level_keypoints(self, l):
        n = self.L.orb_oracle_level_keypoints(self.h, l, None)
        out = np.zeros(n, KP_DTYPE)
        self.L.orb_oracle_level_keypoints(self.h, l, _p(out))
        return out


def resize(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().orb_oracle_resize(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw)
    return dst


def blur7(src: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().orb_oracle_blur7(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), src.shape[1])
    return dst


def fast(img: np.ndarray, threshold: int) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size // 4 + 16
    out = np.zeros((cap, 3), np.int32)
    n = lib().orb_oracle_fast(_p(img), img.shape[1], img.shape[0], img.shape[1], threshold, _p(out), cap)
    return out[:n]


def fast_atan2(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    y = np.ascontiguousarray(y, np.float32); x = np.ascontiguousarray(x, np.float32)
    out = np.zeros_like(y)
    lib().orb_oracle_fast_atan2(_p(y), _p(x), _p(out), y.size)
    return out


def sincos_deg(deg: np.ndarray):
    deg = np.ascontiguousarray(deg, np.float32)
    c = np.zeros_like(deg); s = np.zeros_like(deg)
    lib().orb_oracle_sincos(_p(deg), _p(c), _p(s), deg.size)
    return c, s


def descriptor_distance(a: np.ndarray, b: np.ndarray) -> int:
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return int(lib().orb_oracle_descriptor_distance(_p(a), _p(b)))


# ---------------------------------------------------------------------------------------------------
# cv2-based pipeline front half: pyramid + per-cell FAST candidates + blurred levels, built from the
# same cv2 calls the reference makes.  Used to pin the C++ restatement stage by stage.

def cv2_pyramid(img: np.ndarray, inv_scale_factors: np.ndarray):
    import cv2
    levels = []
    h, w = img.shape
    for l, s in enumerate(inv_scale_factors):
        sw = int(np.rint(np.float32(w) * np.float32(s))); sh = int(np.rint(np.float32(h) * np.float32(s)))
        if l == 0:
            cur = img
        else:
            prev = levels[-1][19:-19, 19:-19]
            cur = cv2.resize(prev, (sw, sh), interpolation=cv2.INTER_LINEAR)
        levels.append(cv2.copyMakeBorder(cur, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))
    return levels


def cv2_level_candidates(bordered: np.ndarray, ini_th=20, min_th=7):
    """ORBextractor.cc:765-829 with cv2.FastFeatureDetector per cell; returns (x, y, response) rows in
    vToDistributeKeys order, coordinates relative to minBorder."""
    import cv2
    im = bordered[19:-19, 19:-19]
    h, w = im.shape
    min_b = 16; max_bx = w - 16; max_by = h - 16
    width = np.float32(max_bx - min_b); height = np.float32(max_by - min_b)
    n_cols = int(width / np.float32(30)); n_rows = int(height / np.float32(30))
    w_cell = int(np.ceil(width / n_cols)); h_cell = int(np.ceil(height / n_rows))
    det_ini = cv2.FastFeatureDetector_create(ini_th, True)
    det_min = cv2.FastFeatureDetector_create(min_th, True)
    out = []
    for i in range(n_rows):
        ini_y = min_b + i * h_cell; max_y = ini_y + h_cell + 6
        if ini_y >= max_by - 3:
            continue
        max_y = min(max_y, max_by)
        for j in range(n_cols):
            ini_x = min_b + j * w_cell; max_x = ini_x + w_cell + 6
            if ini_x >= max_bx - 6:
                continue
            max_x = min(max_x, max_bx)
            cell = np.ascontiguousarray(im[ini_y:max_y, ini_x:max_x])
            kps = det_ini.detect(cell)
            if not kps:
                kps = det_min.detect(cell)
            for kp in kps:
                out.append((kp.pt[0] + j * w_cell, kp.pt[1] + i * h_cell, kp.response))
    return np.array(out, np.float32).reshape(-1, 3)


# ---------------------------------------------------------------------------------------------------
# Matcher oracle (oracle/matcher_oracle.cpp)

def build_grid(kps: np.ndarray, bounds6: np.ndarray):
    kps = np.ascontiguousarray(kps)
    gs = np.zeros(64 * 48 + 1, np.int32); gi = np.zeros(max(len(kps), 1), np.int32)
    n = lib().match_oracle_build_grid(_p(kps), len(kps), C.c_float(bounds6[0]), C.c_float(bounds6[2]),
                                      C.c_float(bounds6[4]), C.c_float(bounds6[5]), _p(gs), _p(gi))
    return gs, gi[:n]


def features_in_area(kps, gs, gi, bounds6, x, y, r, min_level, max_level):
    kps = np.ascontiguousarray(kps); out = np.zeros(max(len(kps), 1), np.int32)
    gi2 = np.ascontiguousarray(np.concatenate([gi, np.zeros(1, np.int32)]))
    b = np.ascontiguousarray(bounds6, np.float32)
    n = lib().match_oracle_features_in_area(_p(kps), len(kps), _p(gs), _p(gi2), _p(b), C.c_float(x), C.c_float(y),
                                            C.c_float(r), int(min_level), int(max_level), _p(out))
    return out[:n]


def search_by_projection_frame(cur_kps, cur_desc, gs, gi, bounds6, K4, scale_factors, Tcw, last_kps, last_flags,
                               last_xw, last_desc, th, check_ori, claimed=None):
    cur_kps = np.ascontiguousarray(cur_kps); cur_desc = np.ascontiguousarray(cur_desc)
    last_kps = np.ascontiguousarray(last_kps)
    n = len(cur_kps)
    gi2 = np.ascontiguousarray(np.concatenate([gi, np.zeros(1, np.int32)]))
    claimed = np.zeros(max(n, 1), np.uint8) if claimed is None else claimed
    match = np.full(max(n, 1), -1, np.int32)
    b = np.ascontiguousarray(bounds6, np.float32); k = np.ascontiguousarray(K4, np.float32)
    sf = np.ascontiguousarray(scale_factors, np.float32); T = np.ascontiguousarray(Tcw, np.float64)
    lf = np.ascontiguousarray(last_flags, np.uint8); lx = np.ascontiguousarray(last_xw, np.float64)
    ld = np.ascontiguousarray(last_desc, np.uint8)
    fn = lib().match_oracle_search_by_projection_frame
    nm = fn(_p(cur_kps), _p(cur_desc), n, _p(gs), _p(gi2), _p(b), _p(k), _p(sf), _p(T), _p(last_kps), len(last_kps),
            _p(lf), _p(lx), _p(ld), C.c_float(th), int(check_ori), _p(claimed), _p(match))
    return match[:n], nm, claimed[:n]


def search_by_projection_points(kps, desc, gs, gi, bounds6, scale_factors, in_view, level, view_cos, proj_xy,
                                mp_desc, has_obs, th, nn_ratio, claimed=None):
    kps = np.ascontiguousarray(kps); desc = np.ascontiguousarray(desc)
    n = len(kps)
    gi2 = np.ascontiguousarray(np.concatenate([gi, np.zeros(1, np.int32)]))
    claimed = np.zeros(max(n, 1), np.uint8) if claimed is None else claimed
    assign = np.full(max(n, 1), -1, np.int32)
    b = np.ascontiguousarray(bounds6, np.float32); sf = np.ascontiguousarray(scale_factors, np.float32)
    iv = np.ascontiguousarray(in_view, np.uint8); lv = np.ascontiguousarray(level, np.int32)
    vc = np.ascontiguousarray(view_cos, np.float32); pj = np.ascontiguousarray(proj_xy, np.float32)
    md = np.ascontiguousarray(mp_desc, np.uint8); ho = np.ascontiguousarray(has_obs, np.uint8)
    fn = lib().match_oracle_search_by_projection_points
    nm = fn(_p(kps), _p(desc), n, _p(gs), _p(gi2), _p(b), _p(sf), len(iv), _p(iv), _p(lv), _p(vc), _p(pj), _p(md),
            _p(ho), C.c_float(th), C.c_float(nn_ratio), _p(claimed), _p(assign))
    return assign[:n], nm, claimed[:n]


def is_in_frustum(pose15, K4, bounds4, log_scale_factor, n_levels, cos_limit, xw, normal, min_dist, max_dist):
    n = len(xw)
    pose15 = np.ascontiguousarray(pose15, np.float64); k = np.ascontiguousarray(K4, np.float32)
    b = np.ascontiguousarray(bounds4, np.float32)
    xw = np.ascontiguousarray(xw, np.float64); normal = np.ascontiguousarray(normal, np.float64)
    mn = np.ascontiguousarray(min_dist, np.float32); mx = np.ascontiguousarray(max_dist, np.float32)
    in_view = np.zeros(n, np.uint8); proj = np.zeros((n, 2), np.float32)
    level = np.zeros(n, np.int32); vcos = np.zeros(n, np.float32)
    lib().match_oracle_is_in_frustum(_p(pose15), _p(k), _p(b), C.c_float(log_scale_factor), int(n_levels),
                                     C.c_float(cos_limit), n, _p(xw), _p(normal), _p(mn), _p(mx), _p(in_view),
                                     _p(proj), _p(level), _p(vcos))
    return in_view, proj, level, vcos


def undistort_points(K4, dist, xy):
    """Frame::UndistortKeyPoints' cv::undistortPoints(mat, mat, K, dist, Mat(), K) restated (oracle/matcher_oracle.cpp)."""
    k = np.ascontiguousarray(K4, np.float32); dc = np.ascontiguousarray(dist, np.float32)
    a = np.ascontiguousarray(xy, np.float32); out = np.zeros_like(a)
    lib().frame_oracle_undistort_points(_p(k), _p(dc), len(dc), _p(a), len(a), _p(out))
    return out


def image_bounds(K4, dist, width, height):
    """Frame::ComputeImageBounds (Frame.cc:357-385) -> min_x, max_x, min_y, max_y (float32)."""
    if np.float32(dist[0]) == 0.0:
        return np.array([0.0, width, 0.0, height], np.float32)
    c = undistort_points(K4, dist, np.array([[0, 0], [width, 0], [0, height], [width, height]], np.float32))
    return np.array([min(c[0, 0], c[2, 0]), max(c[1, 0], c[3, 0]), min(c[0, 1], c[1, 1]), max(c[2, 1], c[3, 1])], np.float32)


# ---------------------------------------------------------------------------------------------------
# Bundle-adjustment oracle (oracle/ba_oracle.cpp)

class BaSummary(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("successful_steps", C.c_int32), ("termination", C.c_int32),
                ("jacobian_evaluations", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


TRACE_COLS = ("cost", "cost_change", "gradient_max_norm", "step_norm", "relative_decrease", "radius", "accepted",
              "valid")


def ba_residual(cam7, X, K4, u, v, inv_sigma2):
    cam7 = np.ascontiguousarray(cam7, np.float64); X = np.ascontiguousarray(X, np.float64)
    K4 = np.ascontiguousarray(K4, np.float64)
    r = np.zeros(2); Jc = np.zeros((2, 6)); Jp = np.zeros((2, 3))
    lib().ba_oracle_residual(_p(cam7), _p(X), _p(K4), C.c_float(u), C.c_float(v), C.c_float(inv_sigma2), _p(r), _p(Jc),
                             _p(Jp))
    return r, Jc, Jp


def quat_plus(q, d):
    q = np.ascontiguousarray(q, np.float64); d = np.ascontiguousarray(d, np.float64); out = np.zeros(4)
    lib().ba_oracle_quat_plus(_p(q), _p(d), _p(out))
    return out


def ba_solve(cams, cam_const, pts, pts_const, obs_cam, obs_pt, uv, inv_sigma2, mode, K4, max_iterations):
    """Restated ceres::Solve on a flattened graph.  Returns cams, pts, summary dict, trace [it+1, 8]."""
    cams = np.array(cams, np.float64, copy=True, order="C"); pts = np.array(pts, np.float64, copy=True, order="C")
    cc = np.ascontiguousarray(cam_const, np.uint8)
    oc = np.ascontiguousarray(obs_cam, np.int32); op = np.ascontiguousarray(obs_pt, np.int32)
    uv = np.ascontiguousarray(uv, np.float32); w = np.ascontiguousarray(inv_sigma2, np.float32)
    md = None if mode is None else np.ascontiguousarray(mode, np.uint8)
    K4 = np.ascontiguousarray(K4, np.float64)
    s = BaSummary(); cap = max_iterations + 2
    trace = np.zeros((cap, 8))
    n = lib().ba_oracle_solve(len(cams), _p(cams), _p(cc), len(pts), _p(pts), int(pts_const), len(oc), _p(oc), _p(op),
                              _p(uv), _p(w), None if md is None else _p(md), _p(K4), int(max_iterations), C.byref(s),
                              _p(trace), cap)
    return cams, pts, s.as_dict(), trace[:n]


def ba_pose_optimization(pose7, xw, uv, inv_sigma2, K4, max_iterations=100):
    pose = np.array(pose7, np.float64, copy=True); xw = np.ascontiguousarray(xw, np.float64)
    uv = np.ascontiguousarray(uv, np.float32); w = np.ascontiguousarray(inv_sigma2, np.float32)
    K4 = np.ascontiguousarray(K4, np.float64)
    out = np.zeros(len(xw), np.uint8); s = BaSummary(); cap = max_iterations + 2
    trace = np.zeros((cap, 8))
    n_in = lib().ba_oracle_pose_optimization(_p(pose), len(xw), _p(xw), _p(uv), _p(w), _p(K4), int(max_iterations),
                                             _p(out), C.byref(s), _p(trace), cap)
    return pose, out, n_in, s.as_dict(), trace[:s.iterations + 1]


def ba_local(cams, cam_flags, pts, obs_cam, obs_pt, uv, inv_sigma2, K4, iters=(5, 10)):
    cams = np.array(cams, np.float64, copy=True, order="C"); pts = np.array(pts, np.float64, copy=True, order="C")
    cf = np.ascontiguousarray(cam_flags, np.uint8)
    oc = np.ascontiguousarray(obs_cam, np.int32); op = np.ascontiguousarray(obs_pt, np.int32)
    uv = np.ascontiguousarray(uv, np.float32); w = np.ascontiguousarray(inv_sigma2, np.float32)
    K4 = np.ascontiguousarray(K4, np.float64)
    erase = np.zeros(len(oc), np.uint8); s = (BaSummary * 2)()
    lib().ba_oracle_local(len(cams), _p(cams), _p(cf), len(pts), _p(pts), len(oc), _p(oc), _p(op), _p(uv), _p(w),
                          _p(K4), int(iters[0]), int(iters[1]), _p(erase), s)
    return cams, pts, erase, [s[0].as_dict(), s[1].as_dict()]


def ba_global(cams, cam_const, pts, obs_cam, obs_pt, uv, inv_sigma2, K4, n_iterations, robust=True):
    cams = np.array(cams, np.float64, copy=True, order="C"); pts = np.array(pts, np.float64, copy=True, order="C")
    cc = np.ascontiguousarray(cam_const, np.uint8)
    oc = np.ascontiguousarray(obs_cam, np.int32); op = np.ascontiguousarray(obs_pt, np.int32)
    uv = np.ascontiguousarray(uv, np.float32); w = np.ascontiguousarray(inv_sigma2, np.float32)
    K4 = np.ascontiguousarray(K4, np.float64)
    s = BaSummary(); cap = n_iterations + 2; trace = np.zeros((cap, 8))
    lib().ba_oracle_global(len(cams), _p(cams), _p(cc), len(pts), _p(pts), len(oc), _p(oc), _p(op), _p(uv), _p(w),
                           _p(K4), int(n_iterations), int(robust), C.byref(s), _p(trace), cap)
    return cams, pts, s.as_dict(), trace[:s.iterations + 1]


# ---------------------------------------------------------------------------------------------------
# The remaining ORBmatcher searches (oracle/matcher2_oracle.cpp)

class View:
    """A Frame or KeyFrame as the searches read it: keypoints, descriptors, the Frame's grid, and the 12 floats
    {min_x, max_x, min_y, max_y, grid inv w, grid inv h, fx, fy, cx, cy, log scale factor, n levels}.  A KeyFrame's
    bounds are truncated to int for the look-ups (KeyFrame.h:179-182); the grid is always the Frame's."""

    def __init__(self, kps, desc, bounds6, K4, scale_factors, scale_factor=1.2, is_keyframe=False):
        self.kps = np.ascontiguousarray(kps); self.desc = np.ascontiguousarray(desc, np.uint8)
        self.n = len(self.kps)
        b = np.asarray(bounds6, np.float32).copy()
        self.gs, gi = build_grid(self.kps, b)
        self.gi = np.ascontiguousarray(np.concatenate([gi, np.zeros(1, np.int32)]))
        if is_keyframe:
            b[:4] = np.trunc(b[:4])
        self.sf = np.ascontiguousarray(scale_factors, np.float32)
        log_sf = np.float32(np.log(np.float64(np.float32(scale_factor))))
        self.v12 = np.concatenate([b, np.asarray(K4, np.float32), [log_sf, np.float32(len(self.sf))]]).astype(np.float32)

    def args(self):
        return (_p(self.kps), _p(self.desc), self.n, _p(self.gs), _p(self.gi), _p(self.v12))


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def search_by_projection_reloc(V: View, Tcw, kf_valid, kf_xw, kf_min_d, kf_max_d, kf_desc, kf_angle, th, orb_dist,
                               check_ori, cur_has_point):
    hp = _c(cur_has_point, np.uint8).copy(); match = np.full(max(V.n, 1), -1, np.int32)
    T = _c(Tcw, np.float64); va = _c(kf_valid, np.uint8); xw = _c(kf_xw, np.float64)
    mn = _c(kf_min_d, np.float32); mx = _c(kf_max_d, np.float32); kd = _c(kf_desc, np.uint8); an = _c(kf_angle, np.float32)
    nm = lib().match2_oracle_search_by_projection_reloc(*V.args(), _p(V.sf), _p(T), len(va), _p(va), _p(xw), _p(mn), _p(mx),
                                                        _p(kd), _p(an), C.c_float(th), int(orb_dist), int(check_ori),
                                                        _p(hp), _p(match))
    return match[:V.n], nm, hp


def search_by_projection_sim3(V: View, Scw, pt_skip, xw, normal, min_d, max_d, pt_desc, th, matched):
    m = _c(matched, np.uint8).copy(); assign = np.full(max(V.n, 1), -1, np.int32)
    S = _c(Scw, np.float64); sk = _c(pt_skip, np.uint8); x = _c(xw, np.float64); nr = _c(normal, np.float64)
    mn = _c(min_d, np.float32); mx = _c(max_d, np.float32); pd = _c(pt_desc, np.uint8)
    nm = lib().match2_oracle_search_by_projection_sim3(*V.args(), _p(V.sf), _p(S), len(sk), _p(sk), _p(x), _p(nr), _p(mn),
                                                       _p(mx), _p(pd), int(th), _p(m), _p(assign))
    return assign[:V.n], nm, m


def fuse(V: View, inv_sigma2, sim3, pose, pt_skip, xw, normal, min_d, max_d, pt_desc, th):
    n = len(pt_skip)
    bi = np.full(max(n, 1), -1, np.int32); bd = np.full(max(n, 1), 256, np.int32)
    P = _c(pose, np.float64); sk = _c(pt_skip, np.uint8); x = _c(xw, np.float64); nr = _c(normal, np.float64)
    mn = _c(min_d, np.float32); mx = _c(max_d, np.float32); pd = _c(pt_desc, np.uint8); iv = _c(inv_sigma2, np.float32)
    nf = lib().match2_oracle_fuse(*V.args(), _p(V.sf), _p(iv), int(sim3), _p(P), n, _p(sk), _p(x), _p(nr), _p(mn), _p(mx),
                                  _p(pd), C.c_float(th), _p(bi), _p(bd))
    return bi[:n], bd[:n], nf


def search_by_sim3(V1: View, V2: View, pose1, pose2, s12, R12, t12, side1, side2, th):
    """side = (valid, already, xw, min_d, max_d, mp_desc) per keypoint of that keyframe."""
    m12 = np.full(max(V1.n, 1), -1, np.int32)
    a = []
    for s in (side1, side2):
        a += [_c(s[0], np.uint8), _c(s[1], np.uint8), _c(s[2], np.float64), _c(s[3], np.float32), _c(s[4], np.float32),
              _c(s[5], np.uint8)]
    p1 = _c(pose1, np.float64); p2 = _c(pose2, np.float64); R = _c(R12, np.float64); t = _c(t12, np.float64)
    nf = lib().match2_oracle_search_by_sim3(*V1.args(), *V2.args(), _p(V1.sf), _p(p1), _p(p2), C.c_float(s12), _p(R), _p(t),
                                            *[_p(x) for x in a], C.c_float(th), _p(m12))
    return m12[:V1.n], nf


def flatten_feature_vector(node_of_feature, order=None):
    """node id per feature (-1 = not in the vector) -> (node_ids ascending, start, features in insertion order)."""
    node_of_feature = np.asarray(node_of_feature)
    idx = np.arange(len(node_of_feature)) if order is None else np.asarray(order)
    idx = idx[node_of_feature[idx] >= 0]
    nodes = np.unique(node_of_feature[idx])
    start = [0]; feats = []
    for nd in nodes:
        f = idx[node_of_feature[idx] == nd]
        feats.extend(f.tolist()); start.append(len(feats))
    return nodes.astype(np.int32), np.array(start, np.int32), np.array(feats, np.int32)


def _fv(fv):
    nodes, start, feats = (_c(fv[0], np.int32), _c(fv[1], np.int32), _c(np.concatenate([fv[2], [0]]), np.int32))
    return nodes, start, feats


def search_by_bow(mode, desc1, angle1, valid1, fv1, desc2, angle2, valid2, fv2, nn_ratio, check_ori):
    d1 = _c(desc1, np.uint8); d2 = _c(desc2, np.uint8); a1 = _c(angle1, np.float32); a2 = _c(angle2, np.float32)
    v1 = _c(valid1, np.uint8); v2 = _c(valid2 if valid2 is not None else np.ones(len(d2)), np.uint8)
    n1, s1, f1 = _fv(fv1); n2, s2, f2 = _fv(fv2)
    n_out = len(d2) if mode == 0 else len(d1)
    match = np.full(max(n_out, 1), -1, np.int32)
    nm = lib().match2_oracle_search_by_bow(int(mode), _p(d1), _p(a1), len(d1), _p(v1), len(n1), _p(n1), _p(s1), _p(f1),
                                           _p(d2), _p(a2), len(d2), _p(v2), len(n2), _p(n2), _p(s2), _p(f2),
                                           C.c_float(nn_ratio), int(check_ori), _p(match))
    return match[:n_out], nm


def search_for_triangulation(kps1, desc1, has1, fv1, kps2, desc2, has2, fv2, F12, Cw, R2w, t2w, K4_2, sf2, sigma2_2,
                             check_ori):
    k1 = _c(kps1, KP_DTYPE); k2 = _c(kps2, KP_DTYPE); d1 = _c(desc1, np.uint8); d2 = _c(desc2, np.uint8)
    h1 = _c(has1, np.uint8); h2 = _c(has2, np.uint8)
    n1, s1, f1 = _fv(fv1); n2, s2, f2 = _fv(fv2)
    F = _c(F12, np.float64); c = _c(Cw, np.float64); R = _c(R2w, np.float64); t = _c(t2w, np.float64)
    K = _c(K4_2, np.float32); sf = _c(sf2, np.float32); sg = _c(sigma2_2, np.float32)
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    nm = lib().match2_oracle_search_for_triangulation(_p(k1), _p(d1), len(k1), _p(h1), len(n1), _p(n1), _p(s1), _p(f1),
                                                      _p(k2), _p(d2), len(k2), _p(h2), len(n2), _p(n2), _p(s2), _p(f2),
                                                      _p(F), _p(c), _p(R), _p(t), _p(K), _p(sf), _p(sg), int(check_ori),
                                                      _p(m12))
    return m12[:len(k1)], nm


def search_for_initialization(kps1, desc1, V2: View, prev_matched, window_size, nn_ratio, check_ori):
    k1 = _c(kps1, KP_DTYPE); d1 = _c(desc1, np.uint8)
    prev = _c(prev_matched, np.float32).copy()
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    nm = lib().match2_oracle_search_for_initialization(_p(k1), _p(d1), len(k1), *V2.args(), _p(prev), int(window_size),
                                                       C.c_float(nn_ratio), int(check_ori), _p(m12))
    return m12[:len(k1)], nm, prev


def distinctive_descriptors(obs_start, desc):
    st = _c(obs_start, np.int32); d = _c(desc, np.uint8)
    best = np.full(max(len(st) - 1, 1), -1, np.int32)
    lib().map_oracle_distinctive_descriptors(len(st) - 1, _p(st), _p(d), _p(best))
    return best[:len(st) - 1]


def update_normal_and_depth(obs_start, obs_kf, Ow, pos, ref_kf, ref_level, sf, normal, min_d, max_d):
    st = _c(obs_start, np.int32); ok = _c(obs_kf, np.int32); ow = _c(Ow, np.float64); x = _c(pos, np.float64)
    rk = _c(ref_kf, np.int32); rl = _c(ref_level, np.int32); s = _c(sf, np.float32)
    nr = _c(normal, np.float64).copy(); mn = _c(min_d, np.float32).copy(); mx = _c(max_d, np.float32).copy()
    lib().map_oracle_update_normal_and_depth(len(st) - 1, _p(st), _p(ok), _p(ow), _p(x), _p(rk), _p(rl), _p(s), len(s),
                                             _p(nr), _p(mn), _p(mx))
    return nr, mn, mx


# ---------------------------------------------------------------------------------------------------
# DBoW2 vocabulary transform (oracle/bow_oracle.cpp)

def make_vocabulary(k=10, L=3, seed=0, stop_frac=0.02):
    """Synthetic vocabulary tree with DBoW2's shape (branching k, depth L, leaves = words): child descriptors are the
    parent's with random bit flips, idf-like weights, a few stop words (weight 0).  Returns the flattened arrays."""
    rng = np.random.default_rng(seed)
    desc = [rng.integers(0, 256, 32).astype(np.uint8)]
    child_start = [0]; children = []; level = [0]
    weight = [0.0]; word = [-1]
    frontier = [0]
    for lev in range(1, L + 1):
        nxt = []
        for parent in frontier:
            ids = list(range(len(desc), len(desc) + k))
            for _ in ids:
                d = desc[parent].copy()
                nb = max(2, 48 >> (lev - 1))
                d[rng.integers(0, 32, nb)] ^= (1 << rng.integers(0, 8, nb)).astype(np.uint8)
                desc.append(d); level.append(lev); weight.append(0.0); word.append(-1)
            nxt += ids
            while len(child_start) <= parent + 1:
                child_start.append(len(children))
            children += ids
            child_start[parent + 1] = len(children)
        frontier = nxt
    n = len(desc)
    cs = np.zeros(n + 1, np.int32)
    # children were appended parent by parent in id order: rebuild the CSR
    kids = {}
    pos = 0
    order = sorted(set(range(n)) - set(frontier))
    for p_ in order:
        kids[p_] = children[pos:pos + k]; pos += k
    flat = []
    for i in range(n):
        cs[i] = len(flat)
        flat += kids.get(i, [])
    cs[n] = len(flat)
    w = np.zeros(n); wd = np.full(n, -1, np.int32)
    for j, leaf in enumerate(frontier):
        wd[leaf] = j
        w[leaf] = 0.0 if rng.random() < stop_frac else float(rng.uniform(0.5, 9.0))
    return dict(child_start=cs, children=np.array(flat, np.int32), desc=np.stack(desc), weight=w, word=wd, L=L, k=k)


def bow_transform(voc, features, levelsup=4):
    f = _c(features, np.uint8); n = len(f)
    bw = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1)); fn = np.zeros(max(n, 1), np.int32)
    fs = np.zeros(n + 2, np.int32); ff = np.zeros(max(n, 1), np.int32); nfn = C.c_int32()
    fw = np.zeros(max(n, 1), np.int32); fnode = np.zeros(max(n, 1), np.int32)
    cs = _c(voc["child_start"], np.int32); ch = _c(voc["children"], np.int32); d = _c(voc["desc"], np.uint8)
    w = _c(voc["weight"], np.float64); wd = _c(voc["word"], np.int32)
    nw = lib().bow_oracle_transform(len(cs) - 1, _p(cs), _p(ch), _p(d), _p(w), _p(wd), int(voc["L"]), _p(f), n, int(levelsup),
                                    _p(bw), _p(bv), _p(fn), _p(fs), _p(ff), C.byref(nfn), _p(fw), _p(fnode))
    m = nfn.value
    return dict(words=bw[:nw], values=bv[:nw], fv_nodes=fn[:m], fv_start=fs[:m + 1], fv_features=ff[:fs[m]],
                feat_word=fw[:n], feat_node=fnode[:n])


# ---------------------------------------------------------------------------------------------------
# OptimizeSim3 (oracle/ba_oracle.cpp, namespace sim3o)

def sim3_exp(lie):
    v = _c(lie, np.float64); s = C.c_double(); R = np.zeros(9); t = np.zeros(3)
    lib().ba_oracle_sim3_exp(_p(v), C.byref(s), _p(R), _p(t))
    return s.value, R.reshape(3, 3), t


def sim3_log(s, R, t):
    out = np.zeros(7)
    lib().ba_oracle_sim3_log(C.c_double(s), _p(_c(R, np.float64)), _p(_c(t, np.float64)), _p(out))
    return out


def sim3_plus(x, delta):
    out = np.zeros(7)
    lib().ba_oracle_sim3_plus(_p(_c(x, np.float64)), _p(_c(delta, np.float64)), _p(out))
    return out


def sim3_error_term(lie, K4, obs, P, inv_sigma, do_inverse):
    r = np.zeros(2); J = np.zeros(14)
    lib().ba_oracle_sim3_error_term(_p(_c(lie, np.float64)), _p(_c(K4, np.float64)), _p(_c(obs, np.float64)),
                                    _p(_c(P, np.float64)), C.c_double(inv_sigma), int(do_inverse), _p(r), _p(J))
    return r, J.reshape(2, 7)


def optimize_sim3(s12, R12, t12, K1, K2, obs1, inv_sigma1, P3D2c, obs2, inv_sigma2, P3D1c, th2=10.0, max_iterations=100):
    n = len(obs1)
    s = C.c_double(s12); R = _c(R12, np.float64).reshape(-1).copy(); t = _c(t12, np.float64).copy()
    k1 = _c(K1, np.float64); k2 = _c(K2, np.float64)
    a = (_c(obs1, np.float32), _c(inv_sigma1, np.float32), _c(P3D2c, np.float64), _c(obs2, np.float32), _c(inv_sigma2, np.float32),
         _c(P3D1c, np.float64))
    bad = np.zeros(max(n, 1), np.uint8); lie = np.zeros(7); summ = BaSummary(); trace = np.zeros((max_iterations + 2, 8))
    fn = lib().ba_oracle_optimize_sim3
    fn.restype = C.c_int
    ret = fn(n, C.byref(s), _p(R), _p(t), _p(k1), _p(k2), *[_p(x) for x in a], C.c_float(th2), int(max_iterations), _p(bad),
             _p(lie), C.byref(summ), _p(trace), len(trace))
    return dict(ret=ret, s=s.value, R=R.reshape(3, 3), t=t, lie=lie, is_bad=bad[:n], iterations=summ.iterations,
                successful_steps=summ.successful_steps, termination=summ.termination, initial_cost=summ.initial_cost,
                final_cost=summ.final_cost, trace=trace)


# ---------------------------------------------------------------------------------------------------
# OptimizeEssentialGraph (oracle/ba_oracle.cpp)

def sim3_adjoint(lie):
    A = np.zeros(49)
    lib().ba_oracle_sim3_adjoint(_p(_c(lie, np.float64)), _p(A))
    return A.reshape(7, 7)


def essential_edge(Sji13, lie_j, lie_i):
    r = np.zeros(7); J = np.zeros(49)
    lib().ba_oracle_essential_edge(_p(_c(Sji13, np.float64)), _p(_c(lie_j, np.float64)), _p(_c(lie_i, np.float64)), _p(r), _p(J))
    return r, J.reshape(7, 7)


def essential_graph(Scw, kf_flags, Snc, edge_j, edge_i, edge_kind, Xw, ref_kf, max_iterations=100):
    """Scw / Snc: [n_kf][13] = scale, rotation (row-major), translation."""
    Scw = _c(Scw, np.float64); Snc = _c(Snc, np.float64); fl = _c(kf_flags, np.uint8)
    ej = _c(edge_j, np.int32); ei = _c(edge_i, np.int32); ek = _c(edge_kind, np.uint8)
    X = _c(Xw, np.float64).reshape(-1, 3); rk = _c(ref_kf, np.int32)
    n = len(Scw); m = len(X)
    lie = np.zeros((n, 7)); T = np.zeros((n, 16)); Xo = np.zeros((max(m, 1), 3)); summ = BaSummary()
    trace = np.zeros((max_iterations + 2, 8))
    fn = lib().ba_oracle_essential_graph
    fn.restype = C.c_int
    fn(n, _p(Scw), _p(fl), _p(Snc), len(ej), _p(ej), _p(ei), _p(ek), int(max_iterations), m, _p(X), _p(rk), _p(lie), _p(T), _p(Xo),
       C.byref(summ), _p(trace), len(trace))
    return dict(lie=lie, Tiw=T.reshape(n, 4, 4), Xw=Xo[:m], iterations=summ.iterations, successful_steps=summ.successful_steps,
                termination=summ.termination, initial_cost=summ.initial_cost, final_cost=summ.final_cost,
                jacobian_evaluations=summ.jacobian_evaluations, trace=trace)

"""Host-side mirror of the reference's CeresOptimizer (include/CeresOptimizer.h:353-387) over the C ABI, working
on flattened graphs instead of Frame*/KeyFrame*/MapPoint* pointers.  All compute is in
libcmos_b200.so (csrc/ba.cu); this file only marshals buffers.

Pose layout everywhere: the reference's 7-vector [t(3), q(x, y, z, w)], camera-from-world
(MatEigenConverter.cc:68-77)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr

TERMINATION = ("max_iterations", "function_tolerance", "parameter_tolerance", "gradient_tolerance", "stop_flag",
               "failure", "min_trust_region_radius")
TRACE_COLS = ("cost", "cost_change", "gradient_max_norm", "step_norm", "relative_decrease", "radius", "accepted",
              "valid")


class BaParams(C.Structure):
    _fields_ = [("max_cams", C.c_int32), ("max_points", C.c_int32), ("max_obs", C.c_int32),
                ("max_pairs_per_obs", C.c_int32), ("max_pose_batch", C.c_int32), ("max_pose_corr", C.c_int32),
                ("device", C.c_int32)]


class BaSummary(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("successful_steps", C.c_int32), ("termination", C.c_int32),
                ("jacobian_evaluations", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


SUMMARY_DTYPE = np.dtype([("iterations", "<i4"), ("successful_steps", "<i4"), ("termination", "<i4"),
                          ("jacobian_evaluations", "<i4"), ("initial_cost", "<f8"), ("final_cost", "<f8")])
assert SUMMARY_DTYPE.itemsize == C.sizeof(BaSummary)


class CeresOptimizer:
    """One engine handle (own CUDA stream and device buffers) per calling thread, like the per-thread context
    SURVEY.md §8b asks for; the reference's methods are static."""

    def __init__(self, max_cams: int = 64, max_points: int = 4096, max_obs: int = 32768, max_pairs_per_obs: int = 8,
                 max_pose_batch: int = 1, max_pose_corr: int = 2048, device: int = 0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        p = BaParams(max_cams, max_points, max_obs, max_pairs_per_obs, max_pose_batch, max_pose_corr, device)
        check(self._L.cmos_ba_create(C.byref(p), C.byref(self._h)))
        self.n_cams = self.n_points = self.n_obs = 0

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.cmos_ba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- int PoseOptimization(Frame*), CeresOptimizer.h:375 ----
    def PoseOptimization(self, pose7, xw, uv, inv_sigma2, K4, n_corr=None, max_iterations: int = 100):
        """Batch form: pose7 [B,7], xw [B,S,3], uv [B,S,2], inv_sigma2 [B,S], n_corr [B] (default S).
        Returns (pose7 [B,7], is_outlier [B,S] uint8, n_inliers [B], summaries [B])."""
        pose = np.array(pose7, np.float64, copy=True, order="C").reshape(-1, 7)
        B = len(pose)
        xw = np.ascontiguousarray(xw, np.float64).reshape(B, -1, 3)
        S = xw.shape[1]
        uv = np.ascontiguousarray(uv, np.float32).reshape(B, S, 2)
        w = np.ascontiguousarray(inv_sigma2, np.float32).reshape(B, S)
        n = np.full(B, S, np.int32) if n_corr is None else np.ascontiguousarray(n_corr, np.int32)
        K4 = np.ascontiguousarray(K4, np.float32)
        out = np.zeros((B, S), np.uint8); inl = np.zeros(B, np.int32); summ = np.zeros(B, SUMMARY_DTYPE)
        check(self._L.cmos_ba_pose_optimization(self._h, B, ptr(pose), ptr(n), ptr(xw), ptr(uv), ptr(w), S, ptr(K4),
                                                int(max_iterations), ptr(out), ptr(inl), ptr(summ), 0, None))
        return pose, out, inl, summ

    def pose_optimization_device(self, n_frames, d_pose7, d_n_corr, d_xw, d_uv, d_inv_sigma2, stride, K4,
                                 max_iterations, d_is_outlier, d_n_inliers, d_summaries=None, stream: int = 0):
        K4 = np.ascontiguousarray(K4, np.float32)
        check(self._L.cmos_ba_pose_optimization(self._h, n_frames, ptr(d_pose7), ptr(d_n_corr), ptr(d_xw), ptr(d_uv),
                                                ptr(d_inv_sigma2), stride, ptr(K4), int(max_iterations),
                                                ptr(d_is_outlier), ptr(d_n_inliers), ptr(d_summaries), 1,
                                                C.c_void_p(stream)))

    def pose_trace(self, frame: int, rows: int):
        t = np.zeros((rows, 8))
        check(self._L.cmos_ba_debug_pose_trace(self._h, frame, ptr(t), rows))
        return t

    # ---- graph upload ----
    # ---- OptimizeSim3(keyframe_1, keyframe_2, matches12, S12, th2, bFixScale), CeresOptimizer.h:372-375 ----
    def OptimizeSim3(self, s12, R12, t12, K1, K2, obs1, inv_sigma1, P3D2c, obs2, inv_sigma2, P3D1c, th2: float = 10.0,
                     max_iterations: int = 100):
        """-> dict(ret, s, R, t, lie, is_bad, summary).  Per correspondence: keypoint of keyframe 1 + its inv_level_sigma2 +
        keyframe 2's map point in camera-2 coordinates, and the same the other way round."""
        n = len(obs1)
        s = C.c_double(float(s12)); R = np.ascontiguousarray(R12, np.float64).reshape(-1).copy()
        t = np.ascontiguousarray(t12, np.float64).copy()
        k1 = np.ascontiguousarray(K1, np.float32); k2 = np.ascontiguousarray(K2, np.float32)
        a = (np.ascontiguousarray(obs1, np.float32), np.ascontiguousarray(inv_sigma1, np.float32),
             np.ascontiguousarray(P3D2c, np.float64), np.ascontiguousarray(obs2, np.float32),
             np.ascontiguousarray(inv_sigma2, np.float32), np.ascontiguousarray(P3D1c, np.float64))
        bad = np.zeros(max(n, 1), np.uint8); lie = np.zeros(7); ninl = C.c_int32(); summ = BaSummary()
        check(self._L.cmos_ba_optimize_sim3(self._h, n, C.byref(s), ptr(R), ptr(t), ptr(k1), ptr(k2), *[ptr(x) for x in a],
                                            C.c_float(th2), int(max_iterations), ptr(bad), ptr(lie), C.byref(ninl),
                                            C.byref(summ)))
        return dict(ret=ninl.value, s=s.value, R=R.reshape(3, 3), t=t, lie=lie, is_bad=bad[:n], summary=summ.as_dict())

    def OptimizeEssentialGraph(self, Scw, kf_flags, Snc, edge_j, edge_i, edge_kind, Xw, ref_kf, max_iterations: int = 100):
        """CeresOptimizer::OptimizeEssentialGraph over the flattened graph (see include/cmos_b200.h) ->
        dict(lie [n_kf][7], Tiw [n_kf][4][4], Xw [n_points][3], summary)."""
        Scw = np.ascontiguousarray(Scw, np.float64); Snc = np.ascontiguousarray(Snc, np.float64)
        fl = np.ascontiguousarray(kf_flags, np.uint8)
        ej = np.ascontiguousarray(edge_j, np.int32); ei = np.ascontiguousarray(edge_i, np.int32)
        ek = np.ascontiguousarray(edge_kind, np.uint8)
        X = np.ascontiguousarray(Xw, np.float64).reshape(-1, 3); rk = np.ascontiguousarray(ref_kf, np.int32)
        n, m = len(Scw), len(X)
        lie = np.zeros((n, 7)); T = np.zeros((n, 16)); Xo = np.zeros((max(m, 1), 3)); summ = BaSummary()
        check(self._L.cmos_ba_optimize_essential_graph(self._h, n, ptr(Scw), ptr(fl), ptr(Snc), len(ej), ptr(ej), ptr(ei), ptr(ek),
                                                       int(max_iterations), m, ptr(X), ptr(rk), ptr(lie), ptr(T), ptr(Xo),
                                                       C.byref(summ)))
        return dict(lie=lie, Tiw=T.reshape(n, 4, 4), Xw=Xo[:m], summary=summ.as_dict())

    def essential_graph_plan(self) -> dict:
        """How the last OptimizeEssentialGraph call solved its normal equations (cmos_ba_debug_essential_graph_plan)."""
        info = np.zeros(6, np.int32)
        check(self._L.cmos_ba_debug_essential_graph_plan(self._h, ptr(info)))
        return dict(zip(("nested_dissection", "keyframes_per_node", "node_unknowns", "nodes", "border_keyframes", "border_unknowns"),
                        (int(v) for v in info)))

    def set_problem(self, cams, cam_flags, points, obs_cam, obs_pt, uv, inv_sigma2, K4):
        cams = np.ascontiguousarray(cams, np.float64); points = np.ascontiguousarray(points, np.float64)
        cf = np.ascontiguousarray(cam_flags, np.uint8)
        oc = np.ascontiguousarray(obs_cam, np.int32); op = np.ascontiguousarray(obs_pt, np.int32)
        uv = np.ascontiguousarray(uv, np.float32); w = np.ascontiguousarray(inv_sigma2, np.float32)
        K4 = np.ascontiguousarray(K4, np.float32)
        check(self._L.cmos_ba_set_problem(self._h, len(cams), ptr(cams), ptr(cf), len(points), ptr(points), len(oc),
                                          ptr(oc), ptr(op), ptr(uv), ptr(w), ptr(K4)))
        self.n_cams, self.n_points, self.n_obs = len(cams), len(points), len(oc)

    def run_local(self, iters=(5, 10), stop_flag=None, stream: int = 0):
        check(self._L.cmos_ba_run_local(self._h, int(iters[0]), int(iters[1]), ptr(stop_flag), C.c_void_p(stream)))

    def run_global(self, n_iterations: int, robust: bool = True, stop_flag=None, stream: int = 0):
        check(self._L.cmos_ba_run_global(self._h, int(n_iterations), int(robust), ptr(stop_flag), C.c_void_p(stream)))

    def get_results(self, stream: int = 0):
        cams = np.zeros((self.n_cams, 7)); pts = np.zeros((self.n_points, 3))
        erase = np.zeros(self.n_obs, np.uint8); summ = np.zeros(2, SUMMARY_DTYPE)
        check(self._L.cmos_ba_get_results(self._h, ptr(cams), ptr(pts), ptr(erase), ptr(summ), C.c_void_p(stream)))
        return cams, pts, erase, summ

    def debug_stop_at(self, pass_: int, iteration: int):
        """Test hook: the device raises the stop flag itself after `iteration` iterations of `pass_` (-1: off)."""
        check(self._L.cmos_ba_debug_stop_at(self._h, int(pass_), int(iteration)))

    def trace(self, pass_: int, rows: int):
        t = np.zeros((rows, 8))
        check(self._L.cmos_ba_debug_trace(self._h, pass_, ptr(t), rows))
        return t

    # ---- void LocalBundleAdjustment(KeyFrame*, bool* stop_flag, Map*), CeresOptimizer.h:364 ----
    def LocalBundleAdjustment(self, cams, cam_flags, points, obs_cam, obs_pt, uv, inv_sigma2, K4, stop_flag=None):
        """Returns (cams, points, erase[n_obs], summaries[2]); erase marks the observations the reference removes
        from the map (CeresOptimizer.cc:573-581)."""
        self.set_problem(cams, cam_flags, points, obs_cam, obs_pt, uv, inv_sigma2, K4)
        self.run_local((5, 10), stop_flag)
        return self.get_results()

    # ---- void BundleAdjustment(...), GlobalBundleAdjustemnt(...), CeresOptimizer.h:355-362 ----
    def BundleAdjustment(self, cams, cam_const, points, obs_cam, obs_pt, uv, inv_sigma2, K4, n_iterations: int = 200,
                         stop_flag=None, is_robust: bool = True):
        self.set_problem(cams, cam_const, points, obs_cam, obs_pt, uv, inv_sigma2, K4)
        self.run_global(n_iterations, is_robust, stop_flag)
        cams, pts, _, summ = self.get_results()
        return cams, pts, summ[0]

    GlobalBundleAdjustemnt = BundleAdjustment   # [sic] the reference's spelling

    # ---- multi-GPU global BA (no reference counterpart) ----
    def comm_init(self, world: int, rank: int, device=None):
        """Collective: creates the engine's NCCL communicator; the 128-byte id travels over torch.distributed."""
        import torch
        import torch.distributed as dist
        idbuf = np.zeros(128, np.uint8)
        if rank == 0:
            check(self._L.cmos_ba_comm_unique_id(ptr(idbuf)))
        t = torch.from_numpy(idbuf)
        if device is not None:
            t = t.to(device)
        dist.broadcast(t, src=0)
        idbuf = np.ascontiguousarray(t.cpu().numpy())
        check(self._L.cmos_ba_comm_init(self._h, ptr(idbuf), world, rank))
        self.world, self.rank = world, rank

    def GlobalBundleAdjustemntSharded(self, cams, cam_const, points, obs_cam, obs_pt, uv, inv_sigma2, K4,
                                      n_iterations: int, world: int, rank: int, is_robust: bool = True, device=None):
        """Every rank passes the full graph; points/observations are partitioned here, keyframes replicated.
        Returns (cams, points, summary) with the full point array on every rank."""
        from . import sharding
        part = sharding.partition_graph(len(points), obs_cam, obs_pt, uv, inv_sigma2, world, rank)
        local_pts = np.ascontiguousarray(np.asarray(points, np.float64)[part["lo"]:part["hi"]])
        self.set_problem(cams, cam_const, local_pts, part["obs_cam"], part["obs_pt"], part["uv"], part["inv_sigma2"], K4)
        self.run_global(n_iterations, is_robust)
        cams_o, pts_o, _, summ = self.get_results()
        return cams_o, sharding.gather_points(pts_o, len(points), world, rank, device), summ[0]

    def launch_count(self) -> int:
        n = C.c_int32()
        check(self._L.cmos_ba_last_launch_count(self._h, C.byref(n)))
        return n.value

    def set_profiling(self, enable: bool = True):
        check(self._L.cmos_ba_set_profiling(self._h, int(enable)))

    def solve_time(self):
        ms, calls = C.c_double(), C.c_int64()
        check(self._L.cmos_ba_solve_time(self._h, C.byref(ms), C.byref(calls)))
        return ms.value, calls.value

"""Bundle-adjustment section of bench.py: BASELINE.json configs[2..4] on the GPU engine.

Metric: M residual-Jacobian evaluations per second = n_obs x sum over solves of (LM iterations + 1) / time of the
whole solve.  `value` is timed with the graph already on the device (CUDA events on the launching
stream, restore-from-initial included); `e2e` goes through the host-buffer C ABI: upload + structure build +
solve + download, wall clock."""
from __future__ import annotations

import time

import numpy as np

from . import synth
from .ceres_optimizer import CeresOptimizer

ALG_BYTES_PER_EVAL = 164          # 20 B observation in + 144 B of Jacobian blocks out


def _evals(n_obs, summaries):
    return n_obs * sum(int(s["iterations"]) + 1 for s in summaries)


def bench_pose(device: int, steps: int, warmup: int, batch: int):
    import torch
    K4 = np.array(synth.KITTI_K, np.float32)
    probs = [synth.make_pose_problem(1500, seed=3 + i) for i in range(batch)]
    S = 1500
    opt = CeresOptimizer(max_cams=1, max_points=1, max_obs=1, max_pose_batch=batch, max_pose_corr=S, device=device)
    pose = np.stack([p["pose"] for p in probs]); xw = np.stack([p["Xw"] for p in probs])
    uv = np.stack([p["uv"] for p in probs]); w = np.stack([p["inv_sigma2"] for p in probs])
    # host-buffer call (e2e) — also gives the iteration counts
    for _ in range(2):
        _, _, inl, summ = opt.PoseOptimization(pose, xw, uv, w, K4, max_iterations=4)
    t0 = time.perf_counter()
    for _ in range(steps):
        opt.PoseOptimization(pose, xw, uv, w, K4, max_iterations=4)
    e2e_s = (time.perf_counter() - t0) / steps
    evals = sum(S * (int(s["iterations"]) + 1) for s in summ)
    # device-resident
    dev = torch.device("cuda", device)
    d_pose0 = torch.from_numpy(pose).to(dev); d_pose = d_pose0.clone()
    d_xw = torch.from_numpy(xw).to(dev); d_uv = torch.from_numpy(uv).to(dev); d_w = torch.from_numpy(w).to(dev)
    d_n = torch.full((batch,), S, dtype=torch.int32, device=dev)
    d_out = torch.zeros((batch, S), dtype=torch.uint8, device=dev); d_inl = torch.zeros(batch, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for it in range(warmup + steps):
            if it == warmup:
                e0.record(stream)
            d_pose.copy_(d_pose0, non_blocking=True)
            opt.pose_optimization_device(batch, d_pose, d_n, d_xw, d_uv, d_w, S, K4, 4, d_out, d_inl,
                                         stream=stream.cuda_stream)
        e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / steps
    assert np.array_equal(d_inl.cpu().numpy(), inl)
    return {"config": f"configs[2]: PoseOptimization, {batch} frame(s) x 1500 correspondences, 4 LM iterations (Huber)",
            "value": evals / (ms * 1e-3) / 1e6, "unit": "Mresid/s", "ms_per_solve": ms, "frames": batch,
            "iterations": [int(s["iterations"]) for s in summ][:4],
            "e2e": {"value": evals / e2e_s / 1e6, "unit": "Mresid/s", "ms_per_solve": e2e_s * 1e3},
            "gpu_launches_per_solve": 1}


def _bench_graph(opt, G, K4, steps, warmup, local: bool, iters, label):
    import torch
    flags = G["fixed"]
    args = (G["poses"], flags, G["points"], G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"], K4)
    n_obs = len(G["obs_cam"])
    opt.set_problem(*args)
    run = (lambda: opt.run_local(iters)) if local else (lambda: opt.run_global(iters, True))
    for _ in range(warmup):
        run()
    _, _, _, summ = opt.get_results()
    opt.set_profiling(True)
    for _ in range(steps):
        run()
    c1, p1, e1_, summ = opt.get_results()
    ms_total, calls = opt.solve_time()
    opt.set_profiling(False)
    ms = ms_total / max(calls, 1)
    summaries = list(summ) if local else [summ[0]]
    evals = _evals(n_obs, summaries)
    launches = opt.launch_count()
    # e2e: upload + structure + solve + download through the host-buffer ABI.  The headline number rebuilds the structure
    # every call (a new local map each time, as in LocalMapping); `same_topology` is the repeated-graph case, where
    # cmos_ba_set_problem recognises the unchanged index arrays and uploads values only.
    import os
    n_e2e = max(1, min(steps, 5))

    def timed_e2e():
        opt.set_problem(*args); run(); opt.get_results()          # warm
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            opt.set_problem(*args)
            run()
            res = opt.get_results()
        return (time.perf_counter() - t0) / n_e2e, res
    os.environ["CMOS_BA_NO_TOPO_CACHE"] = "1"
    try:
        e2e_s, (c2, p2, _, _) = timed_e2e()
    finally:
        del os.environ["CMOS_BA_NO_TOPO_CACHE"]
    e2e_same_s, (c3, p3, _, _) = timed_e2e()
    assert np.array_equal(c1, c2) and np.array_equal(p1, p2), "BA engine is not deterministic run to run"
    assert np.array_equal(c1, c3) and np.array_equal(p1, p3), "values-only upload changed the result"
    return {"config": label, "value": evals / (ms * 1e-3) / 1e6, "unit": "Mresid/s", "ms_per_solve": ms,
            "iterations": [int(s["iterations"]) for s in summaries],
            "successful_steps": [int(s["successful_steps"]) for s in summaries],
            "cost": [[float(s["initial_cost"]), float(s["final_cost"])] for s in summaries],
            "evals_per_solve": evals,
            "roofline": {"bound": "hbm", "algorithmic_bytes_per_eval": ALG_BYTES_PER_EVAL,
                         "achieved_gbs": evals * ALG_BYTES_PER_EVAL / (ms * 1e-3) / 1e9,
                         "note": "latency-bound by construction (SURVEY.md §7 hard part 6)"},
            "e2e": {"value": evals / e2e_s / 1e6, "unit": "Mresid/s", "ms_per_solve": e2e_s * 1e3,
                    "call": "cmos_ba_set_problem (structure build, one packed upload) + run + cmos_ba_get_results, host buffers",
                    "same_topology": {"value": evals / e2e_same_s / 1e6, "unit": "Mresid/s", "ms_per_solve": e2e_same_s * 1e3}},
            "gpu_launches_per_solve": launches}


def bench_essential_graph(device: int, steps: int, n_kf: int = 1000, n_points: int = 100000):
    """CeresOptimizer::OptimizeEssentialGraph after a loop closure over n_kf keyframes (spanning tree + three co-visibility
    edges per keyframe + loop connections) and the correction of n_points map points; the call is synchronous over host
    buffers, so this is an end-to-end time."""
    G = synth.make_essential_graph_problem(n_kf, seed=8, n_group=10, covis=(2, 3, 5), n_points=n_points)
    a = (G["Scw"], G["kf_flags"], G["Snc"], G["edge_j"], G["edge_i"], G["edge_kind"], G["Xw"], G["ref_kf"])
    opt = CeresOptimizer(max_cams=2, max_points=8, max_obs=8, device=device)
    got = opt.OptimizeEssentialGraph(*a)               # warm-up: allocates the storage
    best = None
    for _ in range(steps):
        t0 = time.perf_counter()
        got = opt.OptimizeEssentialGraph(*a)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    launches = opt.launch_count()
    s = got["summary"]
    opt.close()
    E = len(G["edge_j"])
    return {"config": f"OptimizeEssentialGraph, {n_kf} keyframes x {E} edges (7 residuals each), {n_points} map points corrected, "
                      f"max 100 LM iterations", "ms_per_solve": best * 1e3, "iterations": s["iterations"],
            "successful_steps": s["successful_steps"], "termination": s["termination"],
            "cost": [s["initial_cost"], s["final_cost"]], "unknowns": 7 * (n_kf - 1),
            "value": E * s["jacobian_evaluations"] / best / 1e6, "unit": "M edge linearisations/s",
            "gpu_launches_per_solve": launches, "timing": "wall clock around the synchronous C-ABI call, host buffers, best of %d" % steps}


def run(device: int, world: int, args):
    """Returns the "ba" object of bench.py's JSON line (rank 0 formats it; every rank runs its replica)."""
    steps = max(3, min(args.steps, 20)); warmup = 3
    K4 = np.array(synth.KITTI_K, np.float32)
    out = {}
    out["pose_optimization"] = bench_pose(device, steps, warmup, 1)
    out["pose_optimization_batch64"] = bench_pose(device, steps, warmup, 64)
    G = synth.make_ba_problem(20, 3000, 4, seed=4)
    opt = CeresOptimizer(max_cams=20, max_points=3000, max_obs=12000, device=device)
    out["local_ba"] = _bench_graph(opt, G, K4, steps, warmup, True, (5, 10),
                                   "configs[3]: LocalBundleAdjustment, 20 keyframes x 3000 points x 12000 observations, "
                                   "5 Huber + 10 LM iterations")
    opt.close()
    out["essential_graph"] = bench_essential_graph(device, max(2, min(steps, 5)))
    if not getattr(args, "no_global", False):
        G = synth.make_ba_problem_fast(1000, 100000, 5, seed=5)
        opt = CeresOptimizer(max_cams=1000, max_points=100000, max_obs=500000, device=device)
        label = (f"configs[4]: GlobalBundleAdjustemnt, 1000 keyframes x 100000 points x 500000 observations, "
                 f"{args.global_iters} LM iterations (Huber)")
        if world == 1:
            out["global_ba"] = _bench_graph(opt, G, K4, max(2, min(steps, 3)), 1, False, args.global_iters, label)
        else:
            out["global_ba"] = _bench_global_sharded(opt, G, K4, max(2, min(steps, 3)), args.global_iters, label, device, world)
        opt.close()
    return out


def _bench_global_sharded(opt, G, K4, steps, iters, label, device, world):
    """Points/observations partitioned over the ranks, one NCCL all-reduce of the reduced camera system per LM
    iteration (strong scaling of ONE problem; time = max over ranks)."""
    import torch
    import torch.distributed as dist
    from . import sharding
    rank = dist.get_rank()
    dev = torch.device("cuda", device)
    opt.comm_init(world, rank, dev)
    part = sharding.partition_graph(len(G["points"]), G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"], world, rank)
    local_pts = np.ascontiguousarray(G["points"][part["lo"]:part["hi"]])
    opt.set_problem(G["poses"], G["fixed"], local_pts, part["obs_cam"], part["obs_pt"], part["uv"], part["inv_sigma2"], K4)
    opt.run_global(iters, True)
    opt.get_results()
    dist.barrier()
    opt.set_profiling(True)
    for _ in range(steps):
        opt.run_global(iters, True)
    _, _, _, summ = opt.get_results()
    ms_total, calls = opt.solve_time()
    opt.set_profiling(False)
    ms = ms_total / max(calls, 1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    evals = len(G["obs_cam"]) * (int(summ[0]["iterations"]) + 1)
    # equality evidence carried by the bench line itself: rank 0 also solves the UNSHARDED problem once (outside the timed
    # region, own handle, no communicator) and compares the replicated keyframe poses, its own point range and the costs
    cams_s, pts_s, _, _ = opt.get_results()
    check = None
    if rank == 0:
        solo = CeresOptimizer(max_cams=len(G["poses"]), max_points=len(G["points"]), max_obs=len(G["obs_cam"]), device=device)
        c1, p1, s1 = solo.BundleAdjustment(G["poses"], G["fixed"], G["points"], G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"],
                                           K4, n_iterations=iters, is_robust=True)
        solo.close()
        p1 = p1[part["lo"]:part["hi"]]
        check = {"max_rel_diff_vs_1gpu": float(max(np.abs(cams_s - c1).max() / max(1.0, np.abs(c1).max()),
                                                   np.abs(pts_s - p1).max() / max(1.0, np.abs(p1).max()))),
                 "final_cost_rel_diff": float(abs(float(summ[0]["final_cost"]) / float(s1["final_cost"]) - 1.0)),
                 "iterations_equal": bool(int(summ[0]["iterations"]) == int(s1["iterations"]) and
                                          int(summ[0]["successful_steps"]) == int(s1["successful_steps"]))}
    dist.barrier()
    return {"config": label + f", points sharded over {world} GPUs, NCCL all-reduces per LM iteration: H_cc/g_c + scalars, "
                              f"reduced camera system, step statistics", "vs_1gpu": check,
            "value": evals / (ms * 1e-3) / 1e6, "unit": "Mresid/s", "ms_per_solve": ms, "scaling": "strong",
            "iterations": [int(summ[0]["iterations"])], "cost": [[float(summ[0]["initial_cost"]), float(summ[0]["final_cost"])]],
            "evals_per_solve": evals, "gpu_launches_per_solve": opt.launch_count()}

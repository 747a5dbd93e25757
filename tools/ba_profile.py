"""Small driver for ncu: a few LocalBA (configs[3]) solves and optionally a short GlobalBA (configs[4]) solve."""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from ceres_mono_orb_slam2_b200 import CeresOptimizer, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "local"
K4 = np.array(synth.KITTI_K, np.float32)
if which == "local":
    G = synth.make_ba_problem(20, 3000, 4, seed=4)
    opt = CeresOptimizer(max_cams=20, max_points=3000, max_obs=12000)
    opt.set_problem(G["poses"], G["fixed"], G["points"], G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"], K4)
    for _ in range(3):
        opt.run_local((5, 10))
    print(opt.get_results()[3])
elif which == "pose":
    P = synth.make_pose_problem(1500, seed=3)
    opt = CeresOptimizer(max_pose_batch=1, max_pose_corr=1500)
    for _ in range(3):
        print(opt.PoseOptimization(P["pose"][None], P["Xw"][None], P["uv"][None], P["inv_sigma2"][None], K4, max_iterations=4)[3])
elif which == "essential":
    import time
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    E = synth.make_essential_graph_problem(n, seed=8, n_group=10, covis=(2, 3, 5), n_points=100000)
    opt = CeresOptimizer(max_cams=2, max_points=8, max_obs=8)
    a = (E["Scw"], E["kf_flags"], E["Snc"], E["edge_j"], E["edge_i"], E["edge_kind"], E["Xw"], E["ref_kf"])
    for _ in range(2):
        t0 = time.perf_counter()
        r = opt.OptimizeEssentialGraph(*a)
        print(r["summary"], opt.launch_count(), "launches", (time.perf_counter() - t0) * 1e3, "ms")
elif which == "global_time":
    from ceres_mono_orb_slam2_b200.ba_bench import _bench_graph
    G = synth.make_ba_problem_fast(1000, 100000, 5, seed=5)
    opt = CeresOptimizer(max_cams=1000, max_points=100000, max_obs=500000)
    r = _bench_graph(opt, G, K4, 3, 1, False, 10, "x")
    print("global BA: %.2f ms per 10-iteration solve, %.1f Mresid/s, %d launches" % (r["ms_per_solve"], r["value"], r["gpu_launches_per_solve"]))
elif which == "sim3":
    P = synth.make_sim3_problem(n=300, seed=6)
    opt = CeresOptimizer(max_cams=2, max_points=8, max_obs=8)
    for _ in range(3):
        print(opt.OptimizeSim3(P["s0"], P["R0"], P["t0"], P["K"], P["K"], P["obs1"], P["inv_sigma1"], P["P3D2c"], P["obs2"],
                               P["inv_sigma2"], P["P3D1c"])["summary"])
else:
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    G = synth.make_ba_problem_fast(1000, 100000, 5, seed=5)
    opt = CeresOptimizer(max_cams=1000, max_points=100000, max_obs=500000)
    opt.set_problem(G["poses"], G["fixed"], G["points"], G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"], K4)
    for _ in range(2):
        opt.run_global(iters, True)
    print(opt.get_results()[3])

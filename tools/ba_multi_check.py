"""torchrun driver: sharded global bundle adjustment on WORLD_SIZE GPUs must equal the single-GPU solve.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P tools/ba_multi_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ceres_mono_orb_slam2_b200 import CeresOptimizer, synth  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
K4 = np.array(synth.KITTI_K, np.float32)
ok = True
# the third case takes the blocked (out-of-shared-memory) Cholesky path and has co-visibility blocks that only one
# rank's points touch: the ranks must agree on the union block list
# the fourth case (260 keyframes, W = 20) takes the nested-dissection (block cyclic reduction) solver; the fifth starts so far
# from the optimum that the trust region REJECTS steps: no re-linearisation happens in the following iteration, while the
# collectives still run (the non-root ranks must then contribute zeros, see k_post_lin)
rejected_seen = False
for n_cams, n_points, window, iters, noise in [(12, 400, None, 8, None), (60, 2500, 6, 8, None), (150, 3000, 10, 6, None),
                                                (260, 15600, 10, 6, None), (40, 1500, 6, 12, (0.08, 0.6))]:
    kw = {} if noise is None else {"pose_noise": noise}
    if n_cams >= 100:
        G = synth.make_ba_problem_fast(n_cams, n_points, 5, seed=31 + n_cams, window=window, **kw)
    else:
        G = synth.make_ba_problem(n_cams, n_points, 5, seed=31 + n_cams, window=window, **kw)
    opt = CeresOptimizer(max_cams=n_cams, max_points=n_points, max_obs=len(G["obs_cam"]), device=local)
    opt.comm_init(world, rank, dev)
    cams, pts, s = opt.GlobalBundleAdjustemntSharded(G["poses"], G["fixed"], G["points"], G["obs_cam"], G["obs_pt"], G["uv"],
                                                     G["inv_sigma2"], K4, iters, world, rank, True, dev)
    ref = CeresOptimizer(max_cams=n_cams, max_points=n_points, max_obs=len(G["obs_cam"]), device=local)
    rc, rp, rs = ref.BundleAdjustment(G["poses"], G["fixed"], G["points"], G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"],
                                      K4, n_iterations=iters)
    e_c = np.abs(cams - rc).max() / np.abs(rc).max(); e_p = np.abs(pts - rp).max() / np.abs(rp).max()
    same = s["iterations"] == rs["iterations"] and s["successful_steps"] == rs["successful_steps"]
    print(f"rank {rank}: {n_cams} keyframes: iterations {s['iterations']} vs {rs['iterations']}, rel err cams {e_c:.2e} points {e_p:.2e}, "
          f"cost {s['final_cost']:.9e} vs {rs['final_cost']:.9e}", flush=True)
    ok = ok and same and e_c < 1e-8 and e_p < 1e-8
    rejected_seen = rejected_seen or s["successful_steps"] < s["iterations"]
    # keyframes must be replicated bit for bit across ranks
    t = torch.from_numpy(cams).to(dev); t0 = t.clone(); dist.broadcast(t0, src=0)
    ok = ok and bool(torch.equal(t, t0))
    opt.close(); ref.close()
if rank == 0:
    print("a case with rejected steps was exercised:", rejected_seen, flush=True)
flag = torch.tensor([1.0 if ok else 0.0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_OK" if flag.item() == 1.0 else "MULTI_FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)

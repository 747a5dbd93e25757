"""World-size-2 gloo tests (CPU) of the host-side multi-GPU logic: frame sharding and throughput aggregation for
the ORB path, point partitioning / reassembly for the sharded global bundle adjustment, and — with the CPU
oracle's per-observation Jacobians — the identity the engine's one exchange step relies on: the reduced camera
system is the SUM over point shards of the shards' partial systems."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ceres_mono_orb_slam2_b200 import sharding, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _partial_system(G, sel, pts_lo, pts_hi):
    """Dense normal-equation pieces of one shard from the oracle's Jacobians: H_cc (sum of Jc'Jc per camera),
    and the Schur term sum_j W_j Hpp_j^-1 W_j' over the shard's points (unit weights, no damping)."""
    from oracle import pyoracle as po
    K = len(G["poses"]); n = 6 * K
    Hcc = np.zeros((n, n)); schur = np.zeros((n, n))
    by_pt = {}
    for i in sel:
        r, Jc, Jp = po.ba_residual(G["poses"][G["obs_cam"][i]], G["points"][G["obs_pt"][i]], G["K"], G["uv"][i, 0], G["uv"][i, 1],
                                   G["inv_sigma2"][i])
        c = G["obs_cam"][i]
        Hcc[6 * c:6 * c + 6, 6 * c:6 * c + 6] += Jc.T @ Jc
        by_pt.setdefault(int(G["obs_pt"][i]), []).append((c, Jc, Jp))
    for j, lst in by_pt.items():
        assert pts_lo <= j < pts_hi
        Hpp = sum(Jp.T @ Jp for _, _, Jp in lst) + 1e-3 * np.eye(3)
        Hi = np.linalg.inv(Hpp)
        for ca, Jca, Jpa in lst:
            for cb, Jcb, Jpb in lst:
                schur[6 * ca:6 * ca + 6, 6 * cb:6 * cb + 6] += (Jca.T @ Jpa) @ Hi @ (Jcb.T @ Jpb).T
    return Hcc, schur


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- ORB: contiguous frame shards cover the batch exactly once
        lo, hi = sharding.shard_range(64, world, rank)
        cover = torch.zeros(64); cover[lo:hi] = 1
        dist.all_reduce(cover)
        assert torch.all(cover == 1)
        units, ms = sharding.aggregate_throughput(1000.0 * (rank + 1), 2.0 + rank, world)
        assert units == 1000.0 * sum(r + 1 for r in range(world)) and ms == 2.0 + world - 1
        # ---- global BA: partition, exchange identity, reassembly
        G = synth.make_ba_problem(4, 40, 3, seed=7)
        part = sharding.partition_graph(len(G["points"]), G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"], world, rank)
        n_obs = torch.tensor([float(len(part["obs_index"]))]); dist.all_reduce(n_obs)
        assert int(n_obs[0]) == len(G["obs_cam"])
        assert np.array_equal(G["obs_pt"][part["obs_index"]] - part["lo"], part["obs_pt"])
        assert part["obs_pt"].min() >= 0 and part["obs_pt"].max() < part["hi"] - part["lo"]
        Hcc, schur = _partial_system(G, part["obs_index"], part["lo"], part["hi"])
        t = torch.from_numpy(np.stack([Hcc, schur])); dist.all_reduce(t)
        if rank == 0:
            Hf, Sf = _partial_system(G, np.arange(len(G["obs_cam"])), 0, len(G["points"]))
            assert np.allclose(t[0].numpy(), Hf, rtol=1e-12, atol=1e-9)
            assert np.allclose(t[1].numpy(), Sf, rtol=1e-12, atol=1e-9)
        # every rank "optimises" its points (here: a marker transform); gather restores the global order
        local = G["points"][part["lo"]:part["hi"]] * 2.0 + rank
        full = sharding.gather_points(local, len(G["points"]), world, rank)
        expect = G["points"] * 2.0
        for r in range(world):
            a, b = sharding.shard_range(len(G["points"]), world, r)
            expect[a:b] += r
        assert np.array_equal(full, expect)
        q.put((rank, "ok"))
    except Exception as e:   # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_shard_range_is_balanced_and_contiguous():
    for n in (0, 1, 7, 64, 100000):
        for w in (1, 2, 3, 8):
            rs = [sharding.shard_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1

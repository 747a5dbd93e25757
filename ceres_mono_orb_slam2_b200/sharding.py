"""Host-side sharding logic for the multi-GPU paths.  Pure numpy + torch.distributed, so that it
is covered by world_size-2 gloo tests on CPU (tests/test_multiprocess.py).

* ORB extract + match: frames are independent -> contiguous frame ranges per rank, no collective.
* Global bundle adjustment: map points are partitioned into contiguous ranges; every observation follows its
  point; keyframes are replicated.  The per-iteration exchange happens inside the engine (NCCL all-reduce of the
  reduced camera system); this module prepares each rank's slice and reassembles the optimised points."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n items for `rank` (first n % world ranks get one extra)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def partition_graph(n_points: int, obs_cam, obs_pt, uv, inv_sigma2, world: int, rank: int):
    """Slice of the observation arrays that belongs to `rank` (points in its contiguous range), with point
    indices rebased to the local range.  Returns dict(lo, hi, obs_index, obs_cam, obs_pt, uv, inv_sigma2)."""
    obs_pt = np.asarray(obs_pt)
    lo, hi = shard_range(n_points, world, rank)
    sel = np.nonzero((obs_pt >= lo) & (obs_pt < hi))[0]
    return dict(lo=lo, hi=hi, obs_index=sel, obs_cam=np.ascontiguousarray(np.asarray(obs_cam)[sel], np.int32),
                obs_pt=np.ascontiguousarray(obs_pt[sel] - lo, np.int32),
                uv=np.ascontiguousarray(np.asarray(uv)[sel], np.float32),
                inv_sigma2=np.ascontiguousarray(np.asarray(inv_sigma2)[sel], np.float32))


def gather_points(local_points: np.ndarray, n_points: int, world: int, rank: int, device=None) -> np.ndarray:
    """All ranks end with the full [n_points, 3] array (ranges are contiguous and ordered by rank)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local_points
    sizes = [shard_range(n_points, world, r) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((width, 3), dtype=torch.float64, device=device)
    buf[: len(local_points)] = torch.from_numpy(np.ascontiguousarray(local_points)).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[: hi - lo].cpu().numpy() for o, (lo, hi) in zip(out, sizes)], 0)


def aggregate_throughput(units_local: float, ms_local: float, world: int, device=None) -> tuple[float, float]:
    """bench.py's contract: whole-job units (SUM over ranks) and the slowest rank's time (MAX over ranks)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return units_local, ms_local
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    u = torch.tensor([units_local], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u[0]), float(t[0])

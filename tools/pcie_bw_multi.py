"""Aggregate page-locked host->device copy bandwidth with every rank copying AT THE SAME TIME (the bound of the end-to-end
number at N GPUs):  python -m torch.distributed.run --nproc-per-node N tools/pcie_bw_multi.py
Each rank copies 41 MB (one bench step's upload) 40 times between two barriers; rank 0 prints per-rank and aggregate GB/s,
first alone (ranks take turns), then all together, then all together with simultaneous device->host copies of 8 MB."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 41 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(8 * 1000 * 1000, dtype=torch.uint8).pin_memory(); d2 = torch.empty(8 * 1000 * 1000, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream(); s2 = torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run(active, duplex, reps=40):
    barrier()
    t0 = time.perf_counter()
    if active:
        with torch.cuda.stream(s):
            for _ in range(reps):
                d.copy_(h, non_blocking=True)
        if duplex:
            with torch.cuda.stream(s2):
                for _ in range(reps):
                    h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([n * reps / dt / 1e9 if active else 0.0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(gbs)
    barrier()
    return n * reps / dt / 1e9 if active else 0.0, float(gbs[0])


run(True, False, 5)
alone = []
for r in range(world):
    mine, _ = run(rank == r, False)
    t = torch.tensor([mine], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t)
    alone.append(float(t[0]))
_, together = run(True, False)
_, duplex = run(True, True)
if rank == 0:
    print(f"ranks {world}: host->device alone per rank {[round(a, 1) for a in alone]} GB/s; all together {together:.1f} GB/s aggregate "
          f"({together / world:.1f} per rank); with simultaneous device->host copies {duplex:.1f} GB/s aggregate; host cores {os.cpu_count()}")
if world > 1:
    dist.destroy_process_group()

"""Host<->device copy bandwidth with page-locked buffers (the bound of bench.py's e2e number): python tools/pcie_bw.py"""
import torch

for mb in (10, 41, 128):
    n = mb * 1000 * 1000
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    for direction in ("h2d", "d2h"):
        with torch.cuda.stream(s):
            for _ in range(3):
                (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(10):
                (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
            e1.record(s)
        s.synchronize()
        print(f"{direction} {mb:4d} MB: {n * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e9:6.1f} GB/s")

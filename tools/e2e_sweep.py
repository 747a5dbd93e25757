"""Sweep of the end-to-end front-end call (cmos_track_frames) over stream lanes x chunk size, configs[1] batch of 64.
    python tools/e2e_sweep.py > gpurun_out/e2e_sweep.txt"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ceres_mono_orb_slam2_b200 import KP_DTYPE, Camera, ORBextractor, TrackingFrontEnd, synth  # noqa: E402

B = 64
W, H = bench.W, bench.H
frames, offs = bench.make_batch(B, seed=1000)
ext = ORBextractor(bench.NFEAT, bench.SCALE, bench.NLEVELS, bench.INI_TH, bench.MIN_TH, max_width=W, max_height=H, max_batch=B)
cap = ext.capacity
cam = Camera.create(W, H, synth.KITTI_K, ext.GetScaleFactors(), bench.SCALE)
kps, desc, counts = ext.extract_batch(frames)
lk, lcounts, flags, xw, mdesc, T = bench.make_last_views(kps, desc, counts, offs, cap, seed=5000)
ext.close()


def pinned(a):
    t = torch.empty(a.view(np.uint8).shape if a.dtype == KP_DTYPE else a.shape,
                    dtype=torch.uint8 if a.dtype == KP_DTYPE else torch.from_numpy(a[:0].copy()).dtype, pin_memory=True)
    v = t.numpy().view(KP_DTYPE).reshape(a.shape) if a.dtype == KP_DTYPE else t.numpy()
    v[...] = a
    return t, v


keep = []
ins = []
for a in (frames, T, lk, lcounts, flags, xw, mdesc):
    t, v = pinned(np.ascontiguousarray(a)); keep.append(t); ins.append(v)
outs = []
for a in (np.zeros((B, cap), KP_DTYPE), np.zeros((B, cap, 32), np.uint8), np.zeros(B, np.int32), np.zeros((B, cap), np.int32),
          np.zeros(B, np.int32)):
    t, v = pinned(a); keep.append(t); outs.append(v)
feats = int(counts.sum())
ref_nm = None
for lanes, chunk in [(1, 64), (2, 32), (2, 16), (3, 16), (4, 16), (2, 8), (3, 8), (4, 8), (3, 4), (4, 4), (6, 4), (4, 2)]:
    fe = TrackingFrontEnd(cam, bench.NFEAT, bench.SCALE, bench.NLEVELS, bench.INI_TH, bench.MIN_TH, max_width=W, max_height=H,
                          lanes=lanes, chunk_frames=chunk)
    for _ in range(3):
        fe.track(*ins, bench.TH_PROJ, out=tuple(outs))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        fe.track(*ins, bench.TH_PROJ, out=tuple(outs))
    dt = (time.perf_counter() - t0) / n
    nm = int(outs[4].sum())
    ref_nm = nm if ref_nm is None else ref_nm
    assert nm == ref_nm and int(outs[2].sum()) == feats
    print(f"lanes {lanes} chunk {chunk:3d}: {dt * 1e3:7.3f} ms/step  {feats / dt / 1e6:7.2f} Mfeat/s  launches {fe.launch_count()}", flush=True)
    fe.close()

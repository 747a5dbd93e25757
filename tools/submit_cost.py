"""Host cost of the pipelined end-to-end call: wall time inside cmos_track_submit / cmos_track_wait per 64-frame step."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ceres_mono_orb_slam2_b200 import KP_DTYPE, Camera, ORBextractor, TrackingFrontEnd, synth
B = 64; W, H = bench.W, bench.H
frames, offs = bench.make_batch(B, seed=1000)
ext = ORBextractor(bench.NFEAT, bench.SCALE, bench.NLEVELS, bench.INI_TH, bench.MIN_TH, max_width=W, max_height=H, max_batch=B)
cap = ext.capacity
cam = Camera.create(W, H, synth.KITTI_K, ext.GetScaleFactors(), bench.SCALE)
kps, desc, counts = ext.extract_batch(frames)
lk, lcounts, flags, xw, mdesc, T = bench.make_last_views(kps, desc, counts, offs, cap, seed=5000)
ext.close()
def pinned(a):
    t = torch.empty(a.view(np.uint8).shape if a.dtype == KP_DTYPE else a.shape, dtype=torch.uint8 if a.dtype == KP_DTYPE else torch.from_numpy(a[:0].copy()).dtype, pin_memory=True)
    v = t.numpy().view(KP_DTYPE).reshape(a.shape) if a.dtype == KP_DTYPE else t.numpy()
    v[...] = a
    return t, v
keep = []; ins = []
for a in (frames, T, lk, lcounts, flags, xw, mdesc):
    t, v = pinned(np.ascontiguousarray(a)); keep.append(t); ins.append(v)
outs = [[], []]
for k in range(2):
    for a in (np.zeros((B, cap), KP_DTYPE), np.zeros((B, cap, 32), np.uint8), np.zeros(B, np.int32), np.zeros((B, cap), np.int32), np.zeros(B, np.int32)):
        t, v = pinned(a); keep.append(t); outs[k].append(v)
for lanes, chunk in ((8, 16), (4, 16), (2, 32)):
    fe = TrackingFrontEnd(cam, bench.NFEAT, bench.SCALE, bench.NLEVELS, bench.INI_TH, bench.MIN_TH, max_width=W, max_height=H, lanes=lanes, chunk_frames=chunk)
    def run(n):
        ts = tw = 0.0; pending = None
        for i in range(n):
            t0 = time.perf_counter(); tk = fe.submit(*ins, bench.TH_PROJ, out=tuple(outs[i & 1])); ts += time.perf_counter() - t0
            if pending is not None:
                t0 = time.perf_counter(); fe.wait(pending); tw += time.perf_counter() - t0
            pending = tk
        t0 = time.perf_counter(); fe.wait(pending); tw += time.perf_counter() - t0
        return ts / n, tw / n
    run(3); torch.cuda.synchronize()
    t0 = time.perf_counter(); s, w = run(20); tot = (time.perf_counter() - t0) / 20
    print(f"lanes {lanes} chunk {chunk}: step {tot*1e3:.3f} ms, inside submit {s*1e3:.3f} ms, inside wait {w*1e3:.3f} ms, launches/step {fe.launch_count()}, host cores {os.cpu_count()}")
    fe.close()

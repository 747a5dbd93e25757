"""compute-sanitizer over one small invocation of every hot path (SURVEY.md §5 build notes): memcheck (out-of-bounds / misaligned
accesses, API errors) and racecheck (shared-memory hazards — the engine relies on named barriers, warp-specialised phases and
"last CTA" tails).  The workload is tiny because the tools slow kernels down by one to two orders of magnitude.

Not in the workload: round 1's serial band solver (k_band_backsub, used only for bands of fewer than 8 nodes).  racecheck reports
its TMA ring as a WARNING (0 errors): a write-after-read between the lanes' generic-proxy reads of a ring slot and lane 0's
cp.async.bulk refill of that slot, which the kernel orders with __syncwarp() + fence.proxy.async — the tool does not model the
proxy fence."""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SNIPPET = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
from ceres_mono_orb_slam2_b200 import Camera, CeresOptimizer, ORBextractor, ORBmatcher, synth
W, H = 320, 240
frames, offs = synth.make_sequence(W, H, 2, 5, return_offsets=True)
ext = ORBextractor(300, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=2)
kps, desc, counts = ext.extract_batch(frames)
cap = ext.capacity
cam = Camera.create(W, H, synth.TUM2_K, ext.GetScaleFactors(), 1.2)
n0 = int(counts[0])
fl, xw, md = synth.make_last_frame_view(kps[0, :n0], desc[0, :n0], (offs[0] - offs[1]).astype(float), 9, K=synth.TUM2_K)
flags = np.zeros((1, cap), np.uint8); X = np.zeros((1, cap, 3)); D = np.zeros((1, cap, 32), np.uint8)
flags[0, :n0] = fl; X[0, :n0] = xw; D[0, :n0] = md
m = ORBmatcher(0.9, True, max_batch=1, max_keypoints=cap)
m.set_frames(cam, kps[1:2], desc[1:2], counts[1:2], 1, cap)
match, nm = m.SearchByProjectionFrame(np.eye(4).reshape(1, 16), kps[0:1], counts[0:1], flags, X, D, cap, 15.0)
# the stream-pipelined front end, all three forms of the last-frame inputs (two lanes, one frame per chunk, two batches in flight)
from ceres_mono_orb_slam2_b200 import KP_DTYPE, TrackingFrontEnd
from ceres_mono_orb_slam2_b200.tracking import pack_last_points, pack_map_associations
fe = TrackingFrontEnd(cam, 300, 1.2, 8, 20, 7, max_width=W, max_height=H, lanes=2, chunk_frames=1)
T2 = np.tile(np.eye(4).reshape(1, 16), (2, 1))
lk2 = np.stack([kps[0], kps[0]]); lc2 = np.array([n0, n0], np.int32)
fl2_ = np.concatenate([flags, flags]); X2 = np.concatenate([X, X]); D2 = np.concatenate([D, D])
mk = lambda: (np.zeros((2, cap), KP_DTYPE), np.zeros((2, cap, 32), np.uint8), np.zeros(2, np.int32), np.full((2, cap), -1, np.int32), np.zeros(2, np.int32))
oa, ob, oc = mk(), mk(), mk()
ta = fe.submit(frames, T2, lk2, lc2, fl2_, X2, D2, 15.0, out=oa)
pts, pstart = pack_last_points(lk2, lc2, fl2_, X2, D2)
tb = fe.submit_points(frames, T2, pts, pstart, 15.0, out=ob)
fe.wait(ta); fe.wait(tb)
slots = np.full((2, cap), -1, np.int32); us = np.argwhere(fl2_ & 1); slots[us[:, 0], us[:, 1]] = np.arange(len(us), dtype=np.int32)
fe.map_reserve(len(us)); fe.map_update(X2[us[:, 0], us[:, 1]][: len(us) // 2], D2[us[:, 0], us[:, 1]][: len(us) // 2])
rest = np.arange(len(us) // 2, len(us), dtype=np.int32)
fe.map_update(X2[us[rest, 0], us[rest, 1]], D2[us[rest, 0], us[rest, 1]], slots=rest)
assoc, astart = pack_map_associations(lk2, lc2, fl2_, slots)
fe.wait(fe.submit_map(frames, T2, assoc, astart, 15.0, out=oc))
assert np.array_equal(oa[3], ob[3]) and np.array_equal(oa[3], oc[3]) and int(oa[4][1]) == int(nm[0])
K4 = np.array(synth.KITTI_K, np.float32)
G = synth.make_ba_problem(6, 120, 4, seed=11, n_fixed_extra=2)
fl2 = G["fixed"].copy(); fl2[6:] |= 2
opt = CeresOptimizer(max_cams=128, max_points=4000, max_obs=20000, max_pose_batch=1, max_pose_corr=200)
opt.LocalBundleAdjustment(G["poses"], fl2, G["points"], G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"], K4)
P = synth.make_pose_problem(150, seed=8)
opt.PoseOptimization(P["pose"][None], P["Xw"][None], P["uv"][None], P["inv_sigma2"][None], K4)
G2 = synth.make_ba_problem_fast(64, 64 * 40, 5, seed=3, window=4)     # banded: the nested-dissection solver
opt.BundleAdjustment(G2["poses"], G2["fixed"], G2["points"], G2["obs_cam"], G2["obs_pt"], G2["uv"], G2["inv_sigma2"], K4, n_iterations=2)
E = synth.make_essential_graph_problem(40, seed=12, n_group=3, covis=(2, 3), n_points=10)    # band + border nested dissection
r = opt.OptimizeEssentialGraph(E["Scw"], E["kf_flags"], E["Snc"], E["edge_j"], E["edge_i"], E["edge_kind"], E["Xw"], E["ref_kf"],
                               max_iterations=2)
assert opt.essential_graph_plan()["nested_dissection"] == 1
print("SANITIZER_WORKLOAD_DONE", int(counts.sum()), int(nm[0]))
''' % ROOT


@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_compute_sanitizer_is_clean(tool, tmp_path):
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    script = tmp_path / "workload.py"
    script.write_text(SNIPPET)
    r = subprocess.run([exe, "--tool", tool, "--error-exitcode", "7", "--launch-timeout", "120", sys.executable, str(script)],
                       capture_output=True, text=True, timeout=1500)
    tail = (r.stdout + r.stderr)[-4000:]
    assert "SANITIZER_WORKLOAD_DONE" in r.stdout, tail
    assert r.returncode == 0, tail
    out = r.stdout + r.stderr
    clean = "ERROR SUMMARY: 0 errors" in out if tool == "memcheck" else "RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings)" in out
    assert clean, tail

"""BASELINE.json's full sizes through the C ABI, checked by properties that do not need the CPU oracle at that size (plus the
oracle on a sample): configs[1] = 64 frames of 1241x376 / 2000 features, configs[4] = 1000 keyframes x 100 000 points x
500 000 observations."""
import numpy as np
import pytest

from ceres_mono_orb_slam2_b200 import CeresOptimizer, ORBextractor, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def test_orb_full_batch_properties():
    """One 64-frame launch: every frame's result is independent of its batch position (three sampled frames equal their
    single-frame extraction and the CPU oracle bit for bit), repeated runs are identical, counts stay inside the quota band
    (sum of per-level quotas .. + the quadtree's overshoot), keypoints lie inside the image minus the 19-px edge at their level,
    levels are concatenated in order."""
    B = 64
    frames = np.stack([synth.make_image(1241, 376, 1000 + f) for f in range(B)])
    ext = ORBextractor(2000, 1.2, 8, 20, 7, max_width=1241, max_height=376, max_batch=B)
    kps, desc, counts = ext.extract_batch(frames)
    k2, d2, c2 = ext.extract_batch(frames)
    assert np.array_equal(counts, c2) and np.array_equal(kps, k2) and np.array_equal(desc, d2)
    oracle = po.OrbOracle(2000, 1.2, 8, 20, 7)
    quota = int(oracle.quota.sum())
    assert (counts >= quota - 8).all() and (counts <= quota + 24).all(), (counts.min(), counts.max(), quota)
    for f in (0, 29, 63):
        n = int(counts[f])
        ks, ds, cs = ext.extract_batch(frames[f:f + 1])
        assert int(cs[0]) == n and np.array_equal(ks[0, :n], kps[f, :n]) and np.array_equal(ds[0, :n], desc[f, :n])
        ok, od = oracle.extract(frames[f])
        assert len(ok) == n and np.array_equal(kps[f, :n], ok) and np.array_equal(desc[f, :n], od)
    sf = oracle.scale_factors
    for f in range(B):
        k = kps[f, :int(counts[f])]
        assert (np.diff(k["octave"]) >= 0).all() and k["octave"].min() == 0 and k["octave"].max() == 7
        x = k["x"] / sf[k["octave"]]; y = k["y"] / sf[k["octave"]]
        for l in range(8):
            lw, lh = oracle.level_size(l)
            m = k["octave"] == l
            assert (x[m] > 18.5).all() and (x[m] < lw - 18.5).all() and (y[m] > 18.5).all() and (y[m] < lh - 18.5).all()
        assert (k["angle"] >= 0).all() and (k["angle"] < 360).all() and (k["response"] >= 7).all()


def test_global_bundle_adjustment_full_size_properties(monkeypatch):
    """configs[4] at full size: the solve is deterministic (two runs bit-identical), the cost never increases along the
    accepted steps and falls by an order of magnitude, the constant keyframe does not move, and two different linear solvers —
    the banded persistent-CTA Cholesky and the blocked envelope Cholesky (CMOS_BA_NO_BAND=1) — give the same trajectory."""
    G = synth.make_ba_problem_fast(1000, 100000, 5, seed=5)
    K4 = np.array(synth.KITTI_K, np.float32)
    assert len(G["obs_cam"]) == 500000

    def solve():
        opt = CeresOptimizer(max_cams=1000, max_points=100000, max_obs=500000)
        cams, pts, s = opt.BundleAdjustment(G["poses"], G["fixed"], G["points"], G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"],
                                            K4, n_iterations=6, is_robust=True)
        tr = opt.trace(0, int(s["iterations"]) + 1)
        opt.close()
        return cams, pts, s, tr

    c1, p1, s1, t1 = solve()
    c2, p2, s2, t2 = solve()
    assert np.array_equal(c1, c2) and np.array_equal(p1, p2) and np.array_equal(t1, t2)
    assert s1["iterations"] == 6 and s1["successful_steps"] >= 4
    cost = t1[:, 0]
    assert np.all(np.diff(cost) <= 1e-9 * cost[0]) and s1["final_cost"] < 0.1 * s1["initial_cost"]
    fixed = G["fixed"].astype(bool)
    assert np.array_equal(c1[fixed], G["poses"][fixed])
    assert np.abs(np.linalg.norm(c1[:, 3:], axis=1) - 1).max() < 1e-9          # quaternions stay on the manifold
    monkeypatch.setenv("CMOS_BA_NO_BAND", "1")
    c3, p3, s3, t3 = solve()
    assert (s3["iterations"], s3["successful_steps"]) == (s1["iterations"], s1["successful_steps"])
    assert np.abs(t3[:, 0] / t1[:, 0] - 1).max() < 1e-9
    assert np.abs(c3 - c1).max() < 1e-7 and np.abs(p3 - p1).max() < 1e-5 * max(1.0, np.abs(p1).max())


def test_global_bundle_adjustment_full_size_matches_oracle():
    """configs[4] at FULL size (1000 keyframes x 100 000 points x 500 000 observations) against the CPU oracle, whose
    reduced-camera solve factorises inside the row envelope (what CHOLMOD does for Ceres, CeresOptimizer.cc:178-187): same
    iteration count and accept / reject sequence, per-iteration cost to 1e-9, poses and points to 1e-7 (north_star asks 1e-4).
    The GPU solve is the nested-dissection (block cyclic reduction) Cholesky of csrc/band_cr.cuh."""
    from oracle import pyoracle as po
    G = synth.make_ba_problem_fast(1000, 100000, 5, seed=5)
    K4 = np.array(synth.KITTI_K, np.float32)
    iters = 5
    opt = CeresOptimizer(max_cams=1000, max_points=100000, max_obs=500000)
    cams, pts, s = opt.BundleAdjustment(G["poses"], G["fixed"], G["points"], G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"],
                                        K4, n_iterations=iters, is_robust=True)
    tr = opt.trace(0, int(s["iterations"]) + 1)
    opt.close()
    oc, op_, os_, otr = po.ba_global(G["poses"], G["fixed"], G["points"], G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"],
                                     G["K"], iters, True)
    assert s["iterations"] == os_["iterations"] == iters and s["successful_steps"] == os_["successful_steps"]
    n = os_["iterations"] + 1
    assert np.abs(tr[:n, 0] / otr[:n, 0] - 1).max() < 1e-9, "per-iteration cost"
    assert np.abs(cams - oc).max() / max(1.0, np.abs(oc).max()) < 1e-7
    assert np.abs(pts - op_).max() / max(1.0, np.abs(op_).max()) < 1e-7


@pytest.mark.parametrize("n_cams,window", [(260, 10), (150, 4), (500, 12)])
def test_global_bundle_adjustment_cyclic_reduction_sizes(n_cams, window, monkeypatch):
    """Nested-dissection solver at other band widths / node counts (W = 20, 8, 24 blocks; padded last node; node counts that
    are not powers of two) against the oracle, and against the serial one-CTA band walk (CMOS_BA_BAND_SERIAL=1)."""
    from oracle import pyoracle as po
    G = synth.make_ba_problem_fast(n_cams, n_cams * 60, 5, seed=n_cams, window=window)
    K4 = np.array(synth.KITTI_K, np.float32)

    def solve():
        opt = CeresOptimizer(max_cams=n_cams, max_points=n_cams * 60, max_obs=n_cams * 300)
        cams, pts, s = opt.BundleAdjustment(G["poses"], G["fixed"], G["points"], G["obs_cam"], G["obs_pt"], G["uv"],
                                            G["inv_sigma2"], K4, n_iterations=6, is_robust=True)
        tr = opt.trace(0, int(s["iterations"]) + 1)
        opt.close()
        return cams, pts, s, tr

    cams, pts, s, tr = solve()
    oc, op_, os_, otr = po.ba_global(G["poses"], G["fixed"], G["points"], G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"],
                                     G["K"], 6, True)
    assert s["iterations"] == os_["iterations"] and s["successful_steps"] == os_["successful_steps"]
    n = os_["iterations"] + 1
    assert np.abs(tr[:n, 0] / otr[:n, 0] - 1).max() < 1e-9
    assert np.abs(cams - oc).max() < 1e-7 and np.abs(pts - op_).max() < 1e-6
    monkeypatch.setenv("CMOS_BA_BAND_SERIAL", "1")
    c2, p2, s2, t2 = solve()
    assert (s2["iterations"], s2["successful_steps"]) == (s["iterations"], s["successful_steps"])
    assert np.abs(c2 - cams).max() < 1e-8 and np.abs(p2 - pts).max() < 1e-8 * max(1.0, np.abs(pts).max())


def test_extract_and_search_full_batch_properties():
    """configs[1] end to end at full size (64 ring-paired frames, th = 15): every match index is valid and used at most once
    per frame, the number of matched keypoints is the match count minus the re-assigned ones, matched descriptors are within TH_HIGH = 100 of
    their map point's, the host-buffer pipeline (cmos_track_frames) returns the same matches as the unfused device calls, and
    three sampled frames equal the CPU oracle index for index."""
    import bench
    from ceres_mono_orb_slam2_b200 import Camera, ORBmatcher, TrackingFrontEnd
    B, W, H = 64, bench.W, bench.H
    frames, offs = bench.make_batch(B, 1000)
    ext = ORBextractor(2000, 1.2, 8, 20, 7, max_width=W, max_height=H, max_batch=B)
    kps, desc, counts = ext.extract_batch(frames)
    cap = ext.capacity
    lk, lcounts, flags, xw, mdesc, T = bench.make_last_views(kps, desc, counts, offs, cap, seed=5000)
    cam = Camera.create(W, H, synth.KITTI_K, ext.GetScaleFactors(), 1.2)
    m = ORBmatcher(0.9, True, max_batch=B, max_keypoints=cap)
    m.set_frames(cam, kps, desc, counts, B, cap)
    match, nm = m.SearchByProjectionFrame(T, lk, lcounts, flags, xw, mdesc, cap, 15.0)
    assert int(nm.sum()) > 500 * B
    popcnt = np.unpackbits(np.arange(256, dtype=np.uint8)[:, None], axis=1).sum(1)
    for f in range(B):
        n = int(counts[f]); mf = match[f, :n]
        hit = mf >= 0
        # nmatches counts assignments: a keypoint holding a point without observations may be re-assigned (ORBmatcher.cc:1219-1221)
        assert int(nm[f]) - 64 <= int(hit.sum()) <= int(nm[f]) and (match[f, n:] == -1).all()
        assert (mf[hit] < lcounts[f]).all() and len(np.unique(mf[hit])) == int(hit.sum())      # one keypoint per map point
        assert flags[f][mf[hit]].all()                                                          # only valid map points
        d = popcnt[desc[f, :n][hit] ^ mdesc[f][mf[hit]]].sum(1)
        assert (d <= 100).all()
    b = cam.bounds6()
    for f in (0, 31, 63):
        n = int(counts[f]); nl = int(lcounts[f])
        gs, gi = po.build_grid(kps[f, :n], b)
        om, onm, _ = po.search_by_projection_frame(kps[f, :n], desc[f, :n], gs, gi, b, cam.K4(), ext.GetScaleFactors(), T[f],
                                                   lk[f, :nl], flags[f, :nl], xw[f, :nl], mdesc[f, :nl], 15.0, True)
        assert onm == nm[f] and np.array_equal(match[f, :n], om)
    front = TrackingFrontEnd(cam, 2000, 1.2, 8, 20, 7, max_width=W, max_height=H, lanes=4, chunk_frames=16)
    k2, d2, c2, m2, nm2 = front.track(frames, T, lk, lcounts, flags, xw, mdesc, 15.0)
    assert np.array_equal(c2, counts) and np.array_equal(nm2, nm)
    for f in range(B):
        n = int(counts[f])
        assert np.array_equal(m2[f, :n], match[f, :n]) and np.array_equal(d2[f, :n], desc[f, :n])

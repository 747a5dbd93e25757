"""GPU parity tests of the matcher: Frame grid, both SearchByProjection variants and isInFrustum through the
C ABI against the CPU oracle, index-exact.  Covers batches, pre-claimed keypoints, points without
observations (re-claimable), the list-overflow path (many equally good candidates) and empty inputs."""
import numpy as np
import pytest

from ceres_mono_orb_slam2_b200 import KP_DTYPE, Camera, ORBmatcher, synth
from oracle import pyoracle as po
from tests.matcher_scenarios import camera_arrays, extract_sequence, identity_T, points_view

pytestmark = pytest.mark.gpu

W, H, K = 1241, 376, synth.KITTI_K


@pytest.fixture(scope="module")
def seq():
    frames, offs, ext, o = extract_sequence(W, H, 5, 2000, seed=77)
    return offs, ext, o


def _pack(ext, stride):
    B = len(ext)
    kps = np.zeros((B, stride), KP_DTYPE); desc = np.zeros((B, stride, 32), np.uint8); counts = np.zeros(B, np.int32)
    for f, (k, d) in enumerate(ext):
        kps[f, :len(k)] = k; desc[f, :len(k)] = d; counts[f] = len(k)
    return kps, desc, counts


def test_grid_matches_oracle(seq):
    _, ext, o = seq
    cam = Camera.create(W, H, K, o.scale_factors)
    stride = 2024
    kps, desc, counts = _pack(ext, stride)
    m = ORBmatcher(0.9, True, max_batch=len(ext), max_keypoints=stride)
    m.set_frames(cam, kps, desc, counts, len(ext), stride)
    assert np.array_equal(cam.bounds6(), camera_arrays(W, H, K, o.scale_factors)[0])
    for f in range(len(ext)):
        gs, gi = m.debug_grid(f)
        ogs, ogi = po.build_grid(ext[f][0], cam.bounds6())
        assert np.array_equal(gs, ogs) and np.array_equal(gi, ogi), f"frame {f}"


@pytest.mark.parametrize("check_ori,th", [(True, 15.0), (False, 15.0), (True, 30.0)])
def test_search_by_projection_frame_batch(seq, check_ori, th):
    offs, ext, o = seq
    cam = Camera.create(W, H, K, o.scale_factors)
    b, K4, sf = cam.bounds6(), cam.K4(), o.scale_factors
    stride = 2024
    B = len(ext) - 1
    cur = ext[1:]; last = ext[:-1]
    kps, desc, counts = _pack(cur, stride)
    lkps, ldesc_kp, lcounts = _pack(last, stride)
    flags = np.zeros((B, stride), np.uint8); xw = np.zeros((B, stride, 3)); mdesc = np.zeros((B, stride, 32), np.uint8)
    T = np.tile(identity_T(), (B, 1))
    claimed0 = np.zeros((B, stride), np.uint8)
    rng = np.random.default_rng(4)
    for f in range(B):
        shift = (offs[f] - offs[f + 1]).astype(np.float64)
        fl, x, md = synth.make_last_frame_view(last[f][0], last[f][1], shift, seed=100 + f, K=K)
        n = len(fl)
        flags[f, :n] = fl; xw[f, :n] = x; mdesc[f, :n] = md
        claimed0[f, :counts[f]] = rng.random(counts[f]) < 0.05      # some keypoints already hold a point
        # a small camera motion so the fp64 projection is not the identity
        T[f] = np.array([[1, 0.001 * f, 0, 0.01], [-0.001 * f, 1, 0, -0.02], [0, 0, 1, 0.03], [0, 0, 0, 1]]).reshape(-1)
    m = ORBmatcher(0.9, check_ori, max_batch=B, max_keypoints=stride)
    m.set_frames(cam, kps, desc, counts, B, stride)
    claimed = claimed0.copy()
    match, nm = m.SearchByProjectionFrame(T, lkps, lcounts, flags, xw, mdesc, stride, th, claimed=claimed)
    total = 0
    for f in range(B):
        ck, cd = cur[f]
        gs, gi = po.build_grid(ck, b)
        om, onm, ocl = po.search_by_projection_frame(ck, cd, gs, gi, b, K4, sf, T[f], last[f][0], flags[f, :lcounts[f]],
                                                     xw[f, :lcounts[f]], mdesc[f, :lcounts[f]], th, check_ori,
                                                     claimed=claimed0[f, :counts[f]].copy())
        assert nm[f] == onm, f"frame {f}: nmatches {nm[f]} vs {onm}"
        bad = np.nonzero(match[f, :counts[f]] != om)[0]
        assert bad.size == 0, f"frame {f}: {bad.size} keypoints differ, first {bad[:5]}"
        assert (match[f, counts[f]:] == -1).all()
        assert np.array_equal(claimed[f, :counts[f]], ocl)
        total += onm
    assert total > 300 * B


def test_search_frame_overflow_and_reclaim():
    """All descriptors identical: every window candidate ties at distance 0, lists overflow, first candidate in
    grid order must win; points without observations do not block later queries."""
    rng = np.random.default_rng(8)
    n = 1500
    kps = np.zeros(n, KP_DTYPE)
    kps["x"] = rng.uniform(20, W - 20, n).astype(np.float32); kps["y"] = rng.uniform(20, H - 20, n).astype(np.float32)
    kps["octave"] = rng.integers(0, 3, n); kps["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    desc = np.zeros((n, 32), np.uint8)
    o = po.OrbOracle(2000, 1.2, 8, 20, 7)
    cam = Camera.create(W, H, K, o.scale_factors)
    flags = np.where(rng.random(n) < 0.5, 3, 1).astype(np.uint8)       # half the points have no observations
    xw = np.zeros((n, 3)); z = rng.uniform(5, 30, n)
    fx, fy, cx, cy = [float(np.float32(v)) for v in K]
    xw[:, 0] = (kps["x"] - cx) / fx * z; xw[:, 1] = (kps["y"] - cy) / fy * z; xw[:, 2] = z
    T = identity_T()[None]
    m = ORBmatcher(0.9, True, max_batch=1, max_keypoints=n)
    m.set_frames(cam, kps[None], desc[None], np.array([n], np.int32), 1, n)
    match, nm = m.SearchByProjectionFrame(T, kps[None], np.array([n], np.int32), flags[None], xw[None], desc[None], n, 30.0)
    gs, gi = po.build_grid(kps, cam.bounds6())
    om, onm, _ = po.search_by_projection_frame(kps, desc, gs, gi, cam.bounds6(), cam.K4(), o.scale_factors, T[0], kps,
                                               flags, xw, desc, 30.0, True)
    assert nm[0] == onm and np.array_equal(match[0], om)


@pytest.mark.parametrize("th,palette,obs_frac,n", [(6.0, 3, 0.9, 1800), (10.0, 40, 0.5, 1800), (15.0, 8, 1.0, 1800),
                                                   (8.0, 60, 0.9, 4500)])     # 4500 > what the list staging fits
def test_search_frame_contention_chains(th, palette, obs_frac, n):
    """Clustered keypoints with descriptors drawn from a small palette: many queries want the same candidates, so
    the claim chains of the greedy assignment are long (the parallel fixed point needs many rounds) while most
    lists stay within the list capacity."""
    rng = np.random.default_rng(int(th) * 100 + palette)
    centres = np.stack([rng.uniform(60, W - 60, 40), rng.uniform(40, H - 40, 40)], 1)
    which = rng.integers(0, 40, n)
    kps = np.zeros(n, KP_DTYPE)
    kps["x"] = np.clip(centres[which, 0] + rng.normal(0, 14, n), 1, W - 2).astype(np.float32)
    kps["y"] = np.clip(centres[which, 1] + rng.normal(0, 9, n), 1, H - 2).astype(np.float32)
    kps["octave"] = rng.integers(0, 2, n); kps["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    pal = rng.integers(0, 256, (palette, 32)).astype(np.uint8)
    desc = pal[rng.integers(0, palette, n)].copy()
    flip = rng.random(n) < 0.5
    desc[flip, rng.integers(0, 32, flip.sum())] ^= np.uint8(1) << rng.integers(0, 8, flip.sum()).astype(np.uint8)
    o = po.OrbOracle(2000, 1.2, 8, 20, 7)
    cam = Camera.create(W, H, K, o.scale_factors)
    # last frame = a permutation of the same keypoints, slightly moved
    perm = rng.permutation(n)
    lk = kps[perm].copy()
    ldesc = desc[perm].copy()
    flags = (1 | ((rng.random(n) < obs_frac).astype(np.uint8) << 1)).astype(np.uint8)
    flags[rng.random(n) < 0.1] &= 2
    z = rng.uniform(5, 30, n)
    fx, fy, cx, cy = [float(np.float32(v)) for v in K]
    u = lk["x"] + rng.normal(0, 1.5, n); v = lk["y"] + rng.normal(0, 1.5, n)
    xw = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    T = identity_T()[None]
    claimed0 = (rng.random(n) < 0.03).astype(np.uint8)
    m = ORBmatcher(0.9, True, max_batch=1, max_keypoints=n)
    m.set_frames(cam, kps[None], desc[None], np.array([n], np.int32), 1, n)
    claimed = claimed0[None].copy()
    match, nm = m.SearchByProjectionFrame(T, lk[None], np.array([n], np.int32), flags[None], xw[None], ldesc[None], n, th,
                                          claimed=claimed)
    gs, gi = po.build_grid(kps, cam.bounds6())
    om, onm, ocl = po.search_by_projection_frame(kps, desc, gs, gi, cam.bounds6(), cam.K4(), o.scale_factors, T[0], lk,
                                                 flags, xw, ldesc, th, True, claimed=claimed0.copy())
    assert nm[0] == onm
    assert np.array_equal(match[0], om), f"{(match[0] != om).sum()} keypoints differ"
    assert np.array_equal(claimed[0], ocl)
    assert onm > 200


@pytest.mark.parametrize("th,ratio", [(1.0, 0.8), (5.0, 0.8), (3.0, 0.6)])
def test_search_by_projection_points_batch(seq, th, ratio):
    offs, ext, o = seq
    cam = Camera.create(W, H, K, o.scale_factors)
    b, sf = cam.bounds6(), o.scale_factors
    stride, pstride = 2024, 2100
    B = len(ext) - 1
    cur = ext[1:]; last = ext[:-1]
    kps, desc, counts = _pack(cur, stride)
    npts = np.zeros(B, np.int32)
    in_view = np.zeros((B, pstride), np.uint8); level = np.zeros((B, pstride), np.int32)
    vcos = np.zeros((B, pstride), np.float32); proj = np.zeros((B, pstride, 2), np.float32)
    mdesc = np.zeros((B, pstride, 32), np.uint8); has_obs = np.zeros((B, pstride), np.uint8)
    views = []
    for f in range(B):
        shift = (offs[f] - offs[f + 1]).astype(np.float64)
        v = points_view(last[f][0], last[f][1], shift, seed=50 + f, th_noise=1.0 if th == 1.0 else 3.0)
        n = len(v[0]); npts[f] = n
        in_view[f, :n], level[f, :n], vcos[f, :n], proj[f, :n], mdesc[f, :n], has_obs[f, :n] = v
        views.append(v)
    m = ORBmatcher(ratio, True, max_batch=B, max_keypoints=stride, max_points=pstride)
    m.set_frames(cam, kps, desc, counts, B, stride)
    claimed = np.zeros((B, stride), np.uint8)
    assign, nm = m.SearchByProjectionPoints(npts, in_view, level, vcos, proj, mdesc, has_obs, pstride, th, claimed=claimed)
    total = 0
    for f in range(B):
        ck, cd = cur[f]
        gs, gi = po.build_grid(ck, b)
        oa, onm, ocl = po.search_by_projection_points(ck, cd, gs, gi, b, sf, *views[f], th, ratio)
        assert nm[f] == onm, f"frame {f}: {nm[f]} vs {onm}"
        bad = np.nonzero(assign[f, :counts[f]] != oa)[0]
        assert bad.size == 0, f"frame {f}: {bad.size} differ, first {bad[:5]}"
        assert np.array_equal(claimed[f, :counts[f]], ocl)
        total += onm
    assert total > 100 * B


def test_search_points_overflow_ties():
    rng = np.random.default_rng(12)
    n = 1200
    kps = np.zeros(n, KP_DTYPE)
    kps["x"] = rng.uniform(20, W - 20, n).astype(np.float32); kps["y"] = rng.uniform(20, H - 20, n).astype(np.float32)
    kps["octave"] = rng.integers(0, 2, n)
    desc = rng.integers(0, 2, (n, 32)).astype(np.uint8)          # distances are small and tie often
    o = po.OrbOracle(2000, 1.2, 8, 20, 7)
    cam = Camera.create(W, H, K, o.scale_factors)
    npnt = 900
    in_view = np.ones(npnt, np.uint8); level = rng.integers(0, 2, npnt).astype(np.int32)
    vcos = rng.uniform(0.99, 1, npnt).astype(np.float32)
    proj = np.stack([rng.uniform(0, W, npnt), rng.uniform(0, H, npnt)], 1).astype(np.float32)
    mdesc = rng.integers(0, 2, (npnt, 32)).astype(np.uint8); has_obs = (rng.random(npnt) < 0.7).astype(np.uint8)
    m = ORBmatcher(0.8, True, max_batch=1, max_keypoints=n, max_points=npnt)
    m.set_frames(cam, kps[None], desc[None], np.array([n], np.int32), 1, n)
    assign, nm = m.SearchByProjectionPoints(np.array([npnt], np.int32), in_view[None], level[None], vcos[None], proj[None],
                                            mdesc[None], has_obs[None], npnt, 25.0)
    gs, gi = po.build_grid(kps, cam.bounds6())
    oa, onm, _ = po.search_by_projection_points(kps, desc, gs, gi, cam.bounds6(), o.scale_factors, in_view, level, vcos,
                                                proj, mdesc, has_obs, 25.0, 0.8)
    assert nm[0] == onm and np.array_equal(assign[0], oa)


def test_search_points_large_local_map():
    """A local map of 9000 points next to 2000 keypoints: the candidate lists no longer fit the CTA's shared memory and live
    in HBM scratch (the drop-in used to throw above ~6000 points).  Same result as the oracle."""
    rng = np.random.default_rng(99)
    n, npnt = 2000, 9000
    kps = np.zeros(n, KP_DTYPE)
    kps["x"] = rng.uniform(20, W - 20, n).astype(np.float32); kps["y"] = rng.uniform(20, H - 20, n).astype(np.float32)
    kps["octave"] = rng.integers(0, 4, n)
    desc = rng.integers(0, 256, (n, 32)).astype(np.uint8)
    o = po.OrbOracle(2000, 1.2, 8, 20, 7)
    cam = Camera.create(W, H, K, o.scale_factors)
    src = rng.integers(0, n, npnt)                                   # every point is a noisy copy of some keypoint
    in_view = (rng.random(npnt) < 0.9).astype(np.uint8); level = kps["octave"][src].astype(np.int32)
    vcos = rng.uniform(0.99, 1, npnt).astype(np.float32)
    proj = np.stack([kps["x"][src] + rng.normal(0, 1.5, npnt), kps["y"][src] + rng.normal(0, 1.5, npnt)], 1).astype(np.float32)
    mdesc = desc[src].copy()
    mdesc[np.arange(npnt), rng.integers(0, 32, npnt)] ^= (1 << rng.integers(0, 8, npnt)).astype(np.uint8)
    has_obs = (rng.random(npnt) < 0.7).astype(np.uint8)
    m = ORBmatcher(0.8, True, max_batch=1, max_keypoints=n, max_points=npnt)
    m.set_frames(cam, kps[None], desc[None], np.array([n], np.int32), 1, n)
    assign, nm = m.SearchByProjectionPoints(np.array([npnt], np.int32), in_view[None], level[None], vcos[None], proj[None],
                                            mdesc[None], has_obs[None], npnt, 3.0)
    gs, gi = po.build_grid(kps, cam.bounds6())
    oa, onm, _ = po.search_by_projection_points(kps, desc, gs, gi, cam.bounds6(), o.scale_factors, in_view, level, vcos,
                                                proj, mdesc, has_obs, 3.0, 0.8)
    assert nm[0] == onm and np.array_equal(assign[0], oa) and onm > 1000


def test_empty_inputs():
    o = po.OrbOracle(2000, 1.2, 8, 20, 7)
    cam = Camera.create(W, H, K, o.scale_factors)
    n = 64
    kps = np.zeros((1, n), KP_DTYPE); desc = np.zeros((1, n, 32), np.uint8)
    m = ORBmatcher(0.9, True, max_batch=1, max_keypoints=n, max_points=n)
    m.set_frames(cam, kps, desc, np.zeros(1, np.int32), 1, n)             # frame without keypoints
    match, nm = m.SearchByProjectionFrame(identity_T()[None], kps, np.zeros(1, np.int32), np.zeros((1, n), np.uint8),
                                          np.zeros((1, n, 3)), desc, n, 15.0)
    assert nm[0] == 0 and (match == -1).all()
    assign, nm = m.SearchByProjectionPoints(np.zeros(1, np.int32), np.zeros((1, n), np.uint8), np.zeros((1, n), np.int32),
                                            np.zeros((1, n), np.float32), np.zeros((1, n, 2), np.float32), desc,
                                            np.zeros((1, n), np.uint8), n, 3.0)
    assert nm[0] == 0 and (assign == -1).all()


def test_is_in_frustum(seq):
    _, ext, o = seq
    cam = Camera.create(W, H, K, o.scale_factors)
    rng = np.random.default_rng(3)
    n = 3000
    pose7 = np.concatenate([rng.normal(0, 0.2, 3), synth.quat_from_rotvec(rng.normal(0, 0.05, 3))])
    R = synth.quat_to_R(pose7[3:]); t = pose7[:3]; Ow = -R.T @ t
    pose15 = np.concatenate([R.reshape(-1), t, Ow])
    xw = rng.uniform([-30, -10, -5], [30, 10, 60], (n, 3))
    normal = (xw - Ow) / np.linalg.norm(xw - Ow, axis=1, keepdims=True) + rng.normal(0, 0.5, (n, 3))
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    d = np.linalg.norm(xw - Ow, axis=1)
    min_d = (d * rng.uniform(0.3, 1.3, n)).astype(np.float32); max_d = (min_d * rng.uniform(1.5, 4.0, n)).astype(np.float32)
    m = ORBmatcher(0.8, True, max_batch=1, max_keypoints=64, max_points=n)
    iv, pj, lv, vc = m.IsInFrustum(cam, pose15[None], 0.5, np.array([n], np.int32), xw[None], normal[None], min_d[None],
                                   max_d[None], n, 1)
    oiv, opj, olv, ovc = po.is_in_frustum(pose15, cam.K4(), cam.bounds6()[:4], cam.log_scale_factor, 8, 0.5, xw, normal,
                                          min_d, max_d)
    assert 100 < oiv.sum() < n
    assert np.array_equal(iv[0], oiv)
    sel = oiv.astype(bool)
    assert np.array_equal(pj[0][sel], opj[sel]) and np.array_equal(lv[0][sel], olv[sel]) and np.array_equal(vc[0][sel], ovc[sel])


@pytest.mark.parametrize("lanes,chunk", [(1, 8), (2, 2), (3, 1)])
def test_tracking_front_end_equals_unfused_calls(lanes, chunk):
    """cmos_track_frames (chunks pipelined over streams, host buffers) == extract + grid + SearchByProjection."""
    from ceres_mono_orb_slam2_b200 import ORBextractor, TrackingFrontEnd
    w, h, B = 640, 480, 5
    frames, offs = synth.make_sequence(w, h, B, 21, return_offsets=True)
    ext = ORBextractor(1000, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=B)
    kps, desc, counts = ext.extract_batch(frames)
    cap = ext.capacity
    cam = Camera.create(w, h, synth.TUM2_K, ext.GetScaleFactors())
    lk = np.zeros((B, cap), KP_DTYPE); lcounts = np.zeros(B, np.int32)
    flags = np.zeros((B, cap), np.uint8); xw = np.zeros((B, cap, 3)); mdesc = np.zeros((B, cap, 32), np.uint8)
    T = np.tile(identity_T(), (B, 1))
    for f in range(B):
        l = (f - 1) % B
        nl = int(counts[l])
        fl, x, md = synth.make_last_frame_view(kps[l, :nl], desc[l, :nl], (offs[l] - offs[f]).astype(float), 50 + f,
                                               K=synth.TUM2_K)
        lk[f, :nl] = kps[l, :nl]; lcounts[f] = nl; flags[f, :nl] = fl; xw[f, :nl] = x; mdesc[f, :nl] = md
    m = ORBmatcher(0.9, True, max_batch=B, max_keypoints=cap)
    m.set_frames(cam, kps, desc, counts, B, cap)
    match, nm = m.SearchByProjectionFrame(T, lk, lcounts, flags, xw, mdesc, cap, 15.0)
    fe = TrackingFrontEnd(cam, 1000, 1.2, 8, 20, 7, max_width=w, max_height=h, lanes=lanes, chunk_frames=chunk)
    assert fe.capacity == cap
    for _ in range(2):      # the second pass reuses every lane's buffers
        k2, d2, c2, m2, n2 = fe.track(frames, T, lk, lcounts, flags, xw, mdesc, 15.0)
        assert np.array_equal(c2, counts) and np.array_equal(n2, nm)
        for f in range(B):
            n = int(counts[f])
            assert np.array_equal(k2[f, :n], kps[f, :n]) and np.array_equal(d2[f, :n], desc[f, :n])
            assert np.array_equal(m2[f, :n], match[f, :n])
    assert nm.sum() > 100 * B and fe.launch_count() > 0
    # asynchronous pair: three batches in flight (the same batch and two shifted copies), waited for out of order
    from ceres_mono_orb_slam2_b200 import CmosError
    outs, tickets, ins = [], [], []
    for s_ in range(3):
        idx = np.roll(np.arange(B), s_)
        a = tuple(np.ascontiguousarray(x[idx]) for x in (frames, T, lk, lcounts, flags, xw, mdesc))
        o = (np.zeros((B, cap), KP_DTYPE), np.zeros((B, cap, 32), np.uint8), np.zeros(B, np.int32),
             np.full((B, cap), -1, np.int32), np.zeros(B, np.int32))
        ins.append((a, idx)); outs.append(o)
        tickets.append(fe.submit(*a, 15.0, out=o))
    for s_ in (1, 0, 2):
        fe.wait(tickets[s_])
        idx = ins[s_][1]
        k2, d2, c2, m2, n2 = outs[s_]
        assert np.array_equal(c2, counts[idx]) and np.array_equal(n2, nm[idx])
        for j, f in enumerate(idx):
            n = int(counts[f])
            assert np.array_equal(k2[j, :n], kps[f, :n]) and np.array_equal(m2[j, :n], match[f, :n])
    with pytest.raises(CmosError):
        fe.wait(tickets[0])            # already waited for
    # compact last-frame inputs (cmos_track_submit_points): one record per usable map point, same results; a frame without
    # any record and records given in the middle of a batch are part of the case
    from ceres_mono_orb_slam2_b200.tracking import pack_last_points
    fl2 = flags.copy(); fl2[B // 2] = 0                       # one frame whose last frame carries no map point at all
    match2, nm2 = m.SearchByProjectionFrame(T, lk, lcounts, fl2, xw, mdesc, cap, 15.0)
    pts, pstart = pack_last_points(lk, lcounts, fl2, xw, mdesc)
    assert pstart[B // 2 + 1] == pstart[B // 2] and len(pts) == int((fl2 & 1).sum())
    for _ in range(2):
        o = (np.zeros((B, cap), KP_DTYPE), np.zeros((B, cap, 32), np.uint8), np.zeros(B, np.int32),
             np.full((B, cap), -1, np.int32), np.zeros(B, np.int32))
        fe.wait(fe.submit_points(frames, T, pts, pstart, 15.0, out=o))
        k2, d2, c2, m2, n2 = o
        assert np.array_equal(c2, counts) and np.array_equal(n2, nm2)
        for f in range(B):
            n = int(counts[f])
            assert np.array_equal(k2[f, :n], kps[f, :n]) and np.array_equal(m2[f, :n], match2[f, :n])
    assert nm2[B // 2] == 0 and nm2.sum() > 100 * (B - 1)
    # device-resident map points (cmos_track_map_reserve / _update + cmos_track_submit_map): 12-byte association records, position
    # and descriptor gathered from the table; slots are a permutation of the usable keypoints, the table is filled in two steps
    # (a run of consecutive slots, then a slot list after growing it); one record points outside the table
    from ceres_mono_orb_slam2_b200.tracking import pack_map_associations
    rng = np.random.default_rng(77)
    usable = np.argwhere(fl2 & 1)                               # (frame, keypoint) of every usable map point
    n_mp = len(usable)
    perm = rng.permutation(n_mp).astype(np.int32)
    slots = np.full((B, cap), -1, np.int32); slots[usable[:, 0], usable[:, 1]] = perm
    t_xw = np.zeros((n_mp, 3)); t_desc = np.zeros((n_mp, 32), np.uint8)
    t_xw[perm] = xw[usable[:, 0], usable[:, 1]]; t_desc[perm] = mdesc[usable[:, 0], usable[:, 1]]
    with pytest.raises(CmosError):
        fe.submit_map(frames, T, np.zeros(1, np.uint8), np.zeros(B + 1, np.int32), 15.0, out=o)    # no table yet
    half = n_mp // 2
    fe.map_reserve(half)
    fe.map_update(t_xw[:half], t_desc[:half])
    with pytest.raises(CmosError):
        fe.map_update(t_xw[:1], t_desc[:1], slots=np.array([half], np.int32))                     # outside the table
    fe.map_reserve(n_mp)                                        # grows, keeps the first half
    rest = rng.permutation(np.arange(half, n_mp)).astype(np.int32)
    fe.map_update(t_xw[rest], t_desc[rest], slots=rest)
    lost = usable[n_mp // 3]                                    # this keypoint's record gets a slot outside the table
    slots[lost[0], lost[1]] = n_mp + 5
    fl3 = fl2.copy(); fl3[lost[0], lost[1]] = 0
    match3, nm3 = m.SearchByProjectionFrame(T, lk, lcounts, fl3, xw, mdesc, cap, 15.0)
    assoc, astart = pack_map_associations(lk, lcounts, fl2, slots)
    assert assoc.nbytes == 12 * n_mp and np.array_equal(astart, pstart)
    for _ in range(2):
        o = (np.zeros((B, cap), KP_DTYPE), np.zeros((B, cap, 32), np.uint8), np.zeros(B, np.int32),
             np.full((B, cap), -1, np.int32), np.zeros(B, np.int32))
        fe.wait(fe.submit_map(frames, T, assoc, astart, 15.0, out=o))
        k2, d2, c2, m2, n2 = o
        assert np.array_equal(c2, counts) and np.array_equal(n2, nm3)
        for f in range(B):
            n = int(counts[f])
            assert np.array_equal(k2[f, :n], kps[f, :n]) and np.array_equal(m2[f, :n], match3[f, :n])
    # a moved map point: the table entry changes, the records stay
    mv = usable[n_mp // 2]
    xw4 = xw.copy(); xw4[mv[0], mv[1]] += np.array([0.4, -0.3, 0.2])
    fe.map_update(xw4[mv[0], mv[1]][None], mdesc[mv[0], mv[1]][None], slots=np.array([slots[mv[0], mv[1]]], np.int32))
    match4, nm4 = m.SearchByProjectionFrame(T, lk, lcounts, fl3, xw4, mdesc, cap, 15.0)
    fe.wait(fe.submit_map(frames, T, assoc, astart, 15.0, out=o))
    assert np.array_equal(o[4], nm4) and all(np.array_equal(o[3][f, :int(counts[f])], match4[f, :int(counts[f])]) for f in range(B))


def test_undistort_keypoints_and_distorted_bounds():
    """cmos_match_undistort_keypoints / cmos_camera_init_distorted == the oracle (itself bit-exact vs cv2.undistortPoints)."""
    rng = np.random.default_rng(9)
    K4 = np.array([520.908620, 521.007327, 325.141442, 249.701764], np.float32)
    dist = np.array([0.231222, -0.784899, -0.003257, -0.000105, 0.917205], np.float32)
    B, stride = 3, 1200
    kps = np.zeros((B, stride), KP_DTYPE)
    kps["x"] = rng.uniform(0, 640, (B, stride)).astype(np.float32); kps["y"] = rng.uniform(0, 480, (B, stride)).astype(np.float32)
    kps["octave"] = rng.integers(0, 8, (B, stride)); kps["angle"] = rng.uniform(0, 360, (B, stride)).astype(np.float32)
    kps["response"] = rng.uniform(0, 200, (B, stride)).astype(np.float32); kps["size"] = 31; kps["class_id"] = -1
    counts = np.array([stride, 700, 0], np.int32)
    m = ORBmatcher(0.9, True, max_batch=B, max_keypoints=stride)
    und = m.UndistortKeyPoints(K4, dist, kps, counts)
    for f in range(B):
        n = counts[f]
        xy = po.undistort_points(K4, dist, np.stack([kps["x"][f, :n], kps["y"][f, :n]], 1))
        assert np.array_equal(und["x"][f, :n], xy[:, 0]) and np.array_equal(und["y"][f, :n], xy[:, 1])
        for fld in ("size", "angle", "response", "octave", "class_id"):
            assert np.array_equal(und[fld][f, :n], kps[fld][f, :n])
    try:
        import cv2
        K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
        ref = cv2.undistortPoints(np.stack([kps["x"][0], kps["y"][0]], 1).reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2)
        assert np.array_equal(und["x"][0], ref[:, 0]) and np.array_equal(und["y"][0], ref[:, 1])
    except ImportError:
        pass
    # no distortion: plain copy
    same = m.UndistortKeyPoints(K4, np.zeros(5, np.float32), kps, counts)
    assert np.array_equal(same[0], kps[0])
    sf = po.OrbOracle(1000, 1.2, 8, 20, 7).scale_factors
    cam = Camera.create_distorted(640, 480, K4, dist, sf)
    assert np.array_equal(cam.bounds6()[:4], po.image_bounds(K4, dist, 640, 480))
    b = cam.bounds6()
    assert b[4] == np.float32(64) / np.float32(b[1] - b[0]) and b[5] == np.float32(48) / np.float32(b[3] - b[2])

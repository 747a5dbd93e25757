"""Shared scenario builders for the matcher tests (CPU oracle tests and GPU parity tests)."""
import numpy as np

from ceres_mono_orb_slam2_b200 import synth
from oracle import pyoracle as po


def extract_sequence(width, height, n_frames, nfeat, seed):
    frames, offs = synth.make_sequence(width, height, n_frames, seed, return_offsets=True)
    o = po.OrbOracle(nfeat, 1.2, 8, 20, 7)
    out = [o.extract(f) for f in frames]
    return frames, offs, out, o


def camera_arrays(width, height, K, scale_factors):
    """bounds6 / K4 exactly as cmos_camera_init computes them (undistorted camera)."""
    b = np.array([0.0, width, 0.0, height, np.float32(64) / np.float32(width), np.float32(48) / np.float32(height)],
                 np.float32)
    return b, np.array(K, np.float32), np.asarray(scale_factors, np.float32)


def identity_T():
    return np.eye(4, dtype=np.float64).reshape(-1)


def points_view(last_kps, last_desc, shift_xy, seed, n_levels=8, th_noise=1.0):
    """Inputs of SearchByProjection(F, points): one map point per last-frame keypoint."""
    rng = np.random.default_rng(seed)
    n = len(last_kps)
    in_view = (rng.random(n) < 0.9).astype(np.uint8)
    level = np.clip(last_kps["octave"] + rng.integers(-1, 2, n) * (rng.random(n) < 0.2), 0, n_levels - 1).astype(np.int32)
    view_cos = rng.uniform(0.99, 1.0, n).astype(np.float32)
    proj = np.stack([last_kps["x"] + shift_xy[0] + rng.normal(0, th_noise, n),
                     last_kps["y"] + shift_xy[1] + rng.normal(0, th_noise, n)], 1).astype(np.float32)
    desc = last_desc.copy()
    for _ in range(6):
        byte = rng.integers(0, 32, n); bit = rng.integers(0, 8, n)
        desc[np.arange(n), byte] ^= (1 << bit).astype(np.uint8)
    has_obs = (rng.random(n) < 0.95).astype(np.uint8)
    return in_view, level, view_cos, proj, desc, has_obs


# ---------------------------------------------------------------------------------------------------
# Two keyframes that see the same map points (for the relocalisation / loop / BoW / fuse / Sim3 searches)

def _rot(rv):
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.eye(3)
    k = rv / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def make_two_views(n=1200, seed=0, width=1241, height=376, K=synth.KITTI_K, n_levels=8, scale=1.2, n_extra=300,
                   flip_bits=12, n_nodes=60):
    """Returns a dict: view 1 / view 2 keypoints + descriptors, the map points (world = camera-1 frame shifted),
    poses, vocabulary node per feature.  View-2 keypoint j observes point p2[j] (or -1)."""
    from ceres_mono_orb_slam2_b200 import KP_DTYPE
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = [float(np.float32(v)) for v in K]
    sf = np.empty(n_levels, np.float32); sf[0] = 1.0
    for i in range(1, n_levels):
        sf[i] = np.float32(np.float64(sf[i - 1]) * np.float64(np.float32(scale)))
    # camera 1
    R1 = _rot(rng.normal(0, 0.02, 3)); t1 = rng.normal(0, 0.1, 3)
    R2 = _rot(rng.normal(0, 0.03, 3)) @ R1; t2 = t1 + np.array([0.3, 0.02, 0.1]) + rng.normal(0, 0.02, 3)
    k1 = np.zeros(n, KP_DTYPE)
    k1["x"] = rng.uniform(5, width - 5, n).astype(np.float32); k1["y"] = rng.uniform(5, height - 5, n).astype(np.float32)
    k1["octave"] = np.minimum(rng.geometric(0.35, n) - 1, n_levels - 1); k1["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    k1["size"] = 31 * sf[k1["octave"]]; k1["class_id"] = -1
    z = rng.uniform(4, 40, n)
    Xc1 = np.stack([(k1["x"] - cx) / fx * z, (k1["y"] - cy) / fy * z, z], 1)
    Xw = (Xc1 - t1) @ R1          # R1^T (Xc - t1)
    Ow1 = -R1.T @ t1; Ow2 = -R2.T @ t2
    d1 = np.linalg.norm(Xw - Ow1, axis=1)
    max_d = (d1 * sf[k1["octave"]]).astype(np.float32); min_d = (max_d / sf[-1]).astype(np.float32)
    normal = (Xw - Ow1) / d1[:, None] + rng.normal(0, 0.05, (n, 3))   # MapPoint::UpdateNormalAndDepth: camera -> point; normal /= np.linalg.norm(normal, axis=1)[:, None]
    desc1 = rng.integers(0, 256, (n, 32)).astype(np.uint8)
    mp_desc = desc1.copy()
    for _ in range(4):
        mp_desc[np.arange(n), rng.integers(0, 32, n)] ^= (1 << rng.integers(0, 8, n)).astype(np.uint8)
    # camera 2 observations
    Xc2 = Xw @ R2.T + t2
    u2 = fx * Xc2[:, 0] / Xc2[:, 2] + cx + rng.normal(0, 0.8, n); v2 = fy * Xc2[:, 1] / Xc2[:, 2] + cy + rng.normal(0, 0.8, n)
    seen = (Xc2[:, 2] > 0) & (u2 > 2) & (u2 < width - 2) & (v2 > 2) & (v2 < height - 2) & (rng.random(n) < 0.85)
    src = np.nonzero(seen)[0]
    n2 = len(src) + n_extra
    k2 = np.zeros(n2, KP_DTYPE); p2 = np.full(n2, -1, np.int64)
    k2["x"][:len(src)] = u2[src]; k2["y"][:len(src)] = v2[src]
    k2["octave"][:len(src)] = np.clip(k1["octave"][src] + rng.integers(-1, 2, len(src)) * (rng.random(len(src)) < 0.3), 0,
                                      n_levels - 1)
    k2["angle"][:len(src)] = (k1["angle"][src] + rng.normal(0, 4, len(src))) % 360
    p2[:len(src)] = src
    k2["x"][len(src):] = rng.uniform(5, width - 5, n_extra); k2["y"][len(src):] = rng.uniform(5, height - 5, n_extra)
    k2["octave"][len(src):] = np.minimum(rng.geometric(0.35, n_extra) - 1, n_levels - 1)
    k2["angle"][len(src):] = rng.uniform(0, 360, n_extra)
    k2["size"] = 31 * sf[k2["octave"]]; k2["class_id"] = -1
    desc2 = rng.integers(0, 256, (n2, 32)).astype(np.uint8)
    d = desc1[src].copy()
    nflip = rng.integers(0, flip_bits + 1, len(src))
    for b in range(flip_bits):
        on = nflip > b
        d[np.nonzero(on)[0], rng.integers(0, 32, on.sum())] ^= (1 << rng.integers(0, 8, on.sum())).astype(np.uint8)
    desc2[:len(src)] = d
    perm = rng.permutation(n2)
    k2, desc2, p2 = k2[perm], desc2[perm], p2[perm]
    # vocabulary nodes: matching features mostly share a node
    node1 = rng.integers(0, n_nodes, n) * 7 + 3
    node2 = np.where(p2 >= 0, node1[np.maximum(p2, 0)], rng.integers(0, n_nodes, n2) * 7 + 3)
    stray = rng.random(n2) < 0.1
    node2[stray] = rng.integers(0, n_nodes, stray.sum()) * 7 + 3
    return dict(width=width, height=height, K=K, sf=sf, scale=scale, n_levels=n_levels, k1=k1, desc1=desc1, k2=k2,
                desc2=desc2, p2=p2, Xw=Xw, normal=normal, min_d=min_d, max_d=max_d, mp_desc=mp_desc, R1=R1, t1=t1, R2=R2,
                t2=t2, Ow1=Ow1, Ow2=Ow2, node1=node1, node2=node2, rng=rng)


def pose15(R, t):
    return np.concatenate([R.reshape(-1), t, -R.T @ t])


def T44(R, t, s=1.0):
    T = np.eye(4); T[:3, :3] = s * R; T[:3, 3] = s * t if s != 1.0 else t
    return T


def make_map_observations(n_points=3000, n_keyframes=40, seed=0, max_obs=30):
    """CSR of observations for batched MapPoint maintenance: descriptors cluster around a per-point prototype (with a
    few outliers so medians differ), keyframe centres on a path, some points without observations."""
    rng = np.random.default_rng(seed)
    nobs = np.minimum(rng.geometric(0.15, n_points), max_obs)
    nobs[rng.random(n_points) < 0.03] = 0
    nobs[:3] = [1, 2, max_obs]
    start = np.concatenate([[0], np.cumsum(nobs)]).astype(np.int32)
    total = int(start[-1])
    proto = rng.integers(0, 256, (n_points, 32)).astype(np.uint8)
    pt = np.repeat(np.arange(n_points), nobs)
    desc = proto[pt].copy()
    for _ in range(10):
        on = rng.random(total) < 0.5
        desc[np.nonzero(on)[0], rng.integers(0, 32, on.sum())] ^= (1 << rng.integers(0, 8, on.sum())).astype(np.uint8)
    wild = rng.random(total) < 0.08
    desc[wild] = rng.integers(0, 256, (wild.sum(), 32)).astype(np.uint8)
    obs_kf = rng.integers(0, n_keyframes, total).astype(np.int32)
    Ow = np.stack([0.4 * np.arange(n_keyframes), rng.normal(0, 0.05, n_keyframes), rng.normal(0, 0.05, n_keyframes)], 1)
    pos = np.stack([rng.uniform(-5, 0.4 * n_keyframes + 5, n_points), rng.uniform(-3, 3, n_points), rng.uniform(4, 40, n_points)], 1)
    ref_kf = np.where(nobs > 0, obs_kf[np.minimum(start[:-1], max(total - 1, 0))], 0).astype(np.int32)
    ref_level = rng.integers(0, 8, n_points).astype(np.int32)
    return dict(start=start, desc=desc, obs_kf=obs_kf, Ow=Ow, pos=pos, ref_kf=ref_kf, ref_level=ref_level,
                normal0=rng.normal(0, 1, (n_points, 3)), min0=rng.uniform(1, 2, n_points).astype(np.float32),
                max0=rng.uniform(20, 30, n_points).astype(np.float32))

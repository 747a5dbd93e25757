"""Seeded synthetic inputs for the two hot paths.  numpy only.

Images: mid-grey canvas + random rectangles + small blobs + mild noise, with a flat band kept so that
empty FAST cells and the iniThFAST -> minThFAST fallback (ORBextractor.cc:808-816) are exercised.
Frame pairs: frame t+1 is frame t shifted by a few pixels, so SearchByProjection has true matches.
BA graphs: cameras on an arc / loop looking at a point cloud, float32 observations with octave-dependent
noise and gross outliers, float32 intrinsics widened to double as the reference does
(CeresOptimizer.cc:150-154).
"""
from __future__ import annotations

import numpy as np

KITTI_K = (718.856, 718.856, 607.1928, 185.2157)   # configs/KITTI00-02.yaml:8-11
TUM2_K = (520.908620, 521.007327, 325.141442, 249.701764)   # configs/TUM2.yaml:8-11


def make_image(width: int, height: int, seed: int, n_rect: int = 400, n_blob: int = 200,
               noise: int = 6) -> np.ndarray:
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 128, np.int32)
    flat_y0 = int(height * 0.86)          # bottom band stays flat (+ noise only)
    for _ in range(n_rect):
        w = int(rng.integers(8, 121)); h = int(rng.integers(8, 121))
        x = int(rng.integers(-w // 2, width)); y = int(rng.integers(-h // 2, flat_y0))
        v = int(rng.integers(0, 256))
        img[max(y, 0):min(y + h, flat_y0), max(x, 0):x + w] = v
    for _ in range(n_blob):
        r = int(rng.integers(3, 10))
        x = int(rng.integers(0, width)); y = int(rng.integers(0, flat_y0))
        v = int(rng.integers(0, 256))
        yy, xx = np.ogrid[-r:r + 1, -r:r + 1]
        m = (xx * xx + yy * yy) <= r * r
        y0, y1 = max(y - r, 0), min(y + r + 1, flat_y0)
        x0, x1 = max(x - r, 0), min(x + r + 1, width)
        if y1 <= y0 or x1 <= x0:
            continue
        sub = img[y0:y1, x0:x1]
        mm = m[y0 - (y - r):y1 - (y - r), x0 - (x - r):x1 - (x - r)]
        sub[mm] = v
    if noise > 0:
        img += rng.integers(-noise, noise + 1, size=img.shape)
    # a noise-free flat strip inside the band: cells here are empty even at minThFAST
    img[int(height * 0.93):, : width // 2] = 128
    return np.clip(img, 0, 255).astype(np.uint8)


def make_sequence(width: int, height: int, n_frames: int, seed: int, max_shift: int = 6,
                  return_offsets: bool = False):
    """n_frames images; frame t+1 is a shifted crop of the same scene as frame t plus fresh noise.
    With return_offsets also returns the crop origin (ox, oy) of every frame."""
    rng = np.random.default_rng(seed)
    pad = max_shift * min(n_frames, 12) + 8
    scene = make_image(width + 2 * pad, height + 2 * pad, seed, n_rect=int(400 * (1 + 2 * pad / width) ** 2),
                       n_blob=int(200 * (1 + 2 * pad / width) ** 2), noise=0).astype(np.int32)
    out = np.empty((n_frames, height, width), np.uint8)
    offs = np.zeros((n_frames, 2), np.int32)
    ox, oy = pad, pad
    for t in range(n_frames):
        crop = scene[oy:oy + height, ox:ox + width] + rng.integers(-4, 5, size=(height, width))
        out[t] = np.clip(crop, 0, 255).astype(np.uint8)
        offs[t] = (ox, oy)
        ox += int(rng.integers(-max_shift, max_shift + 1))
        oy += int(rng.integers(-max_shift // 2, max_shift // 2 + 1))
        ox = min(max(ox, 0), 2 * pad); oy = min(max(oy, 0), 2 * pad)
    return (out, offs) if return_offsets else out


def make_last_frame_view(last_kps: np.ndarray, last_desc: np.ndarray, shift_xy, seed: int, K=KITTI_K,
                         valid_frac: float = 0.85, obs_frac: float = 0.95, flip_bits: int = 6):
    """Map points for SearchByProjection(cur, last): every last-frame keypoint gets a 3D point at a random
    depth that, under the identity current pose, projects to the keypoint moved by `shift_xy` (+ sub-pixel
    noise); the map-point descriptor is the keypoint's with a few bits flipped.
    Returns flags (bit0 valid, bit1 has observations), Xw [n,3] float64, descriptors [n,32]."""
    rng = np.random.default_rng(seed)
    n = len(last_kps)
    fx, fy, cx, cy = np.array(K, np.float32).astype(np.float64)
    z = rng.uniform(5, 50, n)
    u = last_kps["x"].astype(np.float64) + shift_xy[0] + rng.normal(0, 0.7, n)
    v = last_kps["y"].astype(np.float64) + shift_xy[1] + rng.normal(0, 0.7, n)
    xw = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    flags = (rng.random(n) < valid_frac).astype(np.uint8)
    flags |= ((rng.random(n) < obs_frac).astype(np.uint8) << 1)
    desc = last_desc.copy()
    for _ in range(flip_bits):
        byte = rng.integers(0, 32, n); bit = rng.integers(0, 8, n)
        desc[np.arange(n), byte] ^= (1 << bit).astype(np.uint8)
    return flags, xw, desc


# ------------------------------------------------------------------------------------------------
# SE3 helpers (camera-from-world, q stored x,y,z,w as MatEigenConverter.cc:68-77 does)

def quat_from_rotvec(rv: np.ndarray) -> np.ndarray:
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.array([0.0, 0.0, 0.0, 1.0])
    ax = rv / th
    return np.concatenate([ax * np.sin(th / 2), [np.cos(th / 2)]])


def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz])


def quat_to_R(q: np.ndarray) -> np.ndarray:
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def project(pose7: np.ndarray, X: np.ndarray, K) -> tuple[np.ndarray, np.ndarray]:
    R = quat_to_R(pose7[3:7]); pc = X @ R.T + pose7[:3]
    fx, fy, cx, cy = K
    uv = np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], 1)
    return uv, pc[:, 2]


def _octaves(rng, n, n_levels=8):
    p = 0.8 ** np.arange(n_levels); p /= p.sum()
    return rng.choice(n_levels, size=n, p=p).astype(np.int32)


def inv_sigma2_table(n_levels=8, scale=1.2) -> np.ndarray:
    """float32 table computed as ORBextractor.cc:419-431 does (float products)."""
    sf = np.empty(n_levels, np.float32); sf[0] = 1.0
    s = np.float64(np.float32(scale))
    for i in range(1, n_levels):
        sf[i] = np.float32(np.float64(sf[i - 1]) * s)
    sig = (sf * sf).astype(np.float32)
    return (np.float32(1.0) / sig).astype(np.float32)


def make_pose_problem(n_points: int = 1500, seed: int = 3, K=KITTI_K, outlier_frac: float = 0.1):
    """C3: one frame x n 3D-2D correspondences.  Returns dict of arrays (see keys)."""
    rng = np.random.default_rng(seed)
    Kf = np.array(K, np.float32).astype(np.float64)
    fx, fy, cx, cy = Kf
    true_pose = np.concatenate([rng.normal(0, 0.3, 3), quat_from_rotvec(rng.normal(0, 0.05, 3))])
    # points in the frustum of the true pose
    z = rng.uniform(2, 40, n_points)
    u = rng.uniform(20, 1221, n_points); v = rng.uniform(20, 356, n_points)
    pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1)
    R = quat_to_R(true_pose[3:]); Xw = (pc - true_pose[:3]) @ R      # R^T (pc - t)
    octv = _octaves(rng, n_points)
    inv_s2 = inv_sigma2_table()[octv]
    sigma = 1.0 / np.sqrt(inv_s2.astype(np.float64))
    uv, _ = project(true_pose, Xw, Kf)
    uv = uv + rng.normal(0, 1.0, uv.shape) * sigma[:, None]
    n_out = int(outlier_frac * n_points)
    idx = rng.choice(n_points, n_out, replace=False)
    uv[idx] += rng.uniform(-50, 50, (n_out, 2))
    init = true_pose.copy()
    init[:3] += rng.normal(0, 0.2 / np.sqrt(3), 3)
    init[3:] = quat_mul(quat_from_rotvec(rng.normal(0, 0.05 / np.sqrt(3), 3)), true_pose[3:])
    return dict(pose=init, true_pose=true_pose, Xw=Xw, uv=uv.astype(np.float32), inv_sigma2=inv_s2,
                K=Kf, octave=octv)


def make_ba_problem(n_cams: int, n_points: int, obs_per_point: int, seed: int, n_fixed_extra: int = 0,
                    window: int | None = None, K=KITTI_K, outlier_frac: float = 0.05,
                    pose_noise=(0.01, 0.05), point_noise: float = 0.05):
    """C4/C5: cameras on an arc (or loop) looking outward at a shell of points.
    Every point is seen by exactly `obs_per_point` cameras, chosen inside a +-window camera window when
    `window` is given (C5), else uniformly (C4).  Camera 0 is always fixed (KF id 0 constant,
    CeresOptimizer.cc:115-120); `n_fixed_extra` additional fixed cameras follow the variable ones."""
    rng = np.random.default_rng(seed)
    Kf = np.array(K, np.float32).astype(np.float64)
    fx, fy, cx, cy = Kf
    n_total = n_cams + n_fixed_extra
    # camera centres along an arc, each looking along +z of a slowly turning heading
    poses = np.zeros((n_total, 7))
    step = 0.5
    heading = np.linspace(0, (2 * np.pi if window is not None else 0.6), n_total, endpoint=False)
    centre = np.zeros((n_total, 3))
    for i in range(1, n_total):
        centre[i] = centre[i - 1] + step * np.array([np.cos(heading[i]), 0.0, np.sin(heading[i])])
    for i in range(n_total):
        # camera z axis = heading direction rotated by -90deg about y so that cameras look sideways/outward
        yaw = heading[i] - np.pi / 2
        q_wc = quat_from_rotvec(np.array([0.0, -yaw, 0.0]))     # world-from-camera
        R_wc = quat_to_R(q_wc)
        R_cw = R_wc.T
        q_cw = np.array([-q_wc[0], -q_wc[1], -q_wc[2], q_wc[3]])
        poses[i, :3] = -R_cw @ centre[i]
        poses[i, 3:] = q_cw
    # points: for each point choose an anchor camera, place the point in its frustum
    cam_idx = np.empty((n_points, obs_per_point), np.int32)
    Xw = np.empty((n_points, 3))
    for j in range(n_points):
        for _try in range(100):
            a = int(rng.integers(0, n_total))
            z = rng.uniform(5, 40)
            u = rng.uniform(200, 1041); v = rng.uniform(60, 316)
            pc = np.array([(u - cx) / fx * z, (v - cy) / fy * z, z])
            R = quat_to_R(poses[a, 3:])
            X = R.T @ (pc - poses[a, :3])
            if window is None:
                pool = np.arange(n_total)
            else:
                pool = (a + np.arange(-window, window + 1)) % n_total
            uvp, zp = project_many(poses[pool], X, Kf)
            ok = (zp > 1.0) & (uvp[:, 0] > 0) & (uvp[:, 0] < 1241) & (uvp[:, 1] > 0) & (uvp[:, 1] < 376)
            vis = pool[ok]
            if len(vis) >= obs_per_point:
                cam_idx[j] = np.sort(rng.choice(vis, obs_per_point, replace=False))
                Xw[j] = X
                break
        else:
            raise RuntimeError("could not place point")
    obs_cam = cam_idx.reshape(-1)
    obs_pt = np.repeat(np.arange(n_points, dtype=np.int32), obs_per_point)
    n_obs = obs_cam.size
    octv = _octaves(rng, n_obs)
    inv_s2 = inv_sigma2_table()[octv]
    sigma = 1.0 / np.sqrt(inv_s2.astype(np.float64))
    uv = np.empty((n_obs, 2))
    for c in np.unique(obs_cam):
        m = obs_cam == c
        uv[m], _ = project(poses[c], Xw[obs_pt[m]], Kf)
    uv += rng.normal(0, 1.0, uv.shape) * sigma[:, None]
    n_out = int(outlier_frac * n_obs)
    idx = rng.choice(n_obs, n_out, replace=False)
    uv[idx] += rng.uniform(-50, 50, (n_out, 2))
    fixed = np.zeros(n_total, np.uint8)
    fixed[0] = 1
    fixed[n_cams:] = 1
    init_poses = poses.copy()
    for i in range(n_total):
        if fixed[i]:
            continue
        init_poses[i, :3] += rng.normal(0, pose_noise[1] / np.sqrt(3), 3)
        init_poses[i, 3:] = quat_mul(quat_from_rotvec(rng.normal(0, pose_noise[0] / np.sqrt(3), 3)), poses[i, 3:])
    init_pts = Xw + rng.normal(0, point_noise / np.sqrt(3), Xw.shape)
    return dict(poses=init_poses, true_poses=poses, fixed=fixed, points=init_pts, true_points=Xw,
                obs_cam=obs_cam.astype(np.int32), obs_pt=obs_pt, uv=uv.astype(np.float32),
                inv_sigma2=inv_s2, K=Kf)


def project_many(poses: np.ndarray, X: np.ndarray, K):
    fx, fy, cx, cy = K
    uv = np.empty((len(poses), 2)); z = np.empty(len(poses))
    for i, p in enumerate(poses):
        pc = quat_to_R(p[3:]) @ X + p[:3]
        z[i] = pc[2]
        zz = pc[2] if abs(pc[2]) > 1e-9 else 1e-9
        uv[i] = (fx * pc[0] / zz + cx, fy * pc[1] / zz + cy)
    return uv, z


def make_ba_problem_fast(n_cams: int, n_points: int, obs_per_point: int, seed: int, window: int = 10, K=KITTI_K,
                         outlier_frac: float = 0.05, pose_noise=(0.01, 0.05), point_noise: float = 0.05):
    """Vectorised generator for large windowed graphs (C5: 1000 keyframes x 100k points x 500k observations).
    Cameras drive along a straight line looking sideways (+z of the camera is world +z), 0.5 m apart; a point
    anchored at camera a is seen by `obs_per_point` cameras of the window [a-window, a+window].  Same output keys
    as make_ba_problem."""
    rng = np.random.default_rng(seed)
    Kf = np.array(K, np.float32).astype(np.float64)
    fx, fy, cx, cy = Kf
    step = 0.5
    poses = np.zeros((n_cams, 7)); poses[:, 6] = 1.0
    centre_x = step * np.arange(n_cams)
    poses[:, 0] = -centre_x                       # t = -R c with R = I
    # points: anchor camera, depth, pixel in the anchor
    a = rng.integers(0, n_cams, n_points)
    z = rng.uniform(8, 40, n_points)
    u = rng.uniform(300, 941, n_points); v = rng.uniform(60, 316, n_points)
    Xw = np.stack([(u - cx) / fx * z + centre_x[a], (v - cy) / fy * z, z], 1)
    # visible cameras: |fx * (X - c_k)/z + cx| inside the image; choose obs_per_point offsets inside the window
    offs = np.arange(-window, window + 1)
    cam_idx = np.empty((n_points, obs_per_point), np.int64)
    cand = a[:, None] + offs[None, :]
    uu = fx * (Xw[:, 0:1] - step * cand) / z[:, None] + cx
    ok = (cand >= 0) & (cand < n_cams) & (uu > 5) & (uu < 1236)
    score = rng.random(cand.shape)
    score[~ok] = 2.0
    order = np.argsort(score, axis=1)[:, :obs_per_point]
    chosen_ok = np.take_along_axis(ok, order, 1)
    assert chosen_ok.all(), "window too small for obs_per_point visible cameras"
    cam_idx = np.sort(np.take_along_axis(cand, order, 1), axis=1)
    obs_cam = cam_idx.reshape(-1).astype(np.int32)
    obs_pt = np.repeat(np.arange(n_points, dtype=np.int32), obs_per_point)
    n_obs = obs_cam.size
    octv = _octaves(rng, n_obs)
    inv_s2 = inv_sigma2_table()[octv]
    sigma = 1.0 / np.sqrt(inv_s2.astype(np.float64))
    pc = Xw[obs_pt].copy(); pc[:, 0] -= centre_x[obs_cam]
    uv = np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], 1)
    uv += rng.normal(0, 1.0, uv.shape) * sigma[:, None]
    n_out = int(outlier_frac * n_obs)
    idx = rng.choice(n_obs, n_out, replace=False)
    uv[idx] += rng.uniform(-50, 50, (n_out, 2))
    fixed = np.zeros(n_cams, np.uint8); fixed[0] = 1
    init_poses = poses.copy()
    dt = rng.normal(0, pose_noise[1] / np.sqrt(3), (n_cams, 3))
    dr = rng.normal(0, pose_noise[0] / np.sqrt(3), (n_cams, 3))
    dt[0] = 0; dr[0] = 0
    init_poses[:, :3] += dt
    th = np.linalg.norm(dr, axis=1); th_safe = np.where(th > 0, th, 1.0)
    init_poses[:, 3:6] = dr / th_safe[:, None] * np.sin(th / 2)[:, None]
    init_poses[:, 6] = np.cos(th / 2)
    init_pts = Xw + rng.normal(0, point_noise / np.sqrt(3), Xw.shape)
    return dict(poses=init_poses, true_poses=poses, fixed=fixed, points=init_pts, true_points=Xw, obs_cam=obs_cam,
                obs_pt=obs_pt, uv=uv.astype(np.float32), inv_sigma2=inv_s2, K=Kf)


def make_sim3_problem(n: int = 300, seed: int = 6, K=KITTI_K, scale: float = 1.08, outlier_frac: float = 0.1,
                      init_noise=(0.03, 0.1, 0.03)):
    """OptimizeSim3 inputs: n map points seen by two keyframes whose maps differ by a similarity S12 (camera-1 point =
    s12 R12 * camera-2 point + t12).  Returns the per-correspondence arrays of cmos_ba_optimize_sim3, the true S12 and a
    perturbed initial one."""
    rng = np.random.default_rng(seed)
    Kf = np.array(K, np.float32).astype(np.float64)
    fx, fy, cx, cy = Kf
    # points in camera-1 coordinates (metric of map 1)
    z = rng.uniform(4, 40, n)
    u1 = rng.uniform(60, 1180, n); v1 = rng.uniform(40, 336, n)
    P1 = np.stack([(u1 - cx) / fx * z, (v1 - cy) / fy * z, z], 1)
    rv = rng.normal(0, 0.06, 3)
    th = np.linalg.norm(rv); k = rv / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R12 = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    t12 = np.array([0.4, -0.05, 0.2]) + rng.normal(0, 0.05, 3)
    P2 = ((P1 - t12) @ R12) / scale                      # S12^-1 * P1
    u2 = fx * P2[:, 0] / P2[:, 2] + cx; v2 = fy * P2[:, 1] / P2[:, 2] + cy
    oc1 = _octaves(rng, n); oc2 = _octaves(rng, n)
    is1 = inv_sigma2_table()[oc1]; is2 = inv_sigma2_table()[oc2]
    obs1 = np.stack([u1, v1], 1) + rng.normal(0, 1.0, (n, 2)) / np.sqrt(is1.astype(np.float64))[:, None]
    obs2 = np.stack([u2, v2], 1) + rng.normal(0, 1.0, (n, 2)) / np.sqrt(is2.astype(np.float64))[:, None]
    n_out = int(outlier_frac * n)
    idx = rng.choice(n, n_out, replace=False)
    obs1[idx[: n_out // 2]] += rng.uniform(-40, 40, (n_out // 2, 2))
    obs2[idx[n_out // 2:]] += rng.uniform(-40, 40, (n_out - n_out // 2, 2))
    # the two maps' own (noisy) point estimates
    P1n = P1 + rng.normal(0, 0.02, P1.shape); P2n = P2 + rng.normal(0, 0.02, P2.shape)
    rv0 = rng.normal(0, init_noise[0] / np.sqrt(3), 3)
    th0 = np.linalg.norm(rv0); k0 = rv0 / th0
    K0 = np.array([[0, -k0[2], k0[1]], [k0[2], 0, -k0[0]], [-k0[1], k0[0], 0]])
    R0 = (np.eye(3) + np.sin(th0) * K0 + (1 - np.cos(th0)) * K0 @ K0) @ R12
    return dict(obs1=obs1.astype(np.float32), inv_sigma1=is1, P3D2c=P2n, obs2=obs2.astype(np.float32), inv_sigma2=is2, P3D1c=P1n,
                K=np.array(K, np.float32), s_true=scale, R_true=R12, t_true=t12, s0=scale * float(np.exp(rng.normal(0, init_noise[2]))),
                R0=R0, t0=t12 + rng.normal(0, init_noise[1] / np.sqrt(3), 3))


def _rodrigues(rv):
    th = np.linalg.norm(rv)
    if th < 1e-300:
        return np.eye(3)
    k = rv / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def make_essential_graph_problem(n_kf: int = 60, seed: int = 7, n_points: int = 500, n_group: int = 4, loop_kf: int = 1,
                                 drift=(0.004, 0.01, 0.004), covis=(2, 3), radius: float = 20.0):
    """OptimizeEssentialGraph inputs after a loop closure: n_kf keyframes once around a circle, the estimated trajectory
    drifting in rotation / translation / scale (per-step sigmas `drift`), the last keyframe closing the loop on keyframe
    `loop_kf`.  The current keyframe and its n_group predecessors carry corrected Sim3s (and their non-corrected ones);
    edges in the reference's insertion order: loop connections (kind 0), then per keyframe its parent and its
    co-visible predecessors i - covis[k] (kind 1).  Sim3s as [scale, R row-major (9), t (3)]."""
    rng = np.random.default_rng(seed)
    # ground truth Tcw
    Rt, tt = [], []
    for k in range(n_kf):
        a = 2 * np.pi * k / n_kf * (1 - 0.5 / n_kf)
        Rwc = _rodrigues(np.array([0.0, -a, 0.0]))                       # yaw about the camera's y axis
        C = np.array([radius * np.cos(a), 0.2 * np.sin(3 * a), radius * np.sin(a)])
        Rt.append(Rwc.T); tt.append(-Rwc.T @ C)
    # drifting estimate: chain the true relative motions, each perturbed
    Re, te, se = [Rt[0]], [tt[0]], [1.0]
    for k in range(1, n_kf):
        Rrel = Rt[k] @ Rt[k - 1].T
        trel = tt[k] - Rrel @ tt[k - 1]
        s = se[-1] * float(np.exp(rng.normal(0, drift[2])))
        Rn = _rodrigues(rng.normal(0, drift[0] / np.sqrt(3), 3)) @ Rrel
        tn = s * trel + rng.normal(0, drift[1] / np.sqrt(3), 3)
        Re.append(Rn @ Re[-1]); te.append(Rn @ te[-1] + tn); se.append(s)
    Scw = np.zeros((n_kf, 13)); Snc = np.zeros((n_kf, 13)); flags = np.zeros(n_kf, np.uint8)
    for k in range(n_kf):
        Scw[k, 0] = 1.0; Scw[k, 1:10] = Re[k].reshape(-1); Scw[k, 10:] = te[k]
    flags[loop_kf] |= 1
    cur = n_kf - 1
    sig = se[cur]
    # corrected Sim3 of the current keyframe in the loop side's world, propagated to its group through the (non-corrected)
    # relative motions (LoopClosing::CorrectLoop)
    Rc, tc, sc = Rt[cur], sig * tt[cur], sig
    group = list(range(cur - n_group, cur + 1))
    for k in group:
        Ric = Re[k] @ Re[cur].T
        tic = te[k] - Ric @ te[cur]
        Snc[k] = Scw[k]
        flags[k] |= 2
        Scw[k, 0] = sc; Scw[k, 1:10] = (Ric @ Rc).reshape(-1); Scw[k, 10:] = Ric @ tc + tic
    ej, ei, ek = [], [], []
    loop_cluster = [j for j in (loop_kf - 1, loop_kf, loop_kf + 1) if 0 <= j < n_kf]
    for i in group:
        for j in loop_cluster:
            ej.append(j); ei.append(i); ek.append(0)
    for j in loop_cluster:
        for i in group[-2:]:
            ej.append(i); ei.append(j); ek.append(0)
    for i in range(n_kf):
        if i >= 1:
            ej.append(i - 1); ei.append(i); ek.append(1)
        for c in covis:
            if i - c >= 0:
                ej.append(i - c); ei.append(i); ek.append(1)
    ref = rng.integers(0, n_kf, n_points).astype(np.int32)
    Xw = np.zeros((n_points, 3))
    for p in range(n_points):
        k = ref[p]
        Pc = np.array([rng.uniform(-5, 5), rng.uniform(-1, 1), rng.uniform(4, 30)])
        Xw[p] = Re[k].T @ (Pc - te[k])
    true_T = np.zeros((n_kf, 4, 4))
    for k in range(n_kf):
        true_T[k, :3, :3] = Rt[k]; true_T[k, :3, 3] = tt[k]; true_T[k, 3, 3] = 1
    return dict(Scw=Scw, Snc=Snc, kf_flags=flags, edge_j=np.array(ej, np.int32), edge_i=np.array(ei, np.int32),
                edge_kind=np.array(ek, np.uint8), Xw=Xw, ref_kf=ref, true_Tcw=true_T, loop_kf=loop_kf, group=np.array(group))

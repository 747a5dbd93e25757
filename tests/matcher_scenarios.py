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

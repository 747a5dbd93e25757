"""CPU tests: the matcher oracle (oracle/matcher_oracle.cpp) against slow, independent Python restatements of
Frame::GetFeaturesInArea and the two SearchByProjection variants (the reference ships no tests of its own)."""
import numpy as np
import pytest

from ceres_mono_orb_slam2_b200 import synth
from oracle import pyoracle as po
from tests.matcher_scenarios import camera_arrays, extract_sequence, identity_T, points_view

POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def ham(a, b):
    return int(POP[np.bitwise_xor(a, b)].sum())


def py_features_in_area(kps, bounds, x, y, r, lo, hi):
    """Brute force over all keypoints, then ordered like the grid walk: (cell ix, cell iy, index)."""
    min_x, max_x, min_y, max_y, iw, ih = [np.float32(v) for v in bounds]
    x = np.float32(x); y = np.float32(y); r = np.float32(r)
    mcx0 = max(0, int(np.floor((x - min_x - r) * iw))); mcx1 = min(63, int(np.ceil((x - min_x + r) * iw)))
    mcy0 = max(0, int(np.floor((y - min_y - r) * ih))); mcy1 = min(47, int(np.ceil((y - min_y + r) * ih)))
    if mcx0 >= 64 or mcx1 < 0 or mcy0 >= 48 or mcy1 < 0:
        return []
    out = []
    for i, kp in enumerate(kps):
        px = int(np.floor(np.float32((kp["x"] - min_x) * iw) + np.float32(0.5)))   # round() for positives
        py = int(np.floor(np.float32((kp["y"] - min_y) * ih) + np.float32(0.5)))
        if not (0 <= px < 64 and 0 <= py < 48):
            continue
        if not (mcx0 <= px <= mcx1 and mcy0 <= py <= mcy1):
            continue
        if (lo > 0 or hi >= 0):
            if kp["octave"] < lo or (hi >= 0 and kp["octave"] > hi):
                continue
        if abs(np.float32(kp["x"] - x)) < r and abs(np.float32(kp["y"] - y)) < r:
            out.append((px, py, i))
    return [i for _, _, i in sorted(out)]


@pytest.fixture(scope="module")
def scene():
    frames, offs, ext, o = extract_sequence(640, 480, 2, 600, seed=31)
    return frames, offs, ext, o


def test_grid_and_features_in_area(scene):
    _, _, ext, o = scene
    kps, _ = ext[1]
    b, K4, sf = camera_arrays(640, 480, synth.TUM2_K, o.scale_factors)
    gs, gi = po.build_grid(kps, b)
    assert gs[-1] == len(gi) <= len(kps)
    # every cell's indices ascend (insertion order) and each keypoint appears at most once
    assert len(set(gi.tolist())) == len(gi)
    for c in range(64 * 48):
        seg = gi[gs[c]:gs[c + 1]]
        assert (np.diff(seg) > 0).all()
    rng = np.random.default_rng(0)
    for _ in range(150):
        x = rng.uniform(-20, 660); y = rng.uniform(-20, 500); r = rng.uniform(2, 60)
        lo, hi = [(-1, -1), (0, 2), (2, 3), (5, -1), (-1, 4)][rng.integers(0, 5)]
        got = po.features_in_area(kps, gs, gi, b, x, y, r, lo, hi).tolist()
        assert got == py_features_in_area(kps, b, x, y, r, lo, hi)


def py_search_frame(cur_kps, cur_desc, b, K4, sf, T, last_kps, flags, xw, ldesc, th, check_ori):
    n = len(cur_kps)
    match = np.full(n, -1, np.int32); claimed = np.zeros(n, np.uint8)
    hist = [[] for _ in range(30)]
    nm = 0
    T = T.reshape(4, 4)
    for i in range(len(last_kps)):
        if not flags[i] & 1:
            continue
        pc = T[:3, :3] @ xw[i] + T[:3, 3]
        xc, yc = np.float32(pc[0]), np.float32(pc[1]); invz = np.float32(1.0 / pc[2])
        if invz < 0:
            continue
        u = np.float32(np.float32(K4[0] * xc) * invz) + K4[2]; v = np.float32(np.float32(K4[1] * yc) * invz) + K4[3]
        if u < b[0] or u > b[1] or v < b[2] or v > b[3]:
            continue
        oc = int(last_kps[i]["octave"]); radius = np.float32(th) * sf[oc]
        cand = py_features_in_area(cur_kps, b, u, v, radius, oc - 1, oc + 1)
        best, bi = 256, -1
        for i2 in cand:
            if claimed[i2]:
                continue
            d = ham(ldesc[i], cur_desc[i2])
            if d < best:
                best, bi = d, i2
        if best <= 100:
            match[bi] = i; claimed[bi] = (flags[i] >> 1) & 1; nm += 1
            if check_ori:
                rot = np.float32(last_kps[i]["angle"] - cur_kps[bi]["angle"])
                if rot < 0:
                    rot = np.float32(rot + np.float32(360))
                bin_ = int(np.floor(np.float32(rot * np.float32(1.0 / 30)) + np.float32(0.5)))
                if bin_ == 30:
                    bin_ = 0
                hist[bin_].append(bi)
    if check_ori:
        sizes = [len(h) for h in hist]
        m1 = m2 = m3 = 0; i1 = i2_ = i3 = -1
        for i, s in enumerate(sizes):
            if s > m1:
                m3, m2, m1 = m2, m1, s; i3, i2_, i1 = i2_, i1, i
            elif s > m2:
                m3, m2 = m2, s; i3, i2_ = i2_, i
            elif s > m3:
                m3, i3 = s, i
        if m2 < np.float32(0.1) * np.float32(m1):
            i2_ = i3 = -1
        elif m3 < np.float32(0.1) * np.float32(m1):
            i3 = -1
        for i in range(30):
            if i not in (i1, i2_, i3):
                for idx in hist[i]:
                    match[idx] = -1; claimed[idx] = 0; nm -= 1      # ORBmatcher.cc:1262: map_points_[idx] = nullptr
    return match, nm


@pytest.mark.parametrize("check_ori", [True, False])
def test_search_by_projection_frame_vs_python(scene, check_ori):
    _, offs, ext, o = scene
    (lk, ld), (ck, cd) = ext
    b, K4, sf = camera_arrays(640, 480, synth.TUM2_K, o.scale_factors)
    shift = (offs[0] - offs[1]).astype(np.float64)     # scene point moves by -(crop motion)
    flags, xw, mdesc = synth.make_last_frame_view(lk, ld, shift, seed=5, K=synth.TUM2_K)
    gs, gi = po.build_grid(ck, b)
    T = identity_T()
    match, nm, _ = po.search_by_projection_frame(ck, cd, gs, gi, b, K4, sf, T, lk, flags, xw, mdesc, 15.0, check_ori)
    pm, pnm = py_search_frame(ck, cd, b, K4, sf, T, lk, flags, xw, mdesc, 15.0, check_ori)
    assert nm == pnm and np.array_equal(match, pm)
    assert nm > 0.4 * (flags & 1).sum(), "scenario should produce many matches"
    # matched pairs are mostly the true correspondences: descriptors within TH_HIGH
    idx = np.nonzero(match >= 0)[0]
    assert all(ham(mdesc[match[i]], cd[i]) <= 100 for i in idx)


def py_search_points(kps, desc, b, sf, in_view, level, view_cos, proj, mdesc, has_obs, th, ratio):
    n = len(kps)
    assign = np.full(n, -1, np.int32); claimed = np.zeros(n, np.uint8); nm = 0
    for p in range(len(in_view)):
        if not in_view[p]:
            continue
        r = np.float32(2.5) if float(view_cos[p]) > 0.998 else np.float32(4.0)
        if th != 1.0:
            r = np.float32(r * np.float32(th))
        lv = int(level[p])
        cand = py_features_in_area(kps, b, proj[p, 0], proj[p, 1], np.float32(r * sf[lv]), lv - 1, lv)
        b1, l1, b2, l2, bi = 256, -1, 256, -1, -1
        for idx in cand:
            if claimed[idx]:
                continue
            d = ham(mdesc[p], desc[idx])
            if d < b1:
                b2, l2 = b1, l1; b1, l1, bi = d, int(kps[idx]["octave"]), idx
            elif d < b2:
                b2, l2 = d, int(kps[idx]["octave"])
        if b1 <= 100:
            if l1 == l2 and np.float32(b1) > np.float32(ratio) * np.float32(b2):
                continue
            assign[bi] = p; claimed[bi] = has_obs[p]; nm += 1
    return assign, nm


@pytest.mark.parametrize("th", [1.0, 5.0])
def test_search_by_projection_points_vs_python(scene, th):
    _, offs, ext, o = scene
    (lk, ld), (ck, cd) = ext
    b, K4, sf = camera_arrays(640, 480, synth.TUM2_K, o.scale_factors)
    shift = (offs[0] - offs[1]).astype(np.float64)
    in_view, level, view_cos, proj, mdesc, has_obs = points_view(lk, ld, shift, seed=9)
    gs, gi = po.build_grid(ck, b)
    assign, nm, _ = po.search_by_projection_points(ck, cd, gs, gi, b, sf, in_view, level, view_cos, proj, mdesc,
                                                   has_obs, th, 0.8)
    pa, pnm = py_search_points(ck, cd, b, sf, in_view, level, view_cos, proj, mdesc, has_obs, th, 0.8)
    assert nm == pnm and np.array_equal(assign, pa)
    assert nm > 50


def test_descriptor_distance_popcount():
    rng = np.random.default_rng(1)
    for _ in range(200):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b2 = rng.integers(0, 256, 32, dtype=np.uint8)
        assert po.descriptor_distance(a, b2) == ham(a, b2)
    a = np.zeros(32, np.uint8)
    assert po.descriptor_distance(a, ~a) == 256 and po.descriptor_distance(a, a) == 0


def test_undistort_points_is_cv2_bit_exact():
    """Frame::UndistortKeyPoints / ComputeImageBounds call cv::undistortPoints(mat, mat, K, dist, Mat(), K); the
    restatement is pinned to OpenCV 4.13 through cv2 on the TUM1/TUM2 distortion models and a 4-coefficient one."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    cases = [((520.908620, 521.007327, 325.141442, 249.701764), (0.231222, -0.784899, -0.003257, -0.000105, 0.917205)),   # TUM2.yaml
             ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),   # TUM1.yaml
             ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05))]              # 4 coefficients
    for K4, dist in cases:
        K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
        D = np.array(dist, np.float32)
        pts = np.stack([rng.uniform(-30, 780, 40000), rng.uniform(-30, 520, 40000)], 1).astype(np.float32)
        ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)
        got = po.undistort_points(np.array(K4, np.float32), D, pts)
        assert np.array_equal(got, ref)
    # k1 == 0: Frame.cc:330-333 copies, bounds are the image rectangle
    assert np.array_equal(po.image_bounds(np.array(cases[0][0], np.float32), np.zeros(5, np.float32), 640, 480),
                          np.array([0, 640, 0, 480], np.float32))

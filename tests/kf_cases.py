"""Input builders and runners for the keyframe-matcher searches, shared by the CPU oracle tests
(tests/test_oracle_matcher2.py) and the GPU parity tests (tests/test_matcher_kf_gpu.py).  Every `case_*` returns
the flattened inputs of one reference call; `oracle_*` runs oracle/matcher2_oracle.cpp on them and `gpu_*` the
product (KeyFrameMatcher over the C ABI)."""
import numpy as np

from oracle import pyoracle as po
from tests.matcher_scenarios import T44, pose15


def bounds6(S):
    return np.array([0.0, S["width"], 0.0, S["height"], np.float32(64) / np.float32(S["width"]),
                     np.float32(48) / np.float32(S["height"])], np.float32)


def oracle_view(S, which, is_keyframe):
    k, d = (S["k1"], S["desc1"]) if which == 1 else (S["k2"], S["desc2"])
    return po.View(k, d, bounds6(S), np.array(S["K"], np.float32), S["sf"], S["scale"], is_keyframe)


def gpu_camera(S):
    from ceres_mono_orb_slam2_b200 import Camera
    return Camera.create(S["width"], S["height"], S["K"], S["sf"], S["scale"])


def gpu_bind(m, S, slot, which, is_keyframe):
    k, d = (S["k1"], S["desc1"]) if which == 1 else (S["k2"], S["desc2"])
    m.set_view(slot, gpu_camera(S), k, d, is_keyframe)


# ---- SearchByProjection(Frame, KeyFrame, found, th, ORBdist): current frame = view 2, keyframe = view 1 ----------
def case_reloc(S, th=15.0, orb_dist=100, seed=1):
    rng = np.random.default_rng(seed)
    n1, n2 = len(S["k1"]), len(S["k2"])
    T = T44(S["R2"], S["t2"]).reshape(-1)
    return dict(Tcw=T, kf_valid=(rng.random(n1) < 0.9).astype(np.uint8), kf_xw=S["Xw"], kf_min_d=S["min_d"],
                kf_max_d=S["max_d"], kf_desc=S["mp_desc"], kf_angle=S["k1"]["angle"].copy(), th=th, orb_dist=orb_dist,
                cur_has_point=(rng.random(n2) < 0.05).astype(np.uint8))


def oracle_reloc(S, c, check_ori=True):
    V = oracle_view(S, 2, False)
    return po.search_by_projection_reloc(V, c["Tcw"], c["kf_valid"], c["kf_xw"], c["kf_min_d"], c["kf_max_d"], c["kf_desc"],
                                         c["kf_angle"], c["th"], c["orb_dist"], check_ori, c["cur_has_point"])


def gpu_reloc(m, S, c):
    gpu_bind(m, S, 0, 2, False)
    return m.SearchByProjectionKeyFrame(c["Tcw"], c["kf_valid"], c["kf_xw"], c["kf_min_d"], c["kf_max_d"], c["kf_desc"],
                                        c["kf_angle"], c["th"], c["orb_dist"], c["cur_has_point"])


# ---- points projected into view 2 as a keyframe: SearchByProjection(KF, Scw, ...), Fuse x2 -----------------------------
def case_points(S, seed=2, s=1.3):
    rng = np.random.default_rng(seed)
    n, n2 = len(S["Xw"]), len(S["k2"])
    # duplicates and a shuffled order: several points compete for the same keypoint
    order = np.concatenate([rng.permutation(n), rng.integers(0, n, n // 5)])
    Scw = T44(S["R2"], S["t2"], s)
    return dict(order=order, Scw=Scw.reshape(-1), pose15=pose15(S["R2"], S["t2"]),
                pt_skip=(rng.random(len(order)) < 0.1).astype(np.uint8), xw=S["Xw"][order], normal=S["normal"][order],
                min_d=S["min_d"][order], max_d=S["max_d"][order], pt_desc=S["mp_desc"][order],
                matched=(rng.random(n2) < 0.05).astype(np.uint8),
                inv_sigma2=(1.0 / (S["sf"] * S["sf"])).astype(np.float32))


def oracle_proj_sim3(S, c, th=10):
    V = oracle_view(S, 2, True)
    return po.search_by_projection_sim3(V, c["Scw"], c["pt_skip"], c["xw"], c["normal"], c["min_d"], c["max_d"], c["pt_desc"],
                                        th, c["matched"])


def gpu_proj_sim3(m, S, c, th=10):
    gpu_bind(m, S, 0, 2, True)
    return m.SearchByProjectionSim3(c["Scw"], c["pt_skip"], c["xw"], c["normal"], c["min_d"], c["max_d"], c["pt_desc"], th,
                                    c["matched"])


def oracle_fuse(S, c, sim3, th=3.0):
    V = oracle_view(S, 2, True)
    return po.fuse(V, c["inv_sigma2"], sim3, c["Scw"] if sim3 else c["pose15"], c["pt_skip"], c["xw"], c["normal"], c["min_d"],
                   c["max_d"], c["pt_desc"], th)


def gpu_fuse(m, S, c, sim3, th=3.0):
    gpu_bind(m, S, 0, 2, True)
    return m.Fuse(c["Scw"] if sim3 else c["pose15"], c["pt_skip"], c["xw"], c["normal"], c["min_d"], c["max_d"], c["pt_desc"],
                  th, inv_level_sigma2=c["inv_sigma2"], sim3=sim3)


# ---- SearchBySim3: keyframe 1 = view 1, keyframe 2 = view 2 ---------------------------------------------------------------
def case_sim3(S, seed=3, s12=1.0):
    rng = np.random.default_rng(seed)
    n1, n2 = len(S["k1"]), len(S["k2"])
    R12 = S["R1"] @ S["R2"].T
    t12 = S["t1"] - R12 @ S["t2"]
    p2 = np.maximum(S["p2"], 0)
    side1 = ((rng.random(n1) < 0.9).astype(np.uint8), (rng.random(n1) < 0.05).astype(np.uint8), S["Xw"], S["min_d"],
             S["max_d"], S["mp_desc"])
    side2 = (((S["p2"] >= 0) & (rng.random(n2) < 0.9)).astype(np.uint8), (rng.random(n2) < 0.05).astype(np.uint8),
             S["Xw"][p2], S["min_d"][p2], S["max_d"][p2], S["mp_desc"][p2])
    return dict(pose1=np.concatenate([S["R1"].reshape(-1), S["t1"]]), pose2=np.concatenate([S["R2"].reshape(-1), S["t2"]]),
                s12=s12, R12=R12.reshape(-1), t12=t12, side1=side1, side2=side2)


def oracle_sim3(S, c, th=7.5):
    return po.search_by_sim3(oracle_view(S, 1, True), oracle_view(S, 2, True), c["pose1"], c["pose2"], c["s12"], c["R12"],
                             c["t12"], c["side1"], c["side2"], th)


def gpu_sim3(m, S, c, th=7.5):
    gpu_bind(m, S, 0, 1, True); gpu_bind(m, S, 1, 2, True)
    return m.SearchBySim3(c["pose1"], c["pose2"], c["s12"], c["R12"], c["t12"], c["side1"], c["side2"], th)


# ---- bag of words ------------------------------------------------------------------------------------------------------------
def case_bow(S, seed=4):
    rng = np.random.default_rng(seed)
    n1, n2 = len(S["k1"]), len(S["k2"])
    return dict(valid1=(rng.random(n1) < 0.9).astype(np.uint8), valid2=((S["p2"] >= 0) | (rng.random(n2) < 0.3)).astype(np.uint8),
                fv1=po.flatten_feature_vector(S["node1"]), fv2=po.flatten_feature_vector(S["node2"]))


def oracle_bow(S, c, mode, nn_ratio=0.75, check_ori=True):
    return po.search_by_bow(mode, S["desc1"], S["k1"]["angle"], c["valid1"], c["fv1"], S["desc2"], S["k2"]["angle"],
                            c["valid2"] if mode == 1 else None, c["fv2"], nn_ratio, check_ori)


def gpu_bow(m, S, c, mode):
    from ceres_mono_orb_slam2_b200 import FeatureVector
    gpu_bind(m, S, 0, 1, True); gpu_bind(m, S, 1, 2, mode == 1)
    return m.SearchByBoW(mode, c["valid1"], FeatureVector.create(*c["fv1"]), c["valid2"] if mode == 1 else None,
                         FeatureVector.create(*c["fv2"]))


def case_triangulation(S, seed=5):
    rng = np.random.default_rng(seed)
    n1, n2 = len(S["k1"]), len(S["k2"])
    fx, fy, cx, cy = [float(np.float32(v)) for v in S["K"]]
    Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]])
    R12 = S["R1"] @ S["R2"].T
    t12 = S["t1"] - R12 @ S["t2"]
    tx = np.array([[0, -t12[2], t12[1]], [t12[2], 0, -t12[0]], [-t12[1], t12[0], 0]])
    F12 = np.linalg.inv(Km).T @ tx @ R12 @ np.linalg.inv(Km)     # LocalMapping::ComputeF12
    c = case_bow(S, seed)
    c.update(has1=(rng.random(n1) < 0.3).astype(np.uint8), has2=(rng.random(n2) < 0.3).astype(np.uint8), F12=F12.reshape(-1),
             Cw=S["Ow1"], R2w=S["R2"].reshape(-1), t2w=S["t2"], sigma2=(S["sf"] * S["sf"]).astype(np.float32))
    return c


def oracle_triangulation(S, c, check_ori=True):
    return po.search_for_triangulation(S["k1"], S["desc1"], c["has1"], c["fv1"], S["k2"], S["desc2"], c["has2"], c["fv2"],
                                       c["F12"], c["Cw"], c["R2w"], c["t2w"], np.array(S["K"], np.float32), S["sf"],
                                       c["sigma2"], check_ori)


def gpu_triangulation(m, S, c):
    from ceres_mono_orb_slam2_b200 import FeatureVector
    gpu_bind(m, S, 0, 1, True); gpu_bind(m, S, 1, 2, True)
    return m.SearchForTriangulation(c["has1"], FeatureVector.create(*c["fv1"]), c["has2"], FeatureVector.create(*c["fv2"]),
                                    c["F12"], c["Cw"], c["R2w"], c["t2w"], c["sigma2"])


# ---- SearchForInitialization: F1 = view 1, F2 = view 2 -----------------------------------------------------------------------
def case_init(S):
    return dict(prev=np.stack([S["k1"]["x"], S["k1"]["y"]], 1).astype(np.float32))


def oracle_init(S, c, window=100, nn_ratio=0.9, check_ori=True):
    return po.search_for_initialization(S["k1"], S["desc1"], oracle_view(S, 2, False), c["prev"], window, nn_ratio, check_ori)


def gpu_init(m, S, c, window=100):
    gpu_bind(m, S, 0, 1, False); gpu_bind(m, S, 1, 2, False)
    return m.SearchForInitialization(c["prev"], window)


# ---- the REFERENCE's own ORBmatcher over real KeyFrame / MapPoint objects (oracle/_ref, oracle/ref_shim/ref_matcher.cpp) ----
def _ref_view(S, which):
    from oracle.pyoracle import _p
    k, d = (S["k1"], S["desc1"]) if which == 1 else (S["k2"], S["desc2"])
    k = np.ascontiguousarray(k); d = np.ascontiguousarray(d, np.uint8)
    return k, d, (_p(k), _p(d), len(k))


def _ref_cam(S):
    from oracle.pyoracle import _p
    b = bounds6(S); K = np.array(S["K"], np.float32); sf = np.ascontiguousarray(S["sf"], np.float32)
    return (b, K, sf), (_p(b), _p(K), _p(sf), len(sf))


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


def ref_reloc(S, c, check_ori=True):
    import ctypes as C
    from oracle import pyref as pr
    from oracle.pyoracle import _p
    k, d, v = _ref_view(S, 2); keep, cam = _ref_cam(S)
    hp = _c(c["cur_has_point"], np.uint8).copy(); match = np.full(max(len(k), 1), -1, np.int32)
    T = _c(c["Tcw"], np.float64); va = _c(c["kf_valid"], np.uint8); xw = _c(c["kf_xw"], np.float64)
    mn = _c(c["kf_min_d"], np.float32); mx = _c(c["kf_max_d"], np.float32); kd = _c(c["kf_desc"], np.uint8); an = _c(c["kf_angle"], np.float32)
    nm = pr.lib().ref_search_by_projection_reloc(*v, *cam, _p(T), len(va), _p(va), _p(xw), _p(mn), _p(mx), _p(kd), _p(an),
                                                 C.c_float(c["th"]), int(c["orb_dist"]), int(check_ori), _p(hp), _p(match))
    return match[:len(k)], nm, hp


def ref_proj_sim3(S, c, th=10):
    from oracle import pyref as pr
    from oracle.pyoracle import _p
    k, d, v = _ref_view(S, 2); keep, cam = _ref_cam(S)
    m = _c(c["matched"], np.uint8).copy(); assign = np.full(max(len(k), 1), -1, np.int32)
    Scw = _c(c["Scw"], np.float64); sk = _c(c["pt_skip"], np.uint8); x = _c(c["xw"], np.float64); nr = _c(c["normal"], np.float64)
    mn = _c(c["min_d"], np.float32); mx = _c(c["max_d"], np.float32); pd = _c(c["pt_desc"], np.uint8)
    nm = pr.lib().ref_search_by_projection_sim3(*v, *cam, _p(Scw), len(sk), _p(sk), _p(x), _p(nr), _p(mn), _p(mx), _p(pd), int(th),
                                                _p(m), _p(assign))
    return assign[:len(k)], nm, m


def ref_fuse(S, c, sim3, th=3.0):
    """Returns (added_idx, replace_idx, nFused): keypoint index where point p was ADDED to the keyframe (-1: not added), for the
    Sim3 variant the point p that vpReplacePoint[p] names, and the reference's return value."""
    import ctypes as C
    from oracle import pyref as pr
    from oracle.pyoracle import _p
    k, d, v = _ref_view(S, 2); keep, cam = _ref_cam(S)
    n = len(c["pt_skip"])
    added = np.full(max(n, 1), -1, np.int32); repl = np.full(max(n, 1), -1, np.int32)
    R2 = np.asarray(c["pose15"][:9]).reshape(3, 3); t2 = np.asarray(c["pose15"][9:12])
    pose = _c(c["Scw"], np.float64) if sim3 else _c(np.concatenate([R2.reshape(-1), t2]), np.float64)
    sk = _c(c["pt_skip"], np.uint8); x = _c(c["xw"], np.float64); nr = _c(c["normal"], np.float64)
    mn = _c(c["min_d"], np.float32); mx = _c(c["max_d"], np.float32); pd = _c(c["pt_desc"], np.uint8)
    nf = pr.lib().ref_fuse(*v, *cam, int(sim3), _p(pose), n, _p(sk), _p(x), _p(nr), _p(mn), _p(mx), _p(pd), C.c_float(th),
                           _p(added), _p(repl))
    return added[:n], repl[:n], nf


def sim3_consistent_case(c, n2, seed=0):
    """SearchBySim3 derives vbAlreadyMatched2 from vpMatches12 (the keyframe-2 index of every pre-matched point, :990-1000):
    pair the case's two independent `already` sets up so that the oracle and the reference see the same input."""
    rng = np.random.default_rng(seed)
    a1 = np.nonzero(c["side1"][1])[0]; a2 = np.nonzero(c["side2"][1])[0]
    idx2 = np.full(len(c["side1"][1]), -2, np.int32)
    take = a2[: len(a1)]
    idx2[a1] = -1
    idx2[a1[: len(take)]] = take
    already2 = np.zeros(n2, np.uint8); already2[take] = 1
    c2 = dict(c)
    c2["side2"] = (c["side2"][0], already2) + tuple(c["side2"][2:])
    c2["already1_idx2"] = idx2
    return c2


def ref_sim3(S, c, th=7.5):
    import ctypes as C
    from oracle import pyref as pr
    from oracle.pyoracle import _p
    k1, d1, v1 = _ref_view(S, 1); k2, d2, v2 = _ref_view(S, 2); keep, cam = _ref_cam(S)
    s1, s2 = c["side1"], c["side2"]
    a = [_c(c["pose1"], np.float64), _c(c["pose2"], np.float64), _c(c["R12"], np.float64), _c(c["t12"], np.float64),
         _c(s1[0], np.uint8), _c(c["already1_idx2"], np.int32), _c(s1[2], np.float64), _c(s1[3], np.float32), _c(s1[4], np.float32),
         _c(s1[5], np.uint8), _c(s2[0], np.uint8), _c(s2[2], np.float64), _c(s2[3], np.float32), _c(s2[4], np.float32), _c(s2[5], np.uint8)]
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    nf = pr.lib().ref_search_by_sim3(*v1, *v2, *cam, _p(a[0]), _p(a[1]), C.c_float(c["s12"]), _p(a[2]), _p(a[3]), *[_p(x) for x in a[4:]],
                                     C.c_float(th), _p(m12))
    return m12[:len(k1)], nf


def ref_bow(S, c, mode, nn_ratio=0.75, check_ori=True):
    import ctypes as C
    from oracle import pyref as pr
    from oracle.pyoracle import _p, _fv
    k1, d1, v1 = _ref_view(S, 1); k2, d2, v2 = _ref_view(S, 2); keep, cam = _ref_cam(S)
    va1 = _c(c["valid1"], np.uint8); va2 = _c(c["valid2"] if mode == 1 else np.ones(len(k2)), np.uint8)
    n1, s1, f1 = _fv(c["fv1"]); n2, s2, f2 = _fv(c["fv2"])
    n_out = len(k2) if mode == 0 else len(k1)
    match = np.full(max(n_out, 1), -1, np.int32)
    nm = pr.lib().ref_search_by_bow(int(mode), *v1, _p(va1), len(n1), _p(n1), _p(s1), _p(f1), *v2, _p(va2), len(n2), _p(n2), _p(s2),
                                    _p(f2), *cam, C.c_float(nn_ratio), int(check_ori), _p(match))
    return match[:n_out], nm


def ref_triangulation(S, c, check_ori=True):
    from oracle import pyref as pr
    from oracle.pyoracle import _p, _fv
    k1, d1, v1 = _ref_view(S, 1); k2, d2, v2 = _ref_view(S, 2); keep, cam = _ref_cam(S)
    h1 = _c(c["has1"], np.uint8); h2 = _c(c["has2"], np.uint8)
    n1, s1, f1 = _fv(c["fv1"]); n2, s2, f2 = _fv(c["fv2"])
    F = _c(c["F12"], np.float64)
    p1 = _c(np.concatenate([S["R1"].reshape(-1), S["t1"]]), np.float64); p2 = _c(np.concatenate([S["R2"].reshape(-1), S["t2"]]), np.float64)
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    nm = pr.lib().ref_search_for_triangulation(*v1, _p(h1), len(n1), _p(n1), _p(s1), _p(f1), *v2, _p(h2), len(n2), _p(n2), _p(s2),
                                               _p(f2), _p(F), _p(p1), _p(p2), *cam, int(check_ori), _p(m12))
    return m12[:len(k1)], nm


def ref_init(S, c, window=100, nn_ratio=0.9, check_ori=True):
    import ctypes as C
    from oracle import pyref as pr
    from oracle.pyoracle import _p
    k1, d1, v1 = _ref_view(S, 1); k2, d2, v2 = _ref_view(S, 2); keep, cam = _ref_cam(S)
    prev = _c(c["prev"], np.float32).copy(); m12 = np.full(max(len(k1), 1), -1, np.int32)
    nm = pr.lib().ref_search_for_initialization(*v1, *v2, *cam, _p(prev), int(window), C.c_float(nn_ratio), int(check_ori), _p(m12))
    return m12[:len(k1)], nm, prev

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

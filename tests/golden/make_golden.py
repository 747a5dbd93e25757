"""Generates tests/golden/orb_cv2_golden.npz — run in the build container, where cv2 (opencv-python-headless 4.13, the only
OpenCV available; the reference leaves its OpenCV version unpinned, CMakeLists.txt:21) can be imported:

    python tests/golden/make_golden.py

Every array named cv2_* is the output of the OpenCV entry point the reference calls (cv::resize ORBextractor.cc:1120,
cv::copyMakeBorder :1122, cv::FAST :809/:814, cv::GaussianBlur :1086, cv::fastAtan2 :103, cv::undistortPoints Frame.cc:346) on
the seeded inputs stored next to it; they pin the C++ oracle (and through it the CUDA path) without cv2 at test time.
Arrays named oracle_* are outputs of oracle/ (regression vectors for the parts of the path that are the reference's own code:
quadtree, orientation, rBRIEF, the solvers) — they guard against drift, they are not an independent pin."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from ceres_mono_orb_slam2_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

cv2.setNumThreads(1)


def main():
    out = {}
    W, H = 320, 240
    img = synth.make_image(W, H, 21, n_rect=90, n_blob=60)
    out["image"] = img
    # primitives
    out["cv2_resize_267x200"] = cv2.resize(img, (267, 200), interpolation=cv2.INTER_LINEAR)
    out["cv2_blur7"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    for t in (20, 7):
        det = cv2.FastFeatureDetector_create(t, True)
        k = det.detect(img)
        out[f"cv2_fast_t{t}"] = np.array([(p.pt[0], p.pt[1], p.response) for p in k], np.float32).reshape(-1, 3)
    rng = np.random.default_rng(5)
    y = rng.integers(-70000, 70000, 512).astype(np.float32); x = rng.integers(-70000, 70000, 512).astype(np.float32)
    y[:8] = [0, 0, 1, -1, 5, -5, 0, 3]; x[:8] = [0, 1, 0, 0, 5, 5, -2, -3]
    out["atan2_y"] = y; out["atan2_x"] = x
    out["cv2_fast_atan2"] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    # the extractor's front half, stage by stage, from cv2 calls (TUM2-like parameters: 500 features, 8 levels, 1.2)
    o = po.OrbOracle(500, 1.2, 8, 20, 7)
    pyr = po.cv2_pyramid(img, o.inv_scale_factors)
    for l in range(8):
        out[f"cv2_level{l}"] = pyr[l]
        out[f"cv2_candidates{l}"] = po.cv2_level_candidates(pyr[l])
    kps, desc = o.extract(img)
    out["oracle_keypoints"] = kps; out["oracle_descriptors"] = desc
    # cv::undistortPoints as Frame::UndistortKeyPoints calls it (TUM1 and TUM2 distortion models)
    pts = np.stack([rng.uniform(0, 640, 64), rng.uniform(0, 480, 64)], 1).astype(np.float32)
    out["undistort_xy"] = pts
    for name, K4, dist in (("tum1", (517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
                           ("tum2", (520.908620, 521.007327, 325.141442, 249.701764), (0.231222, -0.784899, -0.003257, -0.000105, 0.917205))):
        K = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
        d = np.array(dist, np.float32)
        out[f"undistort_{name}_K4"] = np.array(K4, np.float32); out[f"undistort_{name}_dist"] = d
        out[f"cv2_undistort_{name}"] = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, d, None, K).reshape(-1, 2).astype(np.float32)
    # solver regression vectors (oracle outputs; the reference pins nothing here)
    P = synth.make_pose_problem(200, seed=31)
    pose, outl, inl, s, _ = po.ba_pose_optimization(P["pose"], P["Xw"], P["uv"], P["inv_sigma2"], P["K"], 100)
    out["oracle_pose_result"] = pose; out["oracle_pose_outliers"] = outl; out["oracle_pose_iterations"] = np.int32(s["iterations"])
    G = synth.make_ba_problem(6, 200, 4, seed=11)
    cams, pts3, erase, ss = po.ba_local(G["poses"], G["fixed"], G["points"], G["obs_cam"], G["obs_pt"], G["uv"], G["inv_sigma2"], G["K"])
    out["oracle_local_cams"] = cams; out["oracle_local_points"] = pts3; out["oracle_local_erase"] = erase
    out["oracle_local_iterations"] = np.array([x["iterations"] for x in ss], np.int32)
    E = synth.make_essential_graph_problem(30, seed=33, n_points=40)
    r = po.essential_graph(E["Scw"], E["kf_flags"], E["Snc"], E["edge_j"], E["edge_i"], E["edge_kind"], E["Xw"], E["ref_kf"])
    out["oracle_essential_lie"] = r["lie"]; out["oracle_essential_points"] = r["Xw"]; out["oracle_essential_iterations"] = np.int32(r["iterations"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "orb_cv2_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(out), "arrays; cv2", cv2.__version__)


if __name__ == "__main__":
    main()

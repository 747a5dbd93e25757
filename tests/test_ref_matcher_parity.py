"""Parity pinned to the REFERENCE ITSELF for the per-frame matcher and the Frame glue: oracle/_ref/libref.so contains the
reference's own, unmodified src/{ORBmatcher,Frame,KeyFrame,MapPoint,Map,KeyFrameDatabase}.cc (compiled against the OpenCV /
Eigen / glog stand-ins of oracle/ref_shim/); oracle/ref_shim/ref_matcher.cpp builds real Frame / MapPoint objects from the
arrays the oracle and the CUDA path take and calls the reference's methods.

CPU tests: reference == oracle (oracle/matcher_oracle.cpp).  GPU tests: reference == CUDA through the C ABI.
Rows of SURVEY.md §8: a8 DescriptorDistance, a9 SearchByProjection(F, points), a10 SearchByProjection(cur, last), a22 grid /
GetFeaturesInArea / isInFrustum / PredictScale, 8f-1 Frame::Frame (undistortion, bounds, grid).
"""
import numpy as np
import pytest

from ceres_mono_orb_slam2_b200 import synth
from oracle import pyoracle as po
from oracle import pyref as pr
from tests.matcher_scenarios import camera_arrays, extract_sequence, identity_T, points_view

pytestmark = pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libref.so not built (needs /root/reference once)")


@pytest.fixture(scope="module")
def scene():
    return extract_sequence(640, 480, 2, 600, seed=31)


@pytest.fixture(scope="module")
def kitti():
    return extract_sequence(1241, 376, 3, 2000, seed=77)


def test_descriptor_distance():
    rng = np.random.default_rng(1)
    for _ in range(100):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert pr.descriptor_distance(a, b) == po.descriptor_distance(a, b)


def test_grid_and_features_in_area(scene):
    _, _, ext, o = scene
    kps, _ = ext[1]
    b, K4, sf = camera_arrays(640, 480, synth.TUM2_K, o.scale_factors)
    gs, gi = po.build_grid(kps, b)
    rgs, rgi = pr.build_grid(kps, b, sf)
    assert np.array_equal(gs, rgs) and np.array_equal(gi, rgi)
    rng = np.random.default_rng(0)
    for _ in range(150):
        x = rng.uniform(-20, 660); y = rng.uniform(-20, 500); r = rng.uniform(2, 60)
        lo, hi = [(-1, -1), (0, 2), (2, 3), (5, -1), (-1, 4)][rng.integers(0, 5)]
        assert np.array_equal(pr.features_in_area(kps, b, sf, x, y, r, lo, hi), po.features_in_area(kps, gs, gi, b, x, y, r, lo, hi))


@pytest.mark.parametrize("check_ori,th", [(True, 15.0), (False, 15.0), (True, 7.0), (True, 30.0)])
def test_search_by_projection_frame(scene, check_ori, th):
    _, offs, ext, o = scene
    (lk, ld), (ck, cd) = ext
    b, K4, sf = camera_arrays(640, 480, synth.TUM2_K, o.scale_factors)
    shift = (offs[0] - offs[1]).astype(np.float64)
    flags, xw, mdesc = synth.make_last_frame_view(lk, ld, shift, seed=5, K=synth.TUM2_K)
    gs, gi = po.build_grid(ck, b)
    T = identity_T()
    om, onm, ocl = po.search_by_projection_frame(ck, cd, gs, gi, b, K4, sf, T, lk, flags, xw, mdesc, th, check_ori)
    rm, rnm, rcl = pr.search_by_projection_frame(ck, cd, b, K4, sf, T, lk, flags, xw, mdesc, th, check_ori)
    assert rnm == onm and np.array_equal(rm, om) and np.array_equal(rcl, ocl)
    assert onm > 100


def test_search_by_projection_frame_kitti_with_claimed_keypoints(kitti):
    """configs[1] geometry, a non-identity pose, and keypoints that already hold an observed map point."""
    _, offs, ext, o = kitti
    b, K4, sf = camera_arrays(1241, 376, synth.KITTI_K, o.scale_factors)
    rng = np.random.default_rng(4)
    for f in (1, 2):
        (lk, ld), (ck, cd) = ext[f - 1], ext[f]
        shift = (offs[f - 1] - offs[f]).astype(np.float64)
        flags, xw, mdesc = synth.make_last_frame_view(lk, ld, shift, seed=50 + f)
        T = np.array([[1, 1e-4 * f, 0, 0.002], [-1e-4 * f, 1, 0, -0.001], [0, 0, 1, 0.004], [0, 0, 0, 1]]).reshape(-1)
        claimed = (rng.random(len(ck)) < 0.15).astype(np.uint8)
        gs, gi = po.build_grid(ck, b)
        om, onm, ocl = po.search_by_projection_frame(ck, cd, gs, gi, b, K4, sf, T, lk, flags, xw, mdesc, 15.0, True, claimed=claimed.copy())
        rm, rnm, rcl = pr.search_by_projection_frame(ck, cd, b, K4, sf, T, lk, flags, xw, mdesc, 15.0, True, claimed=claimed.copy())
        assert rnm == onm and np.array_equal(rm, om) and np.array_equal(rcl, ocl)
        assert onm > 300


@pytest.mark.parametrize("th,ratio", [(1.0, 0.8), (5.0, 0.8), (3.0, 0.6)])
def test_search_by_projection_points(scene, th, ratio):
    _, offs, ext, o = scene
    (lk, ld), (ck, cd) = ext
    b, K4, sf = camera_arrays(640, 480, synth.TUM2_K, o.scale_factors)
    shift = (offs[0] - offs[1]).astype(np.float64)
    in_view, level, view_cos, proj, mdesc, has_obs = points_view(lk, ld, shift, seed=9)
    gs, gi = po.build_grid(ck, b)
    oa, onm, ocl = po.search_by_projection_points(ck, cd, gs, gi, b, sf, in_view, level, view_cos, proj, mdesc, has_obs, th, ratio)
    ra, rnm, rcl = pr.search_by_projection_points(ck, cd, b, sf, in_view, level, view_cos, proj, mdesc, has_obs, th, ratio)
    assert rnm == onm and np.array_equal(ra, oa) and np.array_equal(rcl, ocl)
    assert onm > 50


def _frustum_case(seed, n=4000):
    rng = np.random.default_rng(seed)
    ang = 0.3
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([0.3, -0.1, 0.5]); Ow = -R.T @ t
    pose15 = np.concatenate([R.reshape(-1), t, Ow])
    xw = rng.uniform(-30, 30, (n, 3)); xw[:, 2] = rng.uniform(-5, 60, n)
    normal = rng.normal(size=(n, 3)); normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    toward = (Ow - xw); toward /= np.linalg.norm(toward, axis=1, keepdims=True)
    normal[: n // 2] = -toward[: n // 2] + 0.2 * normal[: n // 2]          # half the points face the camera
    normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    dist = np.linalg.norm(xw - Ow, axis=1)
    max_d = (dist * rng.uniform(0.7, 3.0, n)).astype(np.float32); min_d = (max_d / np.float32(3.58)).astype(np.float32)
    return pose15, xw, normal, min_d, max_d


@pytest.mark.parametrize("seed", [0, 1])
def test_is_in_frustum_and_predict_scale(seed):
    o = po.OrbOracle(2000)
    sf = o.scale_factors
    pose15, xw, normal, min_d, max_d = _frustum_case(seed)
    K4 = np.array(synth.KITTI_K, np.float32); b4 = np.array([0, 1241, 0, 376], np.float32)
    a = po.is_in_frustum(pose15, K4, b4, np.float32(np.log(np.float32(sf[1]))), 8, 0.5, xw, normal, min_d, max_d)
    r = pr.is_in_frustum(pose15, K4, b4, sf, 0.5, xw, normal, min_d, max_d)
    assert np.array_equal(a[0], r[0]) and a[0].sum() > 100
    m = a[0].astype(bool)
    for x, y in zip(a[1:], r[1:]):
        assert np.array_equal(x[m], y[m])


@pytest.mark.parametrize("w,h,nfeat,K,dist", [
    (640, 480, 1000, synth.TUM2_K, (0.231222, -0.784899, -0.003257, -0.000105, 0.917205)),     # TUM2.yaml: distorted
    (1241, 376, 2000, synth.KITTI_K, (0.0, 0.0, 0.0, 0.0))])                                   # KITTI00-02.yaml: rectified
def test_frame_constructor(w, h, nfeat, K, dist):
    """Frame::Frame end to end by the reference's own code (extraction, UndistortKeyPoints, ComputeImageBounds, grid pitch,
    AssignFeaturesToGrid) against the oracle chain: ORB oracle -> cv2-pinned undistortion -> grid."""
    img = synth.make_image(w, h, 21)
    kps, un, desc, b6, gs, gi = pr.frame_construct(img, nfeat, K, dist)
    ok, od = po.OrbOracle(nfeat).extract(img)
    assert np.array_equal(kps, ok) and np.array_equal(desc, od)
    K4 = np.array(K, np.float32); D = np.array(dist, np.float32)
    if dist[0] != 0.0:
        xy = po.undistort_points(K4, D, np.stack([ok["x"], ok["y"]], 1))
        exp = ok.copy(); exp["x"] = xy[:, 0]; exp["y"] = xy[:, 1]
    else:
        exp = ok
    assert np.array_equal(un, exp)
    bounds = po.image_bounds(K4, D, w, h)
    exp6 = np.array([bounds[0], bounds[1], bounds[2], bounds[3], np.float32(64) / np.float32(bounds[1] - bounds[0]),
                     np.float32(48) / np.float32(bounds[3] - bounds[2])], np.float32)
    assert np.array_equal(b6, exp6)
    ogs, ogi = po.build_grid(exp, exp6)
    assert np.array_equal(gs, ogs) and np.array_equal(gi, ogi)


# ------------------------------------------------------------------------------------------------ GPU: reference == CUDA

@pytest.mark.gpu
def test_cuda_search_by_projection_frame_equals_reference(kitti):
    from ceres_mono_orb_slam2_b200 import Camera, ORBmatcher
    _, offs, ext, o = kitti
    cam = Camera.create(1241, 376, synth.KITTI_K, o.scale_factors)
    b, K4, sf = cam.bounds6(), cam.K4(), o.scale_factors
    for f in (1, 2):
        (lk, ld), (ck, cd) = ext[f - 1], ext[f]
        shift = (offs[f - 1] - offs[f]).astype(np.float64)
        flags, xw, mdesc = synth.make_last_frame_view(lk, ld, shift, seed=50 + f)
        T = np.array([[1, 1e-4 * f, 0, 0.002], [-1e-4 * f, 1, 0, -0.001], [0, 0, 1, 0.004], [0, 0, 0, 1]]).reshape(-1)
        n, nl = len(ck), len(lk)
        m = ORBmatcher(0.9, True, max_batch=1, max_keypoints=max(n, nl))
        m.set_frames(cam, ck[None], cd[None], np.array([n], np.int32), 1, n)
        stride = max(n, nl)
        lkp = np.zeros((1, stride), lk.dtype); lkp[0, :nl] = lk
        fl = np.zeros((1, stride), np.uint8); fl[0, :nl] = flags
        X = np.zeros((1, stride, 3)); X[0, :nl] = xw
        D = np.zeros((1, stride, 32), np.uint8); D[0, :nl] = mdesc
        match, nm = m.SearchByProjectionFrame(T[None], lkp, np.array([nl], np.int32), fl, X, D, stride, 15.0)
        rm, rnm, _ = pr.search_by_projection_frame(ck, cd, b, K4, sf, T, lk, flags, xw, mdesc, 15.0, True)
        assert nm[0] == rnm and np.array_equal(match[0, :n], rm)


@pytest.mark.gpu
def test_cuda_search_by_projection_points_equals_reference(scene):
    from ceres_mono_orb_slam2_b200 import Camera, ORBmatcher
    _, offs, ext, o = scene
    (lk, ld), (ck, cd) = ext
    cam = Camera.create(640, 480, synth.TUM2_K, o.scale_factors)
    b, sf = cam.bounds6(), o.scale_factors
    shift = (offs[0] - offs[1]).astype(np.float64)
    v = points_view(lk, ld, shift, seed=9)
    n, npnt = len(ck), len(v[0])
    m = ORBmatcher(0.8, True, max_batch=1, max_keypoints=n, max_points=npnt)
    m.set_frames(cam, ck[None], cd[None], np.array([n], np.int32), 1, n)
    assign, nm = m.SearchByProjectionPoints(np.array([npnt], np.int32), v[0][None], v[1][None], v[2][None], v[3][None], v[4][None],
                                            v[5][None], npnt, 3.0)
    ra, rnm, _ = pr.search_by_projection_points(ck, cd, b, sf, *v, 3.0, 0.8)
    assert nm[0] == rnm and np.array_equal(assign[0, :n], ra)


# ------------------------------------------------------------------------------------------------ keyframe searches (a11-a15)

from tests import kf_cases as kc                      # noqa: E402
from tests.matcher_scenarios import make_two_views    # noqa: E402


@pytest.fixture(scope="module", params=["sparse", "crowded"])
def two_views(request):
    if request.param == "sparse":
        return make_two_views(n=1200, seed=0)
    # crowded: few distinct descriptors, many ties and long claim chains (the GPU tests' second scenario)
    S = make_two_views(n=500, seed=5, width=400, height=300, n_extra=100, flip_bits=2)
    rng = np.random.default_rng(7)
    base = rng.integers(0, 256, (6, 32)).astype(np.uint8)
    for key in ("desc1", "desc2", "mp_desc"):
        S[key] = base[rng.integers(0, 6, len(S[key]))]
    return S


def test_reference_reloc_projection(two_views):
    S = two_views
    c = kc.case_reloc(S)
    om, onm, ohp = kc.oracle_reloc(S, c)
    rm, rnm, rhp = kc.ref_reloc(S, c)
    assert rnm == onm and np.array_equal(rm, om) and np.array_equal(rhp, ohp)


def test_reference_sim3_projection(two_views):
    S = two_views
    c = kc.case_points(S)
    oa, onm, om = kc.oracle_proj_sim3(S, c)
    ra, rnm, rm = kc.ref_proj_sim3(S, c)
    assert rnm == onm and np.array_equal(ra, oa) and np.array_equal(rm, om)


@pytest.mark.parametrize("sim3", [0, 1])
def test_reference_fuse(two_views, sim3):
    """The reference mutates the map inside Fuse; the oracle / CUDA path return per-point decisions that the adapter applies.
    Comparable without re-implementing Replace: the return value, and for every point the reference ADDED to the keyframe
    (first to decide for a free keypoint) the keypoint index equals the oracle's decision for it."""
    S = two_views
    c = kc.case_points(S)
    order = c["order"]
    first = np.zeros(len(order), bool)
    seen = set()
    for i, p in enumerate(order):           # the case repeats points; the reference would see the SAME MapPoint* twice
        if p not in seen:
            first[i] = True; seen.add(p)
    c = {k: (v[first] if isinstance(v, np.ndarray) and len(v) == len(order) else v) for k, v in c.items()}
    obi, obd, onf = kc.oracle_fuse(S, c, sim3)
    added, repl, rnf = kc.ref_fuse(S, c, sim3)
    assert rnf == onf
    if sim3 == 0:
        got = added >= 0
        assert got.sum() > 0 and np.array_equal(added[got], obi[got])
        # every oracle decision is either an addition or a replacement of the point that sits there
        dec = np.nonzero(obi >= 0)[0]
        owner = {int(added[p]): p for p in np.nonzero(got)[0]}
        assert all(int(obi[p]) in owner for p in dec)
    else:
        # Fuse(KF, Scw, ...) adds the first decider of a free keypoint and records later ones in vpReplacePoint (:938-947)
        got = added >= 0
        assert got.sum() > 0 and np.array_equal(added[got], obi[got])
        owner = {int(added[p]): p for p in np.nonzero(got)[0]}
        for p in np.nonzero((obi >= 0) & ~got)[0]:
            assert repl[p] == owner[int(obi[p])], "vpReplacePoint names the point that already sits at the decided keypoint"


def test_reference_search_by_sim3(two_views):
    S = two_views
    c = kc.sim3_consistent_case(kc.case_sim3(S), len(S["k2"]))
    om, onf = kc.oracle_sim3(S, c)
    rm, rnf = kc.ref_sim3(S, c)
    assert rnf == onf and np.array_equal(rm, om) and onf > 5


@pytest.mark.parametrize("mode", [0, 1])
def test_reference_search_by_bow(two_views, mode):
    S = two_views
    c = kc.case_bow(S)
    om, onm = kc.oracle_bow(S, c, mode)
    rm, rnm = kc.ref_bow(S, c, mode)
    assert rnm == onm and np.array_equal(rm, om) and onm > 5


def test_reference_search_for_triangulation(two_views):
    S = two_views
    c = kc.case_triangulation(S)
    om, onm = kc.oracle_triangulation(S, c)
    rm, rnm = kc.ref_triangulation(S, c)
    assert rnm == onm and np.array_equal(rm, om) and onm > 5


def test_reference_search_for_initialization(two_views):
    S = two_views
    c = kc.case_init(S)
    om, onm, oprev = kc.oracle_init(S, c)
    rm, rnm, rprev = kc.ref_init(S, c)
    assert rnm == onm and np.array_equal(rm, om) and np.array_equal(rprev, oprev)


def test_reference_map_point_maintenance():
    """MapPoint::ComputeDistinctiveDescriptors and UpdateNormalAndDepth of the reference on real MapPoint / KeyFrame objects
    against the oracle.  min / max distance as the reference stores them (GetMin/MaxDistanceInvariance
    apply 0.8 / 1.2 on read)."""
    import ctypes as C
    from oracle.pyoracle import _p
    from tests.matcher_scenarios import make_map_observations
    M = make_map_observations(n_points=1500, n_keyframes=60, seed=7, max_obs=40)
    rng = np.random.default_rng(3)
    st = M["start"]; obs_kf = M["obs_kf"].copy(); ref_kf = M["ref_kf"].copy()
    for p in range(len(st) - 1):                   # a std::map<KeyFrame*, size_t> holds each keyframe once, in address order
        n = st[p + 1] - st[p]
        if n:
            obs_kf[st[p]:st[p + 1]] = np.sort(rng.choice(60, n, replace=False))
            ref_kf[p] = obs_kf[st[p] + rng.integers(0, n)]
    sf = np.empty(8, np.float32); sf[0] = 1
    for i in range(1, 8):
        sf[i] = np.float32(np.float64(sf[i - 1]) * np.float64(np.float32(1.2)))
    obest = po.distinctive_descriptors(st, M["desc"])
    on, omn, omx = po.update_normal_and_depth(st, obs_kf, M["Ow"], M["pos"], ref_kf, M["ref_level"], sf, M["normal0"], M["min0"], M["max0"])
    npnt = len(st) - 1
    out_desc = np.zeros((npnt, 32), np.uint8)
    nr = np.ascontiguousarray(M["normal0"], np.float64).copy(); mn = M["min0"].copy(); mx = M["max0"].copy()
    b6 = np.array([0, 1241, 0, 376, 64 / 1241, 48 / 376], np.float32); K4 = np.array(synth.KITTI_K, np.float32)
    okf = np.ascontiguousarray(obs_kf, np.int32); od = np.ascontiguousarray(M["desc"], np.uint8)
    Ow = np.ascontiguousarray(M["Ow"], np.float64); pos = np.ascontiguousarray(M["pos"], np.float64)
    rk = np.ascontiguousarray(ref_kf, np.int32); rl = np.ascontiguousarray(M["ref_level"], np.int32)
    pr.lib().ref_map_point_maintenance(npnt, _p(st), _p(okf), _p(od), 60, _p(Ow), _p(pos), _p(rk), _p(rl), _p(sf), 8, _p(b6), _p(K4),
                                       _p(out_desc), _p(nr), _p(mn), _p(mx))
    has = obest >= 0
    assert has.sum() > 1000
    assert np.array_equal(out_desc[has], M["desc"][st[:-1][has] + obest[has]])
    assert not out_desc[~has].any()
    assert np.array_equal(nr, on) and np.array_equal(mn, omn) and np.array_equal(mx, omx)

"""GPU parity tests of the remaining ORBmatcher searches through the C ABI
(cmos_kfmatch_*) against oracle/matcher2_oracle.cpp: index-exact matches, counts and in/out flags."""
import numpy as np
import pytest

from tests import kf_cases as kc
from tests.matcher_scenarios import make_two_views

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    return make_two_views(n=1500, seed=11, n_extra=400)


@pytest.fixture(scope="module")
def S_dense():
    """Small image, many keypoints, few distinct descriptors: crowded windows, ties and long claim chains."""
    S = make_two_views(n=1400, seed=12, n_extra=300, width=400, height=300, flip_bits=3)
    rng = np.random.default_rng(3)
    pal = rng.integers(0, 256, (6, 32)).astype(np.uint8)
    for key, n in (("desc1", len(S["k1"])), ("desc2", len(S["k2"])), ("mp_desc", len(S["k1"]))):
        d = pal[rng.integers(0, 6, n)].copy()
        d[np.arange(n), rng.integers(0, 32, n)] ^= (1 << rng.integers(0, 8, n)).astype(np.uint8) * (rng.random(n) < 0.5)
        S[key] = d
    return S


@pytest.fixture(scope="module")
def matcher():
    from ceres_mono_orb_slam2_b200 import KeyFrameMatcher
    m = KeyFrameMatcher(0.75, True, max_keypoints=4096, max_points=4096, max_nodes=1024)
    yield m
    m.close()


@pytest.mark.parametrize("scn", ["S", "S_dense"])
@pytest.mark.parametrize("check_ori,th,orb_dist", [(True, 15.0, 100), (False, 10.0, 64)])
def test_search_by_projection_reloc(request, matcher, scn, check_ori, th, orb_dist):
    S = request.getfixturevalue(scn)
    c = kc.case_reloc(S, th=th, orb_dist=orb_dist)
    om, onm, ohp = kc.oracle_reloc(S, c, check_ori)
    matcher.check_ori = check_ori
    m, nm, hp = kc.gpu_reloc(matcher, S, c)
    matcher.check_ori = True
    assert nm == onm and np.array_equal(m, om) and np.array_equal(hp[:len(om)], ohp[:len(om)])
    assert onm > (200 if scn == "S" else 50)


@pytest.mark.parametrize("scn", ["S", "S_dense"])
def test_search_by_projection_sim3(request, matcher, scn):
    S = request.getfixturevalue(scn)
    c = kc.case_points(S)
    oa, onm, omt = kc.oracle_proj_sim3(S, c)
    a, nm, mt = kc.gpu_proj_sim3(matcher, S, c)
    assert nm == onm and np.array_equal(a, oa) and np.array_equal(mt[:len(oa)], omt[:len(oa)])
    assert onm > (200 if scn == "S" else 50)


@pytest.mark.parametrize("scn", ["S", "S_dense"])
@pytest.mark.parametrize("sim3", [0, 1])
def test_fuse(request, matcher, scn, sim3):
    S = request.getfixturevalue(scn)
    c = kc.case_points(S)
    obi, obd, onf = kc.oracle_fuse(S, c, sim3)
    bi, bd, nf = kc.gpu_fuse(matcher, S, c, sim3)
    assert nf == onf and np.array_equal(bi, obi) and np.array_equal(bd, obd)
    assert onf > (100 if scn == "S" else 20)


@pytest.mark.parametrize("scn,s12", [("S", 1.0), ("S_dense", 1.0), ("S", 1.03)])
def test_search_by_sim3(request, matcher, scn, s12):
    S = request.getfixturevalue(scn)
    c = kc.case_sim3(S, s12=s12)
    om, onf = kc.oracle_sim3(S, c)
    m, nf = kc.gpu_sim3(matcher, S, c)
    assert nf == onf and np.array_equal(m, om)
    assert onf > (50 if scn == "S" else 5)


@pytest.mark.parametrize("scn", ["S", "S_dense"])
@pytest.mark.parametrize("mode", [0, 1])
def test_search_by_bow(request, matcher, scn, mode):
    S = request.getfixturevalue(scn)
    c = kc.case_bow(S)
    om, onm = kc.oracle_bow(S, c, mode, nn_ratio=matcher.nnratio)
    m, nm = kc.gpu_bow(matcher, S, c, mode)
    assert nm == onm and np.array_equal(m, om)
    if scn == "S":
        assert onm > 300


@pytest.mark.parametrize("scn", ["S", "S_dense"])
def test_search_for_triangulation(request, matcher, scn):
    S = request.getfixturevalue(scn)
    c = kc.case_triangulation(S)
    om, onm = kc.oracle_triangulation(S, c)
    m, nm = kc.gpu_triangulation(matcher, S, c)
    assert nm == onm and np.array_equal(m, om)
    if scn == "S":
        assert onm > 100


@pytest.mark.parametrize("scn,window", [("S", 100), ("S_dense", 60), ("S", 20)])
def test_search_for_initialization(request, matcher, scn, window):
    S = request.getfixturevalue(scn)
    c = kc.case_init(S)
    matcher.nnratio = 0.9
    om, onm, oprev = kc.oracle_init(S, c, window=window, nn_ratio=0.9)
    m, nm, prev = kc.gpu_init(matcher, S, c, window=window)
    matcher.nnratio = 0.75
    assert nm == onm and np.array_equal(m, om) and np.array_equal(prev, oprev)
    assert onm > (20 if scn == "S" else 3)


def test_empty_inputs_and_errors(matcher):
    from ceres_mono_orb_slam2_b200 import CmosError, FeatureVector, KP_DTYPE
    S = make_two_views(n=50, seed=2, n_extra=5)
    cam = kc.gpu_camera(S)
    matcher.set_view(0, cam, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8), True)
    matcher.set_view(1, cam, S["k2"], S["desc2"], True)
    fv0 = FeatureVector.create([], [0], [])
    fv2 = FeatureVector.create(*__import__("oracle.pyoracle", fromlist=["x"]).flatten_feature_vector(S["node2"]))
    m, nm = matcher.SearchByBoW(1, np.zeros(0, np.uint8), fv0, np.ones(len(S["k2"]), np.uint8), fv2)
    assert nm == 0 and len(m) == 0
    bi, bd, nf = matcher.Fuse(np.concatenate([np.eye(3).reshape(-1), np.zeros(6)]), np.zeros(0, np.uint8), np.zeros((0, 3)),
                              np.zeros((0, 3)), np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros((0, 32), np.uint8),
                              3.0, inv_level_sigma2=np.ones(8, np.float32))
    assert nf == 0
    with pytest.raises(CmosError):        # node ids must ascend
        matcher.SearchByBoW(1, np.zeros(0, np.uint8), FeatureVector.create([5, 3], [0, 0, 0], []),
                            np.ones(len(S["k2"]), np.uint8), fv2)


def test_map_point_maintenance():
    """cmos_map_distinctive_descriptors / cmos_map_update_normal_and_depth == the oracle, bit for bit."""
    from ceres_mono_orb_slam2_b200 import MapPointOps
    from oracle import pyoracle as po
    from tests.matcher_scenarios import make_map_observations
    M = make_map_observations(n_points=4000, n_keyframes=60, seed=7, max_obs=40)
    ops = MapPointOps(max_points=4000, max_observations=len(M["desc"]), max_keyframes=60)
    best, desc = ops.ComputeDistinctiveDescriptors(M["start"], M["desc"])
    obest = po.distinctive_descriptors(M["start"], M["desc"])
    assert np.array_equal(best, obest)
    has = obest >= 0
    assert np.array_equal(desc[has], M["desc"][M["start"][:-1][has] + obest[has]])
    sf = np.empty(8, np.float32); sf[0] = 1
    for i in range(1, 8):
        sf[i] = np.float32(np.float64(sf[i - 1]) * np.float64(np.float32(1.2)))
    got = ops.UpdateNormalAndDepth(M["start"], M["obs_kf"], M["Ow"], M["pos"], M["ref_kf"], M["ref_level"], sf, M["normal0"],
                                   M["min0"], M["max0"])
    ref = po.update_normal_and_depth(M["start"], M["obs_kf"], M["Ow"], M["pos"], M["ref_kf"], M["ref_level"], sf, M["normal0"],
                                     M["min0"], M["max0"])
    for g, r in zip(got, ref):
        assert np.array_equal(g, r)
    # empty batch, and a point list with only empty points
    b, _ = ops.ComputeDistinctiveDescriptors(np.zeros(4, np.int32), np.zeros((0, 32), np.uint8))
    assert list(b) == [-1, -1, -1]
    ops.close()


@pytest.mark.parametrize("k,L,levelsup,n", [(10, 4, 2, 2000), (10, 3, 4, 500), (6, 5, 3, 3000), (33, 2, 1, 100)])
def test_vocabulary_transform(k, L, levelsup, n):
    """cmos_voc_transform == DBoW2's transform as restated by the oracle: word ids, L1-normalised values (bit for bit) and
    the feature vector; stop words, repeated words, levelsup beyond the depth (node 0) and branching > 32 included."""
    from ceres_mono_orb_slam2_b200 import ORBVocabulary
    from oracle import pyoracle as po
    voc = po.make_vocabulary(k=k, L=L, seed=k * 10 + L, stop_frac=0.03)
    rng = np.random.default_rng(n)
    leaves = np.nonzero(voc["word"] >= 0)[0]
    f = voc["desc"][rng.choice(leaves, n)].copy()
    for _ in range(3):
        f[np.arange(n), rng.integers(0, 32, n)] ^= (1 << rng.integers(0, 8, n)).astype(np.uint8)
    f[: n // 10] = f[n // 10: 2 * (n // 10)]
    f[-5:] = rng.integers(0, 256, (5, 32)).astype(np.uint8)
    V = ORBVocabulary(voc["child_start"], voc["children"], voc["desc"], voc["weight"], voc["word"], L)
    got = V.transform(f, levelsup)
    ref = po.bow_transform(voc, f, levelsup)
    for key in ("words", "values", "fv_nodes", "fv_start", "fv_features"):
        assert np.array_equal(got[key], ref[key]), key
    assert len(ref["words"]) > 10 and V.launch_count() == 2
    empty = V.transform(np.zeros((0, 32), np.uint8), levelsup)
    assert len(empty["words"]) == 0 and len(empty["fv_nodes"]) == 0
    V.close()

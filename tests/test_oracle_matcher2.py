"""CPU checks of oracle/matcher2_oracle.cpp (the restated relocalisation / loop / BoW / fuse / Sim3 / initialisation /
triangulation searches).  The reference has no tests for these functions (parity unpinned by the reference), so the
restatement is checked against known answers: on a synthetic pair of keyframes with ground-truth correspondences and
random 256-bit descriptors a wrong match is (almost surely) farther than every threshold, hence every match the
oracle returns must be a ground-truth correspondence, recall must be high, and the order-dependent rules (claimed
keypoints, ratio test, rotation histogram, displacement) must show on hand-made inputs."""
import numpy as np
import pytest

from oracle import pyoracle as po
from tests import kf_cases as kc
from tests.matcher_scenarios import make_two_views


@pytest.fixture(scope="module")
def S():
    return make_two_views(n=900, seed=5, n_extra=200)


def _gt_of_kp2(S):
    return S["p2"]          # view-2 keypoint -> view-1 keypoint / map point index (or -1)


def test_reloc_matches_are_ground_truth(S):
    c = kc.case_reloc(S)
    match, nm, has = kc.oracle_reloc(S, c, check_ori=False)
    got = np.nonzero(match >= 0)[0]
    assert nm == len(got) > 300
    assert np.array_equal(match[got], _gt_of_kp2(S)[got])                       # only true correspondences
    assert not np.any(c["cur_has_point"][got]) and np.all(has[got] == 1)        # pre-assigned keypoints are never taken
    assert np.all(c["kf_valid"][match[got]] == 1)
    # with the orientation check the survivors are a subset
    m2, nm2, _ = kc.oracle_reloc(S, c, check_ori=True)
    assert nm2 <= nm and np.all((m2 == match) | (m2 == -1))


def test_projection_sim3_claims_each_keypoint_once(S):
    c = kc.case_points(S)
    assign, nm, matched = kc.oracle_proj_sim3(S, c)
    got = np.nonzero(assign >= 0)[0]
    assert nm == len(got) > 250
    assert np.array_equal(c["order"][assign[got]], _gt_of_kp2(S)[got])
    assert not np.any(c["matched"][got]) and np.all(matched[got] == 1)
    assert not np.any(c["pt_skip"][assign[got]])
    # a duplicated point that comes later never steals the keypoint: the assigned index is the first occurrence
    first = {}
    for i, p in enumerate(c["order"]):
        if not c["pt_skip"][i]:
            first.setdefault(p, i)
    assert all(assign[j] == first[c["order"][assign[j]]] for j in got)


@pytest.mark.parametrize("sim3", [0, 1])
def test_fuse_decisions(S, sim3):
    c = kc.case_points(S)
    bi, bd, nf = kc.oracle_fuse(S, c, sim3)
    got = np.nonzero(bi >= 0)[0]
    assert nf == len(got) > 150
    assert np.array_equal(_gt_of_kp2(S)[bi[got]], c["order"][got])
    assert np.all(bd[got] <= 50) and np.all(bd[bi < 0] == 256) and not np.any(c["pt_skip"][got])
    # decisions do not depend on each other: duplicates of a point get the same keypoint
    for p in np.unique(c["order"][got]):
        assert len(set(bi[got][c["order"][got] == p])) == 1


def test_search_by_sim3_is_mutual(S):
    c = kc.case_sim3(S)
    m12, nf = kc.oracle_sim3(S, c)
    got = np.nonzero(m12 >= 0)[0]
    assert nf == len(got) > 200
    assert np.array_equal(_gt_of_kp2(S)[m12[got]], got)
    assert not np.any(c["side1"][1][got]) and np.all(c["side1"][0][got] == 1)
    assert not np.any(c["side2"][1][m12[got]]) and np.all(c["side2"][0][m12[got]] == 1)


@pytest.mark.parametrize("mode", [0, 1])
def test_search_by_bow(S, mode):
    c = kc.case_bow(S)
    match, nm = kc.oracle_bow(S, c, mode, check_ori=False)
    got = np.nonzero(match >= 0)[0]
    assert nm == len(got) > 300
    if mode == 0:       # match[idx2] = idx1
        assert np.array_equal(match[got], _gt_of_kp2(S)[got]) and np.all(c["valid1"][match[got]] == 1)
    else:               # match[idx1] = idx2
        assert np.array_equal(_gt_of_kp2(S)[match[got]], got) and np.all(c["valid2"][match[got]] == 1)
        assert len(set(match[got])) == len(got)
    # orientation check: flip the angle of some matched keyframe keypoints -> exactly those matches disappear
    S2 = dict(S); k1 = S["k1"].copy()
    idx1 = match[got] if mode == 0 else got
    bad = idx1[:: 7]
    k1["angle"][bad] = (k1["angle"][bad] + np.random.default_rng(0).uniform(60, 300, len(bad))) % 360   # spread over bins
    S2["k1"] = k1
    m2, nm2 = kc.oracle_bow(S2, c, mode, check_ori=True)
    lost = set((match[got] if mode == 0 else got)[(m2[got] < 0)])
    assert set(bad) <= lost and nm2 == (m2 >= 0).sum()


def test_bow_ratio_test_and_claims():
    """Two identical candidates in one node fail the ratio test; a keypoint taken by an earlier feature is not offered
    to a later one."""
    d = np.zeros((4, 32), np.uint8); d[1, 0] = 1; d[2, 0] = 3; d[3, :8] = 255
    fv = po.flatten_feature_vector(np.array([5, 5, 5, 5]))
    ang = np.zeros(4, np.float32)
    # side 1 feature 0 (all zero) against side 2 = {zero, zero}: best == second -> rejected
    m, nm = po.search_by_bow(0, d[[0]], ang[:1], [1], po.flatten_feature_vector(np.array([5])), d[[0, 0]], ang[:2], None,
                             po.flatten_feature_vector(np.array([5, 5])), 0.6, False)
    assert nm == 0 and np.all(m == -1)
    # side 1 = {zero, one-bit}; side 2 = {zero, far}: feature 0 takes keypoint 0, feature 1 finds only `far` -> no match
    m, nm = po.search_by_bow(0, d[[0, 1]], ang[:2], [1, 1], po.flatten_feature_vector(np.array([5, 5])), d[[0, 3]], ang[:2],
                             None, po.flatten_feature_vector(np.array([5, 5])), 0.6, False)
    assert nm == 1 and m[0] == 0 and m[1] == -1


def test_triangulation_respects_epipolar_gate(S):
    c = kc.case_triangulation(S)
    m12, nm = kc.oracle_triangulation(S, c, check_ori=False)
    got = np.nonzero(m12 >= 0)[0]
    assert nm == len(got) > 150
    assert np.array_equal(_gt_of_kp2(S)[m12[got]], got)
    assert not np.any(c["has1"][got]) and not np.any(c["has2"][m12[got]])
    # a wrong fundamental matrix (transposed) kills almost everything
    c2 = dict(c); c2["F12"] = c["F12"].reshape(3, 3).T.reshape(-1).copy()
    _, nm_bad = kc.oracle_triangulation(S, c2, check_ori=False)
    assert nm_bad < nm // 4


def test_search_for_initialization(S):
    c = kc.case_init(S)
    m12, nm, prev = kc.oracle_init(S, c, check_ori=False)
    got = np.nonzero(m12 >= 0)[0]
    assert nm == len(got) > 100
    assert np.all(S["k1"]["octave"][got] == 0) and np.all(S["k2"]["octave"][m12[got]] == 0)
    assert np.array_equal(_gt_of_kp2(S)[m12[got]], got)
    assert len(set(m12[got])) == len(got)                                  # vnMatches21 keeps the map one-to-one
    assert np.array_equal(prev[got], np.stack([S["k2"]["x"][m12[got]], S["k2"]["y"][m12[got]]], 1))
    untouched = np.setdiff1d(np.arange(len(m12)), got)
    assert np.array_equal(prev[untouched], c["prev"][untouched])


def test_initialization_displacement():
    """A later, closer keypoint takes over an already matched F2 keypoint (ORBmatcher.cc:419-429)."""
    from oracle.pyoracle import KP_DTYPE
    k1 = np.zeros(2, KP_DTYPE); k1["x"] = [100, 104]; k1["y"] = [100, 100]
    k2 = np.zeros(1, KP_DTYPE); k2["x"] = [102]; k2["y"] = [100]
    d2 = np.zeros((1, 32), np.uint8)
    d1 = np.zeros((2, 32), np.uint8); d1[0, 0] = 0b111; d1[1, 0] = 0b1       # distances 3 and 1
    b = np.array([0.0, 640, 0.0, 480, np.float32(64) / np.float32(640), np.float32(48) / np.float32(480)], np.float32)
    V2 = po.View(k2, d2, b, np.array([500, 500, 320, 240], np.float32), np.ones(1, np.float32))
    m12, nm, prev = po.search_for_initialization(k1, d1, V2, np.stack([k1["x"], k1["y"]], 1), 50, 0.9, False)
    assert nm == 1 and list(m12) == [-1, 0]
    assert np.array_equal(prev[1], [102, 100]) and np.array_equal(prev[0], [100, 100])


def test_distinctive_descriptor_and_normal_depth_known_answers():
    """MapPoint::ComputeDistinctiveDescriptors / UpdateNormalAndDepth against numpy on small inputs."""
    from tests.matcher_scenarios import make_map_observations
    M = make_map_observations(n_points=200, n_keyframes=12, seed=4, max_obs=12)
    best = po.distinctive_descriptors(M["start"], M["desc"])
    bits = np.unpackbits(M["desc"], axis=1).astype(np.int32)
    for p in range(200):
        o, e = M["start"][p], M["start"][p + 1]
        if e == o:
            assert best[p] == -1
            continue
        D = (bits[o:e, None, :] != bits[None, o:e, :]).sum(2)
        med = np.sort(D, axis=1)[:, int(0.5 * (e - o - 1))]
        assert best[p] == int(np.argmin(med))                       # first minimum
    sf = (1.2 ** np.arange(8)).astype(np.float32)
    nr, mn, mx = po.update_normal_and_depth(M["start"], M["obs_kf"], M["Ow"], M["pos"], M["ref_kf"], M["ref_level"], sf,
                                            M["normal0"], M["min0"], M["max0"])
    for p in range(200):
        o, e = M["start"][p], M["start"][p + 1]
        if e == o:
            assert np.array_equal(nr[p], M["normal0"][p]) and mn[p] == M["min0"][p] and mx[p] == M["max0"][p]
            continue
        d = M["pos"][p] - M["Ow"][M["obs_kf"][o:e]]
        assert np.allclose(nr[p], (d / np.linalg.norm(d, axis=1)[:, None]).mean(0), rtol=1e-12, atol=1e-15)
        dist = np.float32(np.linalg.norm(M["pos"][p] - M["Ow"][M["ref_kf"][p]]))
        assert mx[p] == dist * sf[M["ref_level"][p]] and mn[p] == mx[p] / sf[7]


def test_bow_transform_against_brute_force():
    """DBoW2 transform restated (oracle/bow_oracle.cpp) vs a brute-force numpy descent on a synthetic vocabulary."""
    voc = po.make_vocabulary(k=7, L=3, seed=3, stop_frac=0.05)
    rng = np.random.default_rng(5)
    leaves = np.nonzero(voc["word"] >= 0)[0]
    f = voc["desc"][rng.choice(leaves, 400)].copy()
    f[np.arange(400), rng.integers(0, 32, 400)] ^= (1 << rng.integers(0, 8, 400)).astype(np.uint8)
    f[:40] = f[40:80]                                             # repeated words
    r = po.bow_transform(voc, f, levelsup=2)
    bits = np.unpackbits(voc["desc"], axis=1).astype(np.int16)
    fb = np.unpackbits(f, axis=1).astype(np.int16)
    words, nodes, weights = [], [], []
    for i in range(400):
        node, level, nid = 0, 0, 0
        while voc["child_start"][node + 1] > voc["child_start"][node]:
            level += 1
            ch = voc["children"][voc["child_start"][node]:voc["child_start"][node + 1]]
            node = int(ch[np.argmin((bits[ch] != fb[i]).sum(1))])     # argmin = first minimum
            if level == voc["L"] - 2:
                nid = node
        words.append(voc["word"][node]); nodes.append(nid); weights.append(voc["weight"][node])
    words, nodes, weights = np.array(words), np.array(nodes), np.array(weights)
    assert np.array_equal(r["feat_word"], words) and np.array_equal(r["feat_node"], nodes)
    keep = weights > 0
    uw = np.unique(words[keep])
    assert np.array_equal(r["words"], uw)
    val = np.array([weights[keep][words[keep] == w].sum() for w in uw])
    assert np.allclose(r["values"], val / val.sum(), rtol=1e-13) and abs(r["values"].sum() - 1) < 1e-12
    un = np.unique(nodes[keep])
    assert np.array_equal(r["fv_nodes"], un)
    for j, nd in enumerate(un):
        assert np.array_equal(r["fv_features"][r["fv_start"][j]:r["fv_start"][j + 1]], np.nonzero(keep & (nodes == nd))[0])

"""Parity pinned to the REFERENCE ITSELF: oracle/_ref/libref.so is the reference's own, unmodified
`src/ORBextractor.cc` and vendored `lib/DBoW2` compiled against a minimal OpenCV stand-in (oracle/ref_shim/, recipe
oracle/ref_shim/Makefile; the five OpenCV image primitives delegate to the cv2-pinned restatements).

CPU tests (`-m "not gpu"`): `_ref` == oracle/orb_oracle.cpp and `_ref` DBoW2 == oracle/bow_oracle.cpp, bit for bit, on the
BASELINE configs[0] / configs[1] frames and on edge cases.  GPU tests: `_ref` == the CUDA path through the C ABI.
With these the chain is  reference source == oracle == CUDA  for SURVEY.md §8 rows a1-a7 and 8f-2.
"""
import numpy as np
import pytest

from ceres_mono_orb_slam2_b200 import synth
from oracle import pyoracle as po
from oracle import pyref as pr

pytestmark = pytest.mark.skipif(not pr.available(), reason="oracle/_ref/libref.so not built (needs /root/reference once)")

CONFIG0 = (640, 480, 1000)     # BASELINE.json configs[0]: TUM2.yaml
CONFIG1 = (1241, 376, 2000)    # BASELINE.json configs[1]: KITTI00-02.yaml


def _bow_features(voc, n, seed):
    rng = np.random.default_rng(seed)
    leaves = np.nonzero(voc["word"] >= 0)[0]
    f = voc["desc"][rng.choice(leaves, n)].copy()
    for _ in range(3):
        f[np.arange(n), rng.integers(0, 32, n)] ^= (1 << rng.integers(0, 8, n)).astype(np.uint8)
    f[: n // 10] = f[n // 10: 2 * (n // 10)]          # repeated words
    f[-5:] = rng.integers(0, 256, (5, 32)).astype(np.uint8)
    return f


# ------------------------------------------------------------------------------------------------ CPU: _ref == oracle

@pytest.mark.parametrize("w,h,nfeat,seed", [CONFIG0 + (11,), CONFIG0 + (12,), CONFIG1 + (1000,), CONFIG1 + (3,),
                                             (320, 240, 500, 1), (752, 480, 1200, 5)])
def test_reference_extractor_equals_oracle(w, h, nfeat, seed):
    img = synth.make_image(w, h, seed)
    ref = pr.RefOrbExtractor(nfeat, 1.2, 8, 20, 7); ora = po.OrbOracle(nfeat, 1.2, 8, 20, 7)
    assert np.array_equal(ref.scale_factors, ora.scale_factors) and np.array_equal(ref.inv_scale_factors, ora.inv_scale_factors)
    assert np.array_equal(ref.sigma2, ora.sigma2) and np.array_equal(ref.inv_sigma2, ora.inv_sigma2)
    rk, rd = ref.extract(img); ok, od = ora.extract(img)
    for l in range(8):
        assert np.array_equal(ref.level_image(l), ora.level_image(l)), f"pyramid level {l} (incl. border)"
    assert len(rk) == len(ok) >= nfeat * 0.9
    for field in rk.dtype.names:
        assert np.array_equal(rk[field], ok[field]), field
    assert np.array_equal(rd, od)


def test_reference_extractor_sequence_and_reuse():
    """One extractor object over a moving sequence (the arena rewinds per call, the pyramid buffers are reallocated)."""
    frames = synth.make_sequence(*CONFIG1[:2], 4, seed=21)
    ref = pr.RefOrbExtractor(CONFIG1[2]); ora = po.OrbOracle(CONFIG1[2])
    for f in frames:
        rk, rd = ref.extract(f); ok, od = ora.extract(f)
        assert np.array_equal(rk, ok) and np.array_equal(rd, od)


@pytest.mark.parametrize("kind", ["flat", "noise", "checker", "dark_low_contrast"])
def test_reference_extractor_edge_images(kind):
    """Empty result (the reference releases the descriptor matrix, ORBextractor.cc:1064), the minThFAST fallback
    cells, and dense ties in the quadtree."""
    w, h = 400, 300
    rng = np.random.default_rng(5)
    if kind == "flat":
        img = np.full((h, w), 128, np.uint8)
    elif kind == "noise":
        img = rng.integers(0, 256, (h, w)).astype(np.uint8)
    elif kind == "checker":
        yy, xx = np.mgrid[0:h, 0:w]
        img = (((xx // 9 + yy // 9) & 1) * 200 + 20).astype(np.uint8)
    else:
        img = (synth.make_image(w, h, 9).astype(np.int32) // 12 + 30).astype(np.uint8)
    ref = pr.RefOrbExtractor(800); ora = po.OrbOracle(800)
    rk, rd = ref.extract(img); ok, od = ora.extract(img)
    assert len(rk) == len(ok)
    assert np.array_equal(rk, ok) and np.array_equal(rd, od)
    if kind == "flat":
        assert len(rk) == 0


@pytest.mark.parametrize("k,L,levelsup,n", [(10, 4, 2, 2000), (10, 3, 4, 500), (6, 5, 3, 3000), (10, 6, 4, 1500), (20, 2, 0, 300)])
def test_reference_dbow2_equals_oracle(k, L, levelsup, n):
    """ORBvoc.txt's own shape is k=10, L=6, levelsup=4 (Frame.cc:326); loadFromTextFile limits k <= 20, L <= 10."""
    voc = po.make_vocabulary(k=k, L=L, seed=k * 10 + L, stop_frac=0.03)
    V = pr.RefVocabulary(voc)
    assert V.size() == int((voc["word"] >= 0).sum())
    f = _bow_features(voc, n, n)
    ref = V.transform(f, levelsup); ora = po.bow_transform(voc, f, levelsup)
    for key in ("words", "values", "fv_nodes", "fv_start", "fv_features"):
        assert np.array_equal(ref[key], ora[key]), key
    assert len(ref["words"]) > 10
    g = _bow_features(voc, n, n + 1)
    s = V.score(ref, V.transform(g, levelsup))               # L1 score of two different frames, in (0, 1)
    assert 0.0 <= s < 1.0 and abs(V.score(ref, ref) - 1.0) < 1e-12
    e = V.transform(np.zeros((0, 32), np.uint8), levelsup)
    assert len(e["words"]) == 0 and len(e["fv_nodes"]) == 0


# ------------------------------------------------------------------------------------------------ GPU: _ref == CUDA

@pytest.mark.gpu
@pytest.mark.parametrize("w,h,nfeat,seed", [CONFIG0 + (11,), CONFIG1 + (1000,)])
def test_cuda_extractor_equals_reference(w, h, nfeat, seed):
    from ceres_mono_orb_slam2_b200 import ORBextractor
    img = synth.make_image(w, h, seed)
    ext = ORBextractor(nfeat, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    ref = pr.RefOrbExtractor(nfeat, 1.2, 8, 20, 7)
    assert np.array_equal(ext.GetScaleFactors(), ref.scale_factors)
    assert np.array_equal(ext.GetInverseScaleSigmaSquares(), ref.inv_sigma2)
    kps, desc = ext(img)
    rk, rd = ref.extract(img)
    for l in range(8):
        assert np.array_equal(ext.debug_level_image(0, l), ref.level_image(l)), f"pyramid level {l}"
    assert len(kps) == len(rk)
    for field in rk.dtype.names:
        assert np.array_equal(kps[field], rk[field]), field
    assert np.array_equal(desc, rd)


@pytest.mark.gpu
def test_cuda_extractor_batch_equals_reference():
    """Eight frames of the configs[1] batch, extracted in one batched launch sequence, each against the reference."""
    from ceres_mono_orb_slam2_b200 import ORBextractor
    frames = synth.make_sequence(*CONFIG1[:2], 8, seed=64)
    ext = ORBextractor(CONFIG1[2], 1.2, 8, 20, 7, max_width=CONFIG1[0], max_height=CONFIG1[1], max_batch=8)
    kps, desc, counts = ext.extract_batch(frames)
    ref = pr.RefOrbExtractor(CONFIG1[2])
    for f in range(8):
        rk, rd = ref.extract(frames[f])
        n = int(counts[f])
        assert n == len(rk) and np.array_equal(kps[f, :n], rk) and np.array_equal(desc[f, :n], rd), f"frame {f}"


@pytest.mark.gpu
@pytest.mark.parametrize("k,L,levelsup,n", [(10, 4, 2, 2000), (10, 6, 4, 2000), (6, 5, 3, 3000)])
def test_cuda_vocabulary_equals_reference(k, L, levelsup, n):
    from ceres_mono_orb_slam2_b200 import ORBVocabulary
    voc = po.make_vocabulary(k=k, L=L, seed=k * 10 + L, stop_frac=0.03)
    f = _bow_features(voc, n, n)
    ref = pr.RefVocabulary(voc).transform(f, levelsup)
    V = ORBVocabulary(voc["child_start"], voc["children"], voc["desc"], voc["weight"], voc["word"], L)
    got = V.transform(f, levelsup)
    for key in ("words", "values", "fv_nodes", "fv_start", "fv_features"):
        assert np.array_equal(got[key], ref[key]), key
    V.close()

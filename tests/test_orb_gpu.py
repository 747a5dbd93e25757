"""GPU parity tests of the ORB extractor: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded frames.  Bit-exact: pyramid bytes, FAST candidate sets, blurred bytes, keypoint records (index order,
coordinates, angle, response, octave) and all 256 descriptor bits."""
import ctypes as C
import os

import numpy as np
import pytest

from ceres_mono_orb_slam2_b200 import ORBextractor, synth, _lib
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _dump(name, **arrays):
    os.makedirs("gpurun_out", exist_ok=True)
    np.savez_compressed(os.path.join("gpurun_out", name), **arrays)


def _compare_frame(ext, oracle, img, frame, tag):
    okps, odesc = oracle.extract(img)
    for l in range(oracle.nlevels):
        a = ext.debug_level_image(frame, l); b = oracle.level_image(l)
        assert a.shape == b.shape, (tag, l, a.shape, b.shape)
        nbad = int((a != b).sum())
        assert nbad == 0, f"{tag}: pyramid level {l}: {nbad} bytes differ"
        gc = ext.debug_level_candidates(frame, l)
        oc = oracle.level_candidates(l)
        ocs = np.stack([oc["x"], oc["y"], oc["response"]], 1).astype(np.int32)
        ocs = ocs[np.lexsort((ocs[:, 0], ocs[:, 1]))]
        if not np.array_equal(gc, ocs):
            _dump(f"cand_{tag}_{l}", gpu=gc, oracle=ocs)
            raise AssertionError(f"{tag}: FAST candidates level {l}: gpu {len(gc)} vs oracle {len(ocs)}")
        if len(oracle.level_keypoints(l)) == 0:
            continue   # the reference (and the oracle) only blur levels that hold keypoints (:1079-1086)
        gb = ext.debug_level_blurred(frame, l); ob = oracle.level_blurred(l)
        nbad = int((gb != ob).sum())
        assert nbad == 0, f"{tag}: blurred level {l}: {nbad} bytes differ"
    return okps, odesc


@pytest.mark.parametrize("w,h,nfeat,seed", [(640, 480, 1000, 11), (1241, 376, 2000, 1000)])
def test_extract_single_frame_bit_exact(w, h, nfeat, seed):
    img = synth.make_image(w, h, seed)
    ext = ORBextractor(nfeat, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    oracle = po.OrbOracle(nfeat, 1.2, 8, 20, 7)
    assert np.array_equal(ext.features_per_level, oracle.quota)
    assert np.array_equal(ext.GetScaleFactors(), oracle.scale_factors)
    assert np.array_equal(ext.GetInverseScaleSigmaSquares(), oracle.inv_sigma2)
    kps, desc = ext(img)
    okps, odesc = _compare_frame(ext, oracle, img, 0, f"{w}x{h}")
    if len(kps) != len(okps) or not np.array_equal(kps, okps):
        _dump(f"kps_{w}x{h}", gpu=kps, oracle=okps)
    assert len(kps) == len(okps)
    for field in ("octave", "x", "y", "response", "size", "angle", "class_id"):
        bad = np.nonzero(kps[field] != okps[field])[0]
        assert bad.size == 0, f"{field}: {bad.size} keypoints differ, first {bad[:5]}: {kps[bad[:5]]} vs {okps[bad[:5]]}"
    bad = np.nonzero((desc != odesc).any(1))[0]
    assert bad.size == 0, f"{bad.size} descriptors differ, first {bad[:5]}"


def test_extract_batch_matches_per_frame():
    frames = synth.make_sequence(1241, 376, 6, seed=21)
    ext = ORBextractor(2000, 1.2, 8, 20, 7, max_width=1241, max_height=376, max_batch=6)
    oracle = po.OrbOracle(2000, 1.2, 8, 20, 7)
    kps, desc, counts = ext.extract_batch(frames)
    for f in range(len(frames)):
        okps, odesc = oracle.extract(frames[f])
        n = int(counts[f])
        assert n == len(okps), f"frame {f}: {n} vs {len(okps)}"
        assert np.array_equal(kps[f, :n], okps), f"frame {f} keypoints"
        assert np.array_equal(desc[f, :n], odesc), f"frame {f} descriptors"
    # partial batch on the same handle, different content
    k2, d2, c2 = ext.extract_batch(frames[3:5])
    assert np.array_equal(c2, counts[3:5]) and np.array_equal(k2[0, :c2[0]], kps[3, :counts[3]])


@pytest.mark.parametrize("kind", ["flat", "noise", "checker", "small"])
def test_extract_edge_cases(kind):
    rng = np.random.default_rng(5)
    if kind == "flat":
        img = np.full((376, 1241), 77, np.uint8)                    # no corners anywhere -> 0 keypoints
    elif kind == "noise":
        img = rng.integers(0, 256, (376, 1241), dtype=np.uint8)     # every cell saturated with candidates
    elif kind == "checker":
        yy, xx = np.mgrid[0:376, 0:1241]
        img = (((xx // 4 + yy // 4) & 1) * 200 + 20).astype(np.uint8)   # many equal scores -> NMS ties
    else:
        img = synth.make_image(160, 120, 9, n_rect=40, n_blob=20)   # top levels have a single cell row
    h, w = img.shape
    ext = ORBextractor(1000, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
    oracle = po.OrbOracle(1000, 1.2, 8, 20, 7)
    kps, desc = ext(img)
    okps, odesc = _compare_frame(ext, oracle, img, 0, kind)
    assert len(kps) == len(okps)
    assert np.array_equal(kps, okps) and np.array_equal(desc, odesc)
    if kind == "flat":
        assert len(kps) == 0


def test_other_parameters_and_resize_of_handle():
    """Different nfeatures / levels / scale on one handle sized for the larger image."""
    ext = ORBextractor(1500, 1.3, 5, 25, 9, max_width=800, max_height=600, max_batch=2)
    oracle = po.OrbOracle(1500, 1.3, 5, 25, 9)
    for (w, h, seed) in [(800, 600, 1), (752, 480, 2)]:
        img = synth.make_image(w, h, seed)
        kps, desc = ext(img)
        okps, odesc = oracle.extract(img)
        assert np.array_equal(kps, okps) and np.array_equal(desc, odesc), (w, h)


@pytest.mark.parametrize("scale,nlevels", [(1.25, 6), (1.5, 5), (2.0, 4)])
def test_scale_factors_pick_a_resize_tile_that_holds_the_source_rows(scale, nlevels):
    """k_resize2's shared-memory tile holds a fixed number of source rows: 64 output rows fit up to scale 1.22, 32 up to 1.25,
    16 up to 1.4, beyond that the handle falls back to the per-row kernel.  Every level byte and the result stay exact."""
    w, h = 752, 480
    img = synth.make_image(w, h, 31)
    ext = ORBextractor(1200, scale, nlevels, 20, 7, max_width=w, max_height=h, max_batch=1)
    oracle = po.OrbOracle(1200, scale, nlevels, 20, 7)
    kps, desc = ext(img)
    okps, odesc = _compare_frame(ext, oracle, img, 0, f"scale{scale}")
    assert np.array_equal(kps, okps) and np.array_equal(desc, odesc)


@pytest.mark.parametrize("w,h,nfeat,scale,nlevels,ini,mn", [
    (160, 120, 200, 1.2, 4, 20, 7),        # tiny image
    (641, 479, 1000, 1.2, 8, 20, 7),       # odd sizes: every level has partial words / partial cells
    (1920, 1080, 3000, 1.2, 8, 20, 7),     # large cells (the general FAST tile), 3 x the KITTI pixel count
    (640, 480, 5000, 1.2, 8, 12, 5),       # more features asked for than most levels can give
    (640, 480, 50, 1.2, 8, 20, 7),         # quotas of one to a dozen keypoints per level
    (640, 480, 1000, 1.2, 1, 20, 7),       # a single level: no resize at all
    (752, 480, 1000, 1.1, 8, 20, 7),       # fine pyramid
    (640, 480, 1000, 1.2, 8, 7, 7),        # iniThFAST == minThFAST: the fallback pass repeats the first one
    (320, 240, 1000, 1.2, 12, 20, 7),      # twelve levels, the last one 43 x 32
])
def test_parameter_sweep_bit_exact(w, h, nfeat, scale, nlevels, ini, mn):
    """Unusual but legal parameter sets (the reference reads all of them from the settings file, ORBextractor.cc:410-470): every
    pyramid byte, FAST candidate, blurred byte, keypoint and descriptor bit equals the oracle."""
    img = synth.make_image(w, h, 77 + w + nlevels)
    ext = ORBextractor(nfeat, scale, nlevels, ini, mn, max_width=w, max_height=h, max_batch=1)
    oracle = po.OrbOracle(nfeat, scale, nlevels, ini, mn)
    assert np.array_equal(ext.features_per_level, oracle.quota)
    kps, desc = ext(img)
    okps, odesc = _compare_frame(ext, oracle, img, 0, f"sweep{w}x{h}_{nfeat}_{scale}_{nlevels}_{ini}_{mn}")
    assert len(kps) == len(okps) and np.array_equal(kps, okps) and np.array_equal(desc, odesc)


def test_empty_image_and_errors():
    ext = ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=1)
    kps, desc = ext(np.zeros((0, 0), np.uint8))
    assert len(kps) == 0 and desc.shape == (0, 32)
    with pytest.raises(_lib.CmosError):
        ext(np.zeros((481, 640), np.uint8))          # larger than the handle
    with pytest.raises(_lib.CmosError):
        ext.extract_batch(np.zeros((2, 480, 640), np.uint8))   # batch over max_batch


def test_device_float_helpers_bit_exact():
    L = _lib.lib()
    deg = np.concatenate([np.linspace(0, 360, 400001, dtype=np.float32),
                          np.random.default_rng(2).uniform(0, 360, 1000000).astype(np.float32)])
    c = np.zeros_like(deg); s = np.zeros_like(deg)
    _lib.check(L.cmos_debug_sincos_deg(_lib.ptr(deg), _lib.ptr(c), _lib.ptr(s), deg.size))
    oc, os_ = po.sincos_deg(deg)
    assert np.array_equal(c, oc) and np.array_equal(s, os_)
    rng = np.random.default_rng(3)
    y = rng.integers(-200000, 200000, 1000000).astype(np.float32)
    x = rng.integers(-200000, 200000, 1000000).astype(np.float32)
    y[:4] = [0, 0, 3, -3]; x[:4] = [0, 3, 0, 0]
    out = np.zeros_like(y)
    _lib.check(L.cmos_debug_fast_atan2(_lib.ptr(y), _lib.ptr(x), _lib.ptr(out), y.size))
    assert np.array_equal(out, po.fast_atan2(y, x))

"""CPU tests: pin the C++ ORB oracle (oracle/orb_oracle.cpp) to OpenCV 4.13 via cv2, primitive by
primitive and stage by stage, on seeded synthetic frames."""
import numpy as np
import pytest

from ceres_mono_orb_slam2_b200 import synth
from oracle import pyoracle as po

cv2 = pytest.importorskip("cv2")
cv2.setNumThreads(1)


@pytest.mark.parametrize("shape,dst", [((376, 1241), (1034, 313)), ((480, 640), (533, 400)), ((105, 346), (288, 88)),
                                       ((61, 97), (81, 51))])
def test_resize_matches_cv2(shape, dst):
    rng = np.random.default_rng(shape[0])
    src = rng.integers(0, 256, shape, dtype=np.uint8)
    ref = cv2.resize(src, dst, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(ref, po.resize(src, dst[0], dst[1]))


@pytest.mark.parametrize("shape", [(376, 1241), (105, 346), (40, 57), (8, 9)])
def test_blur_matches_cv2(shape):
    rng = np.random.default_rng(shape[1])
    src = rng.integers(0, 256, shape, dtype=np.uint8)
    ref = cv2.GaussianBlur(src, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(ref, po.blur7(src))


@pytest.mark.parametrize("threshold", [20, 7])
@pytest.mark.parametrize("shape", [(38, 37), (46, 38), (200, 300), (7, 7), (6, 30), (12, 9)])
def test_fast_matches_cv2(shape, threshold):
    img = synth.make_image(max(shape[1], 64), max(shape[0], 64), seed=shape[0] * 100 + threshold,
                           n_rect=60, n_blob=30)[:shape[0], :shape[1]]
    img = np.ascontiguousarray(img)
    det = cv2.FastFeatureDetector_create(threshold, True)
    ref = np.array([(kp.pt[0], kp.pt[1], kp.response) for kp in det.detect(img)], np.float32).reshape(-1, 3)
    got = po.fast(img, threshold).astype(np.float32)
    assert np.array_equal(ref, got)


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(7)
    y = rng.integers(-70000, 70000, 20000).astype(np.float32)
    x = rng.integers(-70000, 70000, 20000).astype(np.float32)
    y[:4] = [0, 0, 5, -5]; x[:4] = [0, 5, 0, 0]
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    assert np.array_equal(ref, po.fast_atan2(y, x))


def test_constructor_tables():
    o = po.OrbOracle(2000, 1.2, 8, 20, 7)
    assert o.quota.tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert o.umax.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    o1 = po.OrbOracle(1000, 1.2, 8, 20, 7)
    assert o1.quota.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert abs(float(o.scale_factors[7]) - 3.583182) < 1e-5


@pytest.mark.parametrize("w,h,nfeat,seed", [(640, 480, 1000, 11), (1241, 376, 2000, 1000)])
def test_pipeline_stages_match_cv2(w, h, nfeat, seed):
    """Pyramid, per-cell FAST candidates (incl. 20->7 fallback and order) and blurred levels of the C++
    oracle equal what the reference's own cv2 calls produce."""
    img = synth.make_image(w, h, seed)
    o = po.OrbOracle(nfeat, 1.2, 8, 20, 7)
    kps, desc = o.extract(img)
    pyr = po.cv2_pyramid(img, o.inv_scale_factors)
    total = 0
    for l in range(8):
        assert np.array_equal(o.level_image(l), pyr[l]), f"pyramid level {l}"
        c_cv = po.cv2_level_candidates(pyr[l])
        c_or = o.level_candidates(l)
        assert len(c_cv) == len(c_or)
        assert np.array_equal(c_cv[:, 0], c_or["x"]) and np.array_equal(c_cv[:, 1], c_or["y"])
        assert np.array_equal(c_cv[:, 2], c_or["response"])
        assert len(c_or) >= 3 * o.quota[l], "synthetic frame too poor in corners"
        bl = cv2.GaussianBlur(np.ascontiguousarray(pyr[l][19:-19, 19:-19]), (7, 7), 2, 2,
                              borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(bl, o.level_blurred(l))
        nl = len(o.level_keypoints(l))
        assert o.quota[l] <= nl <= o.quota[l] + 2
        total += nl
    assert total == len(kps) and desc.shape == (total, 32)
    # keypoints: inside the image minus the 19-px edge at their level, angle in [0,360)
    assert (kps["angle"] >= 0).all() and (kps["angle"] < 360).all()
    assert (np.diff(kps["octave"]) >= 0).all()
    lvl = kps["octave"]
    x = kps["x"] / o.scale_factors[lvl]; y = kps["y"] / o.scale_factors[lvl]
    for l in range(8):
        lw, lh = o.level_size(l)
        m = lvl == l
        assert (x[m] > 18.5).all() and (x[m] < lw - 18.5).all() and (y[m] > 18.5).all() and (y[m] < lh - 18.5).all()


def test_fallback_threshold_is_exercised():
    """The synthetic frames contain cells that are empty at iniThFAST but not at minThFAST."""
    img = synth.make_image(1241, 376, 1000)
    o = po.OrbOracle(2000, 1.2, 8, 20, 7)
    o.extract(img)
    c = o.level_candidates(0)
    assert (c["response"] < 20).any(), "no candidate came from the minThFAST fallback"


def test_descriptor_orientation_cv2_cross_check():
    """IC angle of the oracle equals cv2.fastAtan2 of independently computed patch moments, and
    descriptor bits equal an independent numpy evaluation of the rotated pattern on cv2's blur."""
    from oracle.pattern import PATTERN
    img = synth.make_image(640, 480, 5)
    o = po.OrbOracle(500, 1.2, 4, 20, 7)
    kps, desc = o.extract(img)
    umax = o.umax
    rng = np.random.default_rng(0)
    for i in rng.choice(len(kps), 60, replace=False):
        kp = kps[i]; l = int(kp["octave"])
        lim = o.level_image(l).astype(np.int64)
        s = o.scale_factors[l] if l else np.float32(1.0)
        # recover level coordinates exactly as integers
        cx = int(round(float(kp["x"]) / float(s))); cy = int(round(float(kp["y"]) / float(s)))
        m10 = m01 = 0
        for v in range(-15, 16):
            d = umax[abs(v)]
            for u in range(-d, d + 1):
                I = lim[cy + 19 + v, cx + 19 + u]
                m10 += u * I; m01 += v * I
        assert np.float32(cv2.fastAtan2(float(m01), float(m10))) == kp["angle"]
        blur = cv2.GaussianBlur(np.ascontiguousarray(o.level_image(l)[19:-19, 19:-19]), (7, 7), 2, 2,
                                borderType=cv2.BORDER_REFLECT_101)
        ang = np.float32(kp["angle"]) * np.float32(np.pi / np.float32(180.0))
        c, s_ = po.sincos_deg(np.array([kp["angle"]], np.float32))
        a, b = c[0], s_[0]
        bits = np.zeros(32, np.uint8)
        for byte in range(32):
            val = 0
            for k in range(8):
                p0 = PATTERN[byte * 16 + 2 * k]; p1 = PATTERN[byte * 16 + 2 * k + 1]
                def px(p):
                    ry = int(np.rint(np.float32(np.float32(p[0]) * b) + np.float32(np.float32(p[1]) * a)))
                    rx = int(np.rint(np.float32(np.float32(p[0]) * a) - np.float32(np.float32(p[1]) * b)))
                    return int(blur[cy + ry, cx + rx])
                val |= (px(p0) < px(p1)) << k
            bits[byte] = val
        assert np.array_equal(bits, desc[i])


def test_sincos_restatement_matches_libm():
    """The glibc sincosf algorithm restated for the device (csrc/orb_device.cuh) equals libm's cosf/sinf
    on the host for a dense sample of angles; the exhaustive [0, 6.3] sweep is documented in DESIGN.md."""
    from oracle.sincosf_restated import sincosf_restated
    deg = np.concatenate([np.linspace(0, 360, 200001, dtype=np.float32),
                          np.random.default_rng(1).uniform(0, 360, 300000).astype(np.float32)])
    c, s = po.sincos_deg(deg)
    c2, s2 = sincosf_restated(deg * np.float32(np.pi / np.float32(180.0)))
    assert np.array_equal(c, c2) and np.array_equal(s, s2)

"""TEST INFRASTRUCTURE ONLY — ctypes front for oracle/_ref/libref.so: the reference's OWN, unmodified
`src/ORBextractor.cc` and vendored `lib/DBoW2` compiled against the OpenCV stand-in of oracle/ref_shim/
(recipe: oracle/ref_shim/Makefile).  Used by tests/test_ref_parity.py to pin oracle/orb_oracle.cpp,
oracle/bow_oracle.cpp and the CUDA path to the reference itself, and by bench.py's CPU legs
(`cpu_baseline.kind == "reference"`).  Never imported by the product package.

/root/reference exists only in the build container; on the GPU box the prebuilt library (git-ignored, shipped with
the snapshot) is loaded as is.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from .pyoracle import KP_DTYPE, _p, build as _build_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref.so")
_LIB = None


def available() -> bool:
    return os.path.exists(_SO) or os.path.exists("/root/reference/src/ORBextractor.cc")


def build() -> str:
    """(Re)build when the reference tree is present; otherwise use the prebuilt library."""
    _build_oracle()
    if os.path.exists("/root/reference/src/ORBextractor.cc"):
        r = subprocess.run(["make", "-C", os.path.join(_HERE, "ref_shim")], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle/_ref build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    if not os.path.exists(_SO):
        raise FileNotFoundError(_SO + " (built from /root/reference by oracle/ref_shim/Makefile)")
    return _SO


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ref_orb_create.restype = C.c_void_p
        L.ref_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.ref_orb_destroy.argtypes = [C.c_void_p]
        L.ref_bow_load_text.restype = C.c_void_p
        L.ref_bow_load_text.argtypes = [C.c_char_p]
        L.ref_bow_destroy.argtypes = [C.c_void_p]
        L.ref_bow_score.restype = C.c_double
        for name in ("ref_orb_tables", "ref_orb_extract", "ref_orb_result", "ref_orb_level_size",
                     "ref_orb_level_image_bordered", "ref_bow_transform", "ref_bow_size", "ref_bow_depth",
                     "ref_bow_branching", "ref_bow_score"):
            getattr(L, name).argtypes = None
        _LIB = L
    return _LIB


class RefOrbExtractor:
    """ORB_SLAM2::ORBextractor of the reference (include/ORBextractor.h:46-61), unmodified."""

    def __init__(self, nfeatures=2000, scale=1.2, nlevels=8, ini_th=20, min_th=7):
        self.L = lib()
        self.nlevels = nlevels
        self.h = C.c_void_p(self.L.ref_orb_create(nfeatures, C.c_float(scale), nlevels, ini_th, min_th))
        sf, isf, s2, is2 = (np.zeros(nlevels, np.float32) for _ in range(4))
        self.L.ref_orb_tables(self.h, _p(sf), _p(isf), _p(s2), _p(is2))
        self.scale_factors, self.inv_scale_factors, self.sigma2, self.inv_sigma2 = sf, isf, s2, is2

    def __del__(self):
        try:
            self.L.ref_orb_destroy(self.h)
        except Exception:
            pass

    def extract(self, img: np.ndarray):
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        n = self.L.ref_orb_extract(self.h, _p(img), w, h, w)
        kps = np.zeros(n, KP_DTYPE); desc = np.zeros((n, 32), np.uint8)
        self.L.ref_orb_result(self.h, _p(kps), _p(desc))
        return kps, desc

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        self.L.ref_orb_level_size(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_image(self, l):
        """Pyramid level with its 19-pixel border, like OrbOracle.level_image."""
        w, h = self.level_size(l)
        out = np.zeros((h + 38, w + 38), np.uint8)
        self.L.ref_orb_level_image_bordered(self.h, l, _p(out))
        return out


def write_vocabulary_text(voc: dict, path: str) -> None:
    """Flattened tree (oracle.pyoracle.make_vocabulary layout) -> the ORBvoc.txt format that
    TemplatedVocabulary::loadFromTextFile reads (TemplatedVocabulary.h:1338-1421): header `k L scoring weighting`,
    then one line per node in id order: `parent isLeaf d0..d31 weight`.  No trailing newline: the reference's
    `while(!f.eof())` loop would turn an empty last line into a node with an uninitialised descriptor under the root."""
    cs = voc["child_start"]; parent = np.full(len(cs) - 1, -1, np.int64)
    for p in range(len(cs) - 1):
        kids = voc["children"][cs[p]:cs[p + 1]]
        parent[kids] = p
        assert list(kids) == sorted(kids), "text format stores children in id order"
    lines = [f"{int(voc['k'])} {int(voc['L'])} 0 0"]      # L1_NORM, TF_IDF — ORBvoc.txt's own header values
    for i in range(1, len(cs) - 1):
        assert 0 <= parent[i] < i, "a parent must precede its children in the file"
        leaf = int(cs[i + 1] == cs[i])
        lines.append(f"{parent[i]} {leaf} " + " ".join(str(int(b)) for b in voc["desc"][i]) + f" {float(voc['weight'][i])!r}")
    with open(path, "w") as f:
        f.write("\n".join(lines))


class RefVocabulary:
    """DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (include/ORBVocabulary.h:30-31), loaded with loadFromTextFile."""

    def __init__(self, voc: dict):
        self.L = lib()
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
            path = f.name
        try:
            write_vocabulary_text(voc, path)
            h = self.L.ref_bow_load_text(path.encode())
        finally:
            os.unlink(path)
        if not h:
            raise RuntimeError("loadFromTextFile rejected the vocabulary")
        self.h = C.c_void_p(h)

    def __del__(self):
        try:
            self.L.ref_bow_destroy(self.h)
        except Exception:
            pass

    def size(self):
        return self.L.ref_bow_size(self.h)

    def transform(self, features, levelsup=4):
        f = np.ascontiguousarray(features, np.uint8); n = len(f)
        bw = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1)); fn = np.zeros(max(n, 1), np.int32)
        fs = np.zeros(n + 2, np.int32); ff = np.zeros(max(n, 1), np.int32); nfn = C.c_int32()
        nw = self.L.ref_bow_transform(self.h, _p(f), n, int(levelsup), _p(bw), _p(bv), _p(fn), _p(fs), _p(ff), C.byref(nfn))
        m = nfn.value
        return dict(words=bw[:nw], values=bv[:nw], fv_nodes=fn[:m], fv_start=fs[:m + 1], fv_features=ff[:fs[m]])

    def score(self, a, b):
        w1 = np.ascontiguousarray(a["words"], np.int32); v1 = np.ascontiguousarray(a["values"], np.float64)
        w2 = np.ascontiguousarray(b["words"], np.int32); v2 = np.ascontiguousarray(b["values"], np.float64)
        return float(self.L.ref_bow_score(self.h, _p(w1), _p(v1), len(w1), _p(w2), _p(v2), len(w2)))


# ---------------------------------------------------------------------------------------------------
# The reference's own Frame / KeyFrame / MapPoint / ORBmatcher (oracle/ref_shim/ref_matcher.cpp).  Same argument meaning as
# the matching functions of oracle.pyoracle, so tests can run the two side by side.

def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def frame_construct(img, nfeatures, K4, dist, scale=1.2, nlevels=8, ini_th=20, min_th=7):
    """Frame::Frame(imgGray, ...) of the reference: returns (keypoints, undistorted keypoints, descriptors, bounds6,
    grid_start, grid_idx)."""
    img = np.ascontiguousarray(img, np.uint8); h, w = img.shape
    cap = 4 * nfeatures + 64
    kps = np.zeros(cap, KP_DTYPE); un = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
    b6 = np.zeros(6, np.float32); gs = np.zeros(64 * 48 + 1, np.int32); gi = np.zeros(cap, np.int32)
    k = _f32(K4); d = _f32(dist)
    n = lib().ref_frame_construct(_p(img), w, h, w, int(nfeatures), C.c_float(scale), int(nlevels), int(ini_th), int(min_th),
                                  _p(k), _p(d), len(d), cap, _p(kps), _p(un), _p(desc), _p(b6), _p(gs), _p(gi))
    assert n <= cap
    return kps[:n], un[:n], desc[:n], b6, gs, gi[:gs[-1]]


def build_grid(kps, bounds6, scale_factors):
    kps = np.ascontiguousarray(kps); sf = _f32(scale_factors); b = _f32(bounds6)
    gs = np.zeros(64 * 48 + 1, np.int32); gi = np.zeros(max(len(kps), 1), np.int32)
    n = lib().ref_build_grid(_p(kps), len(kps), _p(b), _p(sf), len(sf), _p(gs), _p(gi))
    return gs, gi[:n]


def features_in_area(kps, bounds6, scale_factors, x, y, r, min_level, max_level):
    kps = np.ascontiguousarray(kps); sf = _f32(scale_factors); b = _f32(bounds6)
    out = np.zeros(max(len(kps), 1), np.int32)
    n = lib().ref_features_in_area(_p(kps), len(kps), _p(b), _p(sf), len(sf), C.c_float(x), C.c_float(y), C.c_float(r),
                                   int(min_level), int(max_level), _p(out))
    return out[:n]


def search_by_projection_frame(cur_kps, cur_desc, bounds6, K4, scale_factors, Tcw, last_kps, last_flags, last_xw, last_desc,
                               th, check_ori, claimed=None, nn_ratio=0.9):
    cur_kps = np.ascontiguousarray(cur_kps); cur_desc = np.ascontiguousarray(cur_desc, np.uint8)
    last_kps = np.ascontiguousarray(last_kps); n = len(cur_kps)
    claimed = np.zeros(max(n, 1), np.uint8) if claimed is None else claimed
    match = np.full(max(n, 1), -1, np.int32)
    b = _f32(bounds6); k = _f32(K4); sf = _f32(scale_factors); T = np.ascontiguousarray(Tcw, np.float64)
    lf = np.ascontiguousarray(last_flags, np.uint8); lx = np.ascontiguousarray(last_xw, np.float64)
    ld = np.ascontiguousarray(last_desc, np.uint8)
    nm = lib().ref_search_by_projection_frame(_p(cur_kps), _p(cur_desc), n, _p(b), _p(k), _p(sf), len(sf), _p(T), _p(last_kps),
                                              len(last_kps), _p(lf), _p(lx), _p(ld), C.c_float(th), int(check_ori),
                                              C.c_float(nn_ratio), _p(claimed), _p(match))
    return match[:n], nm, claimed[:n]


def search_by_projection_points(kps, desc, bounds6, scale_factors, in_view, level, view_cos, proj_xy, mp_desc, has_obs, th,
                                nn_ratio, claimed=None):
    kps = np.ascontiguousarray(kps); desc = np.ascontiguousarray(desc, np.uint8); n = len(kps)
    claimed = np.zeros(max(n, 1), np.uint8) if claimed is None else claimed
    assign = np.full(max(n, 1), -1, np.int32)
    b = _f32(bounds6); sf = _f32(scale_factors)
    iv = np.ascontiguousarray(in_view, np.uint8); lv = np.ascontiguousarray(level, np.int32)
    vc = _f32(view_cos); pj = _f32(proj_xy); md = np.ascontiguousarray(mp_desc, np.uint8)
    ho = np.ascontiguousarray(has_obs, np.uint8)
    nm = lib().ref_search_by_projection_points(_p(kps), _p(desc), n, _p(b), _p(sf), len(sf), len(iv), _p(iv), _p(lv), _p(vc),
                                               _p(pj), _p(md), _p(ho), C.c_float(th), C.c_float(nn_ratio), _p(claimed), _p(assign))
    return assign[:n], nm, claimed[:n]


def is_in_frustum(pose15, K4, bounds4, scale_factors, cos_limit, xw, normal, min_dist, max_dist):
    n = len(xw)
    pose15 = np.ascontiguousarray(pose15, np.float64); k = _f32(K4); b = _f32(bounds4); sf = _f32(scale_factors)
    xw = np.ascontiguousarray(xw, np.float64); normal = np.ascontiguousarray(normal, np.float64)
    mn = _f32(min_dist); mx = _f32(max_dist)
    in_view = np.zeros(n, np.uint8); proj = np.zeros((n, 2), np.float32); level = np.zeros(n, np.int32); vcos = np.zeros(n, np.float32)
    lib().ref_is_in_frustum(_p(pose15), _p(k), _p(b), _p(sf), len(sf), C.c_float(cos_limit), n, _p(xw), _p(normal), _p(mn),
                            _p(mx), _p(in_view), _p(proj), _p(level), _p(vcos))
    return in_view, proj, level, vcos


def descriptor_distance(a, b):
    return lib().ref_descriptor_distance(_p(np.ascontiguousarray(a, np.uint8)), _p(np.ascontiguousarray(b, np.uint8)))

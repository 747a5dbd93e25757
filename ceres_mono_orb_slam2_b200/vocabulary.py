"""Host-side mirror of ORBVocabulary::transform (DBoW2 TemplatedVocabulary, lib/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1260)
as Frame::ComputeBoW / KeyFrame::ComputeBoW call it, over the C ABI (cmos_voc_* in include/cmos_b200.h).  The tree is
uploaded once, flattened; all compute is in libcmos_b200.so."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


def _c(a, dt):
    return np.ascontiguousarray(a, dt)


class ORBVocabulary:
    def __init__(self, child_start, children, node_descriptors, node_weights, node_word_ids, depth_levels: int, device: int = 0):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        a = (_c(child_start, np.int32), _c(children, np.int32), _c(node_descriptors, np.uint8), _c(node_weights, np.float64),
             _c(node_word_ids, np.int32))
        check(self._L.cmos_voc_create(len(a[0]) - 1, ptr(a[0]), ptr(a[1]), ptr(a[2]), ptr(a[3]), ptr(a[4]), int(depth_levels),
                                      device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.cmos_voc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform(self, descriptors, levelsup: int = 4):
        """-> dict(words, values, fv_nodes, fv_start, fv_features): BowVector and FeatureVector, flattened."""
        d = _c(descriptors, np.uint8); n = len(d)
        bw = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1)); fn = np.zeros(max(n, 1), np.int32)
        fs = np.zeros(n + 1, np.int32); ff = np.zeros(max(n, 1), np.int32)
        nw, nf = C.c_int32(), C.c_int32()
        check(self._L.cmos_voc_transform(self._h, ptr(d), n, int(levelsup), ptr(bw), ptr(bv), C.byref(nw), ptr(fn), ptr(fs),
                                         ptr(ff), C.byref(nf)))
        m = nf.value
        return dict(words=bw[:nw.value], values=bv[:nw.value], fv_nodes=fn[:m], fv_start=fs[:m + 1], fv_features=ff[:fs[m]])

    def launch_count(self) -> int:
        n = C.c_int32()
        check(self._L.cmos_voc_last_launch_count(self._h, C.byref(n)))
        return n.value

"""Host-side mirror of the reference's ORBextractor class (include/ORBextractor.h:46-111) over the C ABI.

Same constructor arguments, same getters, same call semantics (mask ignored, empty image -> nothing),
plus a batched call because the B200 path processes many frames per launch.  All compute happens in
libcmos_b200.so; this file only marshals buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, OrbParams, check, ptr


class ORBextractor:
    HARRIS_SCORE = 0
    FAST_SCORE = 1

    def __init__(self, nfeatures: int, scaleFactor: float, nlevels: int, iniThFAST: int, minThFAST: int,
                 max_width: int = 1241, max_height: int = 376, max_batch: int = 1, device: int = 0):
        self._L = _lib.lib()
        p = OrbParams(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_width, max_height, max_batch,
                      device)
        self._h = C.c_void_p()
        check(self._L.cmos_orb_create(C.byref(p), C.byref(self._h)))
        self.nlevels = nlevels
        self.max_batch = max_batch
        cap = C.c_int32()
        check(self._L.cmos_orb_keypoint_capacity(self._h, C.byref(cap)))
        self.capacity = cap.value
        sf, isf, s2, is2 = (np.zeros(nlevels, np.float32) for _ in range(4))
        check(self._L.cmos_orb_get_scale_factors(self._h, ptr(sf), ptr(isf), ptr(s2), ptr(is2)))
        self._sf, self._isf, self._s2, self._is2 = sf, isf, s2, is2
        self._scale = float(np.float32(scaleFactor))
        q = np.zeros(nlevels, np.int32)
        check(self._L.cmos_orb_get_features_per_level(self._h, ptr(q)))
        self.features_per_level = q

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.cmos_orb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- getters, ORBextractor.h:63-83 ----
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self._scale
    def GetScaleFactors(self): return self._sf.copy()
    def GetInverseScaleFactors(self): return self._isf.copy()
    def GetScaleSigmaSquares(self): return self._s2.copy()
    def GetInverseScaleSigmaSquares(self): return self._is2.copy()

    # ---- operator() ----
    def __call__(self, image: np.ndarray, mask=None):
        """One CV_8UC1 image -> (keypoints[KP_DTYPE], descriptors[N,32] uint8).  `mask` is ignored like the
        reference (ORBextractor.h:58)."""
        if image is None or image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        kps, desc, counts = self.extract_batch(image[None])
        n = int(counts[0])
        return kps[0, :n].copy(), desc[0, :n].copy()

    def extract_batch(self, images: np.ndarray, out=None):
        """images [B,H,W] uint8 (host) -> keypoints [B,cap], descriptors [B,cap,32], counts [B]."""
        assert images.dtype == np.uint8 and images.ndim == 3
        images = np.ascontiguousarray(images)
        B, H, W = images.shape
        if out is None:
            kps = np.zeros((B, self.capacity), KP_DTYPE)
            desc = np.zeros((B, self.capacity, 32), np.uint8)
            counts = np.zeros(B, np.int32)
        else:
            kps, desc, counts = out
        check(self._L.cmos_orb_extract(self._h, ptr(images), C.c_int64(H * W), W, W, H, B, ptr(kps), ptr(desc),
                                       ptr(counts), self.capacity))
        return kps, desc, counts

    def extract_device(self, d_images, frame_stride: int, pitch: int, width: int, height: int, n_frames: int,
                       stream: int = 0):
        """Images already on the device (torch uint8 tensor or raw pointer).  Asynchronous."""
        check(self._L.cmos_orb_extract_device(self._h, ptr(d_images), C.c_int64(frame_stride), pitch, width, height,
                                              n_frames, C.c_void_p(stream)))

    def download(self, n_frames: int, stream: int = 0, out=None):
        if out is None:
            kps = np.zeros((n_frames, self.capacity), KP_DTYPE)
            desc = np.zeros((n_frames, self.capacity, 32), np.uint8)
            counts = np.zeros(n_frames, np.int32)
        else:
            kps, desc, counts = out
        check(self._L.cmos_orb_download(self._h, n_frames, ptr(kps), ptr(desc), ptr(counts), self.capacity,
                                        C.c_void_p(stream)))
        return kps, desc, counts

    def device_results(self):
        """Raw device pointers (ints): keypoints, descriptors, counts, level_counts, capacity."""
        a, b, c, d = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        cap = C.c_int32()
        check(self._L.cmos_orb_device_results(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(cap)))
        return a.value, b.value, c.value, d.value, cap.value

    def launch_count(self) -> int:
        n = C.c_int32()
        check(self._L.cmos_orb_last_launch_count(self._h, C.byref(n)))
        return n.value

    def set_profiling(self, enable: bool = True):
        check(self._L.cmos_orb_set_profiling(self._h, int(enable)))

    def stage_times(self):
        """-> (dict stage name -> accumulated ms, number of extractions covered)."""
        ms = (C.c_double * 8)()
        n, calls = C.c_int32(), C.c_int64()
        check(self._L.cmos_orb_stage_times(self._h, ms, 8, C.byref(n), C.byref(calls)))
        names = ["pyramid", "fast", "quadtree", "blur", "describe"]
        return {names[i]: ms[i] for i in range(n.value)}, calls.value

    # ---- verification taps ----
    def level_size(self, level: int):
        w, h = C.c_int32(), C.c_int32()
        check(self._L.cmos_orb_level_size(self._h, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    def debug_level_image(self, frame: int, level: int) -> np.ndarray:
        w, h = self.level_size(level)
        out = np.zeros((h + 38, w + 38), np.uint8)
        check(self._L.cmos_orb_debug_level_image(self._h, frame, level, ptr(out)))
        return out

    def debug_level_blurred(self, frame: int, level: int) -> np.ndarray:
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        check(self._L.cmos_orb_debug_level_blurred(self._h, frame, level, ptr(out)))
        return out

    def debug_level_candidates(self, frame: int, level: int) -> np.ndarray:
        """-> int array [n,3] of (x, y, score), coordinates relative to minBorder, sorted by (y, x)."""
        n = C.c_int32()
        check(self._L.cmos_orb_debug_level_candidates(self._h, frame, level, None, 0, C.byref(n)))
        raw = np.zeros(max(n.value, 1), np.uint32)
        check(self._L.cmos_orb_debug_level_candidates(self._h, frame, level, ptr(raw), raw.size, C.byref(n)))
        raw = raw[:n.value]
        out = np.stack([raw & 0xfff, (raw >> 12) & 0xfff, raw >> 24], 1).astype(np.int32)
        return out[np.lexsort((out[:, 0], out[:, 1]))]

"""Host-side mirror of the per-frame work of the reference's Tracking thread — Frame::Frame (ExtractORB,
AssignFeaturesToGrid; src/Frame.cc:98-156) followed by ORBmatcher::SearchByProjection(current, last, th)
(src/Tracking.cc:632) — as ONE batched call over host buffers (cmos_track_* in include/cmos_b200.h).  The library
pipelines chunks of the batch over CUDA streams; this file only marshals buffers."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, OrbParams, check, ptr
from .orb_matcher import Camera


LAST_POINT_DTYPE = np.dtype([("descriptor", "u1", (32,)), ("xw", "<f8", (3,)), ("angle", "<f4"), ("index", "<u2"), ("octave", "i1"),
                             ("flags", "u1")])
assert LAST_POINT_DTYPE.itemsize == 64      # == sizeof(cmos_last_point)


def pack_last_points(last_keypoints, last_counts, last_flags, last_xw, last_descriptors):
    """Per-keypoint last-frame arrays [B,S]... -> (records, point_start): one cmos_last_point per keypoint with flag bit 0 set,
    frame after frame, in increasing keypoint index (what a caller would build straight from LastFrame.map_points_)."""
    B = len(last_counts)
    recs, start = [], [0]
    for f in range(B):
        n = int(last_counts[f])
        idx = np.nonzero(last_flags[f, :n] & 1)[0]
        r = np.zeros(len(idx), LAST_POINT_DTYPE)
        r["descriptor"] = last_descriptors[f, idx]; r["xw"] = last_xw[f, idx]; r["angle"] = last_keypoints["angle"][f, idx]
        r["index"] = idx; r["octave"] = last_keypoints["octave"][f, idx]; r["flags"] = last_flags[f, idx]
        recs.append(r); start.append(start[-1] + len(idx))
    return np.concatenate(recs) if recs else np.zeros(0, LAST_POINT_DTYPE), np.array(start, np.int32)


ASSOC_DTYPE = np.dtype([("slot", "<i4"), ("angle", "<f4"), ("index", "<u2"), ("octave", "i1"), ("flags", "u1")])
assert ASSOC_DTYPE.itemsize == 12           # == sizeof(cmos_track_assoc)


def pack_map_associations(last_keypoints, last_counts, last_flags, slots):
    """Per-keypoint last-frame arrays [B,S] + the map-point slot of every keypoint (slots [B,S], read where flag bit 0 is set) ->
    (records, assoc_start): one cmos_track_assoc per usable keypoint, frame after frame, in increasing keypoint index."""
    B = len(last_counts)
    recs, start = [], [0]
    for f in range(B):
        n = int(last_counts[f])
        idx = np.nonzero(last_flags[f, :n] & 1)[0]
        r = np.zeros(len(idx), ASSOC_DTYPE)
        r["slot"] = slots[f, idx]; r["angle"] = last_keypoints["angle"][f, idx]; r["index"] = idx
        r["octave"] = last_keypoints["octave"][f, idx]; r["flags"] = last_flags[f, idx]
        recs.append(r); start.append(start[-1] + len(idx))
    return np.concatenate(recs) if recs else np.zeros(0, ASSOC_DTYPE), np.array(start, np.int32)


class TrackParams(C.Structure):
    _fields_ = [("orb", OrbParams), ("lanes", C.c_int32), ("chunk_frames", C.c_int32)]


class TrackingFrontEnd:
    def __init__(self, cam: Camera, nfeatures: int, scaleFactor: float, nlevels: int, iniThFAST: int, minThFAST: int,
                 max_width: int = 1241, max_height: int = 376, lanes: int = 3, chunk_frames: int = 8, device: int = 0,
                 checkOri: bool = True):
        self._L = _lib.lib()
        p = TrackParams(OrbParams(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_width, max_height,
                                  chunk_frames, device), lanes, chunk_frames)
        self._h = C.c_void_p()
        self._cam = cam
        check(self._L.cmos_track_create(C.byref(p), C.byref(cam), C.byref(self._h)))
        cap = C.c_int32()
        check(self._L.cmos_track_keypoint_capacity(self._h, C.byref(cap)))
        self.capacity = cap.value
        self.check_ori = bool(checkOri)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.cmos_track_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def track(self, images, Tcw, last_keypoints, last_counts, last_flags, last_xw, last_descriptors, th: float,
              out=None):
        """images [B,H,W] uint8; Tcw [B,16]; last_* [B,S]...; -> (keypoints [B,cap], descriptors [B,cap,32],
        counts [B], match [B,cap], nmatches [B]).  All host arrays (numpy or pinned torch tensors)."""
        B, H, W = images.shape
        S = last_flags.shape[1]
        if out is None:
            cap = self.capacity
            out = (np.zeros((B, cap), KP_DTYPE), np.zeros((B, cap, 32), np.uint8), np.zeros(B, np.int32),
                   np.full((B, cap), -1, np.int32), np.zeros(B, np.int32))
        kps, desc, counts, match, nm = out
        cap = match.shape[1]
        check(self._L.cmos_track_frames(self._h, ptr(images), C.c_int64(H * W), W, W, H, B, ptr(Tcw), ptr(last_keypoints),
                                        ptr(last_counts), ptr(last_flags), ptr(last_xw), ptr(last_descriptors), S,
                                        C.c_float(th), int(self.check_ori), ptr(kps), ptr(desc), ptr(counts), cap,
                                        ptr(match), ptr(nm)))
        return out

    def submit(self, images, Tcw, last_keypoints, last_counts, last_flags, last_xw, last_descriptors, th: float, out):
        """Asynchronous half of track(): enqueues the batch and returns a ticket at once (cmos_track_submit).  Every buffer
        (inputs and `out`) must stay valid and untouched until wait(ticket) returns; up to 4 batches may be in flight."""
        B, H, W = images.shape
        S = last_flags.shape[1]
        kps, desc, counts, match, nm = out
        cap = match.shape[1]
        t = C.c_int64(-1)
        check(self._L.cmos_track_submit(self._h, ptr(images), C.c_int64(H * W), W, W, H, B, ptr(Tcw), ptr(last_keypoints),
                                        ptr(last_counts), ptr(last_flags), ptr(last_xw), ptr(last_descriptors), S,
                                        C.c_float(th), int(self.check_ori), ptr(kps), ptr(desc), ptr(counts), cap,
                                        ptr(match), ptr(nm), C.byref(t)))
        return t.value

    def submit_points(self, images, Tcw, points, point_start, th: float, out):
        """submit() with the last-frame inputs as packed 64-byte records (cmos_track_submit_points; pack_last_points builds them):
        the upload carries only the keypoints that have a usable map point.  Same results as submit()."""
        B, H, W = images.shape
        kps, desc, counts, match, nm = out
        cap = match.shape[1]
        t = C.c_int64(-1)
        check(self._L.cmos_track_submit_points(self._h, ptr(images), C.c_int64(H * W), W, W, H, B, ptr(Tcw), ptr(points),
                                               ptr(point_start), C.c_float(th), int(self.check_ori), ptr(kps), ptr(desc),
                                               ptr(counts), cap, ptr(match), ptr(nm), C.byref(t)))
        return t.value

    def map_reserve(self, n_slots: int):
        """Size the device-resident map-point table (cmos_track_map_reserve); growing keeps the slots already written."""
        check(self._L.cmos_track_map_reserve(self._h, int(n_slots)))

    def map_update(self, xw, descriptors, slots=None, first_slot: int = 0):
        """Write map points into the table (cmos_track_map_update): xw [n,3] float64, descriptors [n,32] uint8 into `slots` [n]
        int32, or into first_slot .. first_slot + n - 1 when slots is None.  Waits for the batches in flight."""
        xw = np.ascontiguousarray(xw, np.float64); descriptors = np.ascontiguousarray(descriptors, np.uint8)
        n = len(xw)
        assert xw.shape == (n, 3) and descriptors.shape == (n, 32)
        sp = None
        if slots is not None:
            slots = np.ascontiguousarray(slots, np.int32)
            assert slots.shape == (n,)
            sp = ptr(slots)
        check(self._L.cmos_track_map_update(self._h, n, sp, int(first_slot), ptr(xw), ptr(descriptors)))

    def submit_map(self, images, Tcw, assoc, assoc_start, th: float, out):
        """submit() with the last-frame inputs as 12-byte association records into the device-resident map-point table
        (cmos_track_submit_map; pack_map_associations builds them).  Same results as submit()."""
        B, H, W = images.shape
        kps, desc, counts, match, nm = out
        cap = match.shape[1]
        t = C.c_int64(-1)
        check(self._L.cmos_track_submit_map(self._h, ptr(images), C.c_int64(H * W), W, W, H, B, ptr(Tcw), ptr(assoc),
                                            ptr(assoc_start), C.c_float(th), int(self.check_ori), ptr(kps), ptr(desc),
                                            ptr(counts), cap, ptr(match), ptr(nm), C.byref(t)))
        return t.value

    def wait(self, ticket: int):
        check(self._L.cmos_track_wait(self._h, C.c_int64(ticket)))

    def launch_count(self) -> int:
        n = C.c_int32()
        check(self._L.cmos_track_last_launch_count(self._h, C.byref(n)))
        return n.value

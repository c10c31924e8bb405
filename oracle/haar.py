"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/haar_oracle.c."""
from __future__ import annotations

import ctypes
from ctypes import c_double, c_int, c_void_p

import numpy as np

from . import cbs as _cbs


def _lib():
    L = _cbs.lib()
    if not getattr(L, "_haar_ready", False):
        L.haaro_conv.restype = None
        L.haaro_conv.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p]
        L.haaro_find_peaks.restype = c_int
        L.haaro_find_peaks.argtypes = [c_void_p, c_int, c_void_p]
        L.haaro_fdr_thres.restype = c_double
        L.haaro_fdr_thres.argtypes = [c_void_p, c_int, c_double, c_double]
        L.haaro_segment.restype = c_int
        L.haaro_segment.argtypes = [c_void_p, c_void_p, c_int, c_double, c_void_p, c_void_p]
        L._haar_ready = True
    return L


def conv(x, w, h):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    wp = None
    if w is not None:
        w = np.ascontiguousarray(w, dtype=np.float64)
        wp = w.ctypes.data
    _lib().haaro_conv(x.ctypes.data, wp, len(x), h, out.ctypes.data)
    return out


def find_peaks(s):
    s = np.ascontiguousarray(s, dtype=np.float64)
    out = np.empty(len(s) + 1, dtype=np.int32)
    k = _lib().haaro_find_peaks(s.ctypes.data, len(s), out.ctypes.data)
    return out[:k].astype(np.int64)


def fdr_thres(x, q, sigma):
    x = np.ascontiguousarray(x, dtype=np.float64)
    return _lib().haaro_fdr_thres(x.ctypes.data, len(x), q, sigma)


def segment(x, w, q):
    """haarSeg: (starts, means) of one arm."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    wp = None
    if w is not None:
        w = np.ascontiguousarray(w, dtype=np.float64)
        wp = w.ctypes.data
    starts = np.empty(len(x) + 1, dtype=np.int32)
    means = np.empty(len(x) + 1, dtype=np.float64)
    k = _lib().haaro_segment(x.ctypes.data, wp, len(x), q, starts.ctypes.data, means.ctypes.data)
    return starts[:k].astype(np.int64), means[:k].copy()


def segment_arms(x, w, arm_offsets, q):
    """All arms: (seg_arm, seg_lo, seg_hi, seg_mean) with global half-open bin ranges."""
    arm, lo, hi, mean = [], [], [], []
    off = np.asarray(arm_offsets, dtype=np.int64)
    for a in range(len(off) - 1):
        s, e = int(off[a]), int(off[a + 1])
        if e <= s:
            continue
        st, mn = segment(x[s:e], None if w is None else w[s:e], q)
        ends = np.append(st[1:], e - s)
        for k in range(len(st)):
            arm.append(a); lo.append(s + st[k]); hi.append(s + ends[k]); mean.append(mn[k])
    return (np.asarray(arm, dtype=np.int32), np.asarray(lo, dtype=np.int64), np.asarray(hi, dtype=np.int64),
            np.asarray(mean, dtype=np.float64))

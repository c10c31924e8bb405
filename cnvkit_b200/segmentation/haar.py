"""HaarSeg on the GPU -- host mirror of cnvlib/segmentation/haar.py (depth path)."""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib


def segment_arms(log2, weight, arm_offsets, fdr_q, device_ptrs=None, stream=0):
    """haarSeg (haar.py:257-371) for every arm of a batch in one call.

    Returns (seg_arm, seg_lo, seg_hi, seg_mean): half-open global bin ranges per segment, in arm
    order.  ``weight`` may be None (unweighted convolution and plain means).
    """
    off = _lib.i64(arm_offsets)
    n_arms = len(off) - 1
    nbins = int(off[-1]) if n_arms > 0 else 0
    cap = max(nbins, 1)
    seg_arm = np.empty(cap, dtype=np.int32)
    seg_lo = np.empty(cap, dtype=np.int64)
    seg_hi = np.empty(cap, dtype=np.int64)
    seg_mean = np.empty(cap, dtype=np.float64)
    nseg = ctypes.c_int64(0)
    if device_ptrs is None:
        x = _lib.f64(log2)
        w = None if weight is None else _lib.f64(weight)
        if len(x) != nbins or (w is not None and len(w) != nbins):
            raise ValueError("log2/weight length does not match arm_offsets")
        _lib.call("cnvk_haar_segment", _lib.ptr(x), _lib.ptr(w), _lib.ptr(off), n_arms, float(fdr_q), cap,
                  _lib.ptr(seg_arm), _lib.ptr(seg_lo), _lib.ptr(seg_hi), _lib.ptr(seg_mean), ctypes.addressof(nseg))
    else:
        _lib.call("cnvk_haar_segment_dev", int(device_ptrs[0]), int(device_ptrs[1] or 0), _lib.ptr(off), n_arms,
                  float(fdr_q), cap, _lib.ptr(seg_arm), _lib.ptr(seg_lo), _lib.ptr(seg_hi), _lib.ptr(seg_mean),
                  ctypes.addressof(nseg), int(stream))
    k = nseg.value
    return seg_arm[:k].copy(), seg_lo[:k].copy(), seg_hi[:k].copy(), seg_mean[:k].copy()


def haarSeg(signal, breaksFdrQ, weights=None):
    """Single-array seam with the reference's return dict (haar.py:364-371)."""
    signal = np.asarray(signal, dtype=float)
    n = len(signal)
    arm, lo, hi, mean = segment_arms(signal, weights, np.array([0, n]), breaksFdrQ)
    return {"start": lo, "end": hi - 1, "size": hi - lo, "mean": mean}

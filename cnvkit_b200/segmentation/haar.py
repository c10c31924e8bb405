"""HaarSeg on the GPU -- host mirror of cnvlib/segmentation/haar.py (depth path)."""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib


SIGMA_FLOOR = 1e-10      # haar.py:43
WEIGHT_FLOOR = 1e-6      # haar.py:48


def segment_arms(log2, weight, arm_offsets, fdr_q, device_ptrs=None, stream=0, sigma_override=None):
    """haarSeg (haar.py:257-371) for every arm of a batch in one call.

    Returns (seg_arm, seg_lo, seg_hi, seg_mean): half-open global bin ranges per segment, in arm
    order.  ``weight`` may be None (unweighted convolution and plain means).  ``sigma_override``: one noise
    estimate per arm (NaN = the signal's own), haar.py:334-339 -- host-array path only.
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
    if sigma_override is not None:
        if device_ptrs is not None:
            raise ValueError("sigma_override is supported on the host-array path")
        x = _lib.f64(log2)
        w = None if weight is None else _lib.f64(weight)
        sg = _lib.f64(sigma_override)
        if len(x) != nbins or (w is not None and len(w) != nbins) or len(sg) != n_arms:
            raise ValueError("signal / weight / sigma_override lengths do not match arm_offsets")
        _lib.call("cnvk_haar_segment_sigma", _lib.ptr(x), _lib.ptr(w), _lib.ptr(off), n_arms, float(fdr_q), _lib.ptr(sg),
                  cap, _lib.ptr(seg_arm), _lib.ptr(seg_lo), _lib.ptr(seg_hi), _lib.ptr(seg_mean), ctypes.addressof(nseg))
    elif device_ptrs is None:
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


def level1_sigma(values, arm_offsets):
    """haarSeg's peakSigmaEst per arm: 1.4826 * median |HaarConv(values, None, 1)|, floored at SIGMA_FLOOR
    (haar.py:307-333; with compacted observed BAF values this is _observed_baf_sigma, :154-170)."""
    off = _lib.i64(arm_offsets)
    x = _lib.f64(values)
    out = np.full(len(off) - 1, SIGMA_FLOOR, dtype=np.float64)
    if len(off) > 1 and len(x):
        _lib.call("cnvk_haar_sigma", _lib.ptr(x), _lib.ptr(off), len(off) - 1, _lib.ptr(out))
    return out


def segment_means(signal, weight, seg_lo, seg_hi):
    """SegmentByPeaks' per-segment (weighted) means (haar.py:405-451) for given half-open bin ranges."""
    x = _lib.f64(signal)
    w = None if weight is None else _lib.f64(weight)
    lo, hi = _lib.i64(seg_lo), _lib.i64(seg_hi)
    out = np.empty(len(lo), dtype=np.float64)
    if len(lo):
        _lib.call("cnvk_haar_segment_means", _lib.ptr(x), _lib.ptr(w), len(x), _lib.ptr(lo), _lib.ptr(hi), len(lo),
                  _lib.ptr(out))
    return out


def unify_levels(base, addon, window):
    """UnifyLevels (haar.py:615-663): merge two sorted breakpoint lists, dropping addon values within `window`
    (inclusive) of a base value."""
    base = np.asarray(base, dtype=np.int64)
    addon = np.asarray(addon, dtype=np.int64)
    if not len(addon):
        return base
    if not len(base):
        return addon.copy()
    pos = np.searchsorted(base, addon)
    nb = len(base)
    left = np.clip(pos - 1, 0, nb - 1)
    right = np.clip(pos, 0, nb - 1)
    too_close = ((pos > 0) & ((addon - base[left]) <= window)) | ((pos < nb) & ((base[right] - addon) <= window))
    joined = np.concatenate([base, addon[~too_close]])
    joined.sort()
    return joined

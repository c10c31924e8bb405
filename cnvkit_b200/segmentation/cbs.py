"""Circular binary segmentation on the GPU -- replaces the R/DNAcopy subprocess
(cnvlib/segmentation/cbs.py:1-78, call site cnvlib/segmentation/__init__.py:292-322)."""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib

# DNAcopy::segment defaults (cbs.py:60 only passes weights and alpha) and cbs.py:49's seed
DEFAULTS = dict(nperm=10000, min_width=2, kmax=25, nmin=200, eta=0.05, hybrid=True, seed=0xA5EED)


def make_params(alpha, prune=True, tailp_bracket=True, **overrides) -> _lib.CbsParams:
    kw = dict(DEFAULTS)
    kw.update(overrides)
    prune = int(bool(prune)) | (0 if tailp_bracket else 2)
    return _lib.CbsParams(
        float(alpha), int(kw["nperm"]), int(kw["min_width"]), int(kw["kmax"]), int(kw["nmin"]), float(kw["eta"]),
        int(bool(kw["hybrid"])), int(prune), int(kw["seed"]),
    )


def getbdry(eta: float, nperm: int, max_ones: int) -> np.ndarray:
    out = np.zeros(max_ones * (max_ones + 1) // 2, dtype=np.int32)
    _lib.call("cnvk_cbs_getbdry", float(eta), int(nperm), int(max_ones), _lib.ptr(out))
    return out


def segment_arms(log2, weight, arm_offsets, alpha=1e-4, prune=True, device_ptrs=None, stream=0, tailp_bracket=True,
                 **overrides):
    """Segment every arm of a batch in one call.

    Parameters are numpy arrays (host path) unless ``device_ptrs=(d_log2, d_weight)`` gives raw
    CUDA pointers of resident float64 columns.  Returns (seg_arm, seg_lo, seg_hi, seg_mean, stats)
    with half-open global bin ranges.
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
    params = make_params(alpha, prune=prune, tailp_bracket=tailp_bracket, **overrides)
    stats = _lib.CbsStats()
    if device_ptrs is None:
        x = _lib.f64(log2)
        w = _lib.f64(weight)
        if len(x) != nbins or len(w) != nbins:
            raise ValueError("log2/weight length does not match arm_offsets")
        _lib.call("cnvk_cbs_segment", _lib.ptr(x), _lib.ptr(w), _lib.ptr(off), n_arms, ctypes.addressof(params), cap,
                  _lib.ptr(seg_arm), _lib.ptr(seg_lo), _lib.ptr(seg_hi), _lib.ptr(seg_mean), ctypes.addressof(nseg),
                  ctypes.addressof(stats))
    else:
        _lib.call("cnvk_cbs_segment_dev", int(device_ptrs[0]), int(device_ptrs[1]), _lib.ptr(off), n_arms,
                  ctypes.addressof(params), cap, _lib.ptr(seg_arm), _lib.ptr(seg_lo), _lib.ptr(seg_hi),
                  _lib.ptr(seg_mean), ctypes.addressof(nseg), ctypes.addressof(stats), int(stream))
    k = nseg.value
    return seg_arm[:k].copy(), seg_lo[:k].copy(), seg_hi[:k].copy(), seg_mean[:k].copy(), stats.as_dict()


# DNAcopy::smooth.CNA defaults (cbs.py:57 calls it without arguments)
SMOOTH_DEFAULTS = dict(smooth_region=10, outlier_sd_scale=4.0, smooth_sd_scale=2.0, trim=0.025)


def smooth_cna(log2, arm_offsets, device_ptrs=None, stream=0, **overrides):
    """DNAcopy smooth.CNA of every arm (``--smooth-cbs``, cbs.py:51-58).  Host arrays in / out, or with
    ``device_ptrs=(d_log2, d_out)`` on resident columns (returns None).  Arms of fewer than 2 bins are copied
    unchanged (cbs.py:56: smoothing is skipped for them)."""
    kw = dict(SMOOTH_DEFAULTS)
    kw.update(overrides)
    off = _lib.i64(arm_offsets)
    n_arms = len(off) - 1
    if device_ptrs is not None:
        _lib.call("cnvk_smooth_cna_dev", int(device_ptrs[0]), _lib.ptr(off), n_arms, int(kw["smooth_region"]),
                  float(kw["outlier_sd_scale"]), float(kw["smooth_sd_scale"]), float(kw["trim"]), int(device_ptrs[1]),
                  0, int(stream))
        return None
    x = _lib.f64(log2)
    out = np.empty_like(x)
    _lib.call("cnvk_smooth_cna", _lib.ptr(x), _lib.ptr(off), n_arms, int(kw["smooth_region"]),
              float(kw["outlier_sd_scale"]), float(kw["smooth_sd_scale"]), float(kw["trim"]), _lib.ptr(out), 0)
    return out


def node_diag(log2, weight, node_offsets, alpha=1e-4, **overrides):
    """Decision-chain diagnostics of every interval tested once as a node, nothing decided early
    (cnvk_cbs_node_diag): list of dicts with ``ostat1`` (sqrt T^2), ``ostat`` (0.99999 T^2), ``arc``, ``tailp`` /
    ``delta`` (hybrid nodes), ``nrej`` (permutations, out of nperm, whose statistic reached ``ostat``),
    ``flank_p`` (ternary-split p-values with the permutations forced), ``hybrid`` and ``flat``."""
    off = _lib.i64(node_offsets)
    n_nodes = len(off) - 1
    x = _lib.f64(log2)
    w = _lib.f64(weight)
    params = make_params(alpha, **overrides)
    out = np.zeros((n_nodes, 6), dtype=np.float64)
    iout = np.zeros((n_nodes, 5), dtype=np.int32)
    _lib.call("cnvk_cbs_node_diag", _lib.ptr(x), _lib.ptr(w), _lib.ptr(off), n_nodes, ctypes.addressof(params),
              _lib.ptr(out), _lib.ptr(iout))
    return [dict(ostat1=o[0], tailp=o[1], delta=o[2], ostat=o[3], flank_p=(o[4], o[5]), arc=(int(i[0]), int(i[1])),
                 nrej=int(i[2]), hybrid=bool(i[3]), flat=bool(i[4]), nperm_run=int(params.nperm))
            for o, i in zip(out, iout)]

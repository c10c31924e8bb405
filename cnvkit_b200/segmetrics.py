"""Segment-level bootstrap confidence intervals on the GPU -- the 'ci' statistic of
``cnvlib.segmetrics.do_segmetrics`` (cnvlib/segmetrics.py:24-121), which ``cnvkit.py batch`` computes for
every sample right after segmentation (cnvlib/batch.py:559-566: ``interval_stats=["ci"], alpha=0.5,
smoothed=True``) and which the "ci" call filter reads.

Same signature, column names, warning and error text as the reference.  The other statistics
(location / spread / 'pi') are outside this path: ``do_segmetrics`` raises ``NotImplementedError`` for
them, and ``cnvkit_b200.install`` lets the reference compute those columns and adds ours.
"""
from __future__ import annotations

import ctypes
import logging

import numpy as np

from . import _lib, params

INT32_MAX = 2 ** 31 - 1
NOISE_SEED = 0xC0FFEE


def _smooth_upto(smoothed) -> int:
    """`smoothed` as the reference reads it (segmetrics.py:223): bool -> always / never; int -> k <= n."""
    if isinstance(smoothed, (bool, np.bool_)):
        return INT32_MAX if smoothed else -1
    return int(min(max(int(smoothed), -1), INT32_MAX))


def make_params(alpha, bootstraps=100, smoothed=False, noise_seed=NOISE_SEED) -> _lib.BootParams:
    if not 0 < alpha < 1:
        raise ValueError(f"alpha must be between 0 and 1; got {alpha}")   # segmetrics.py:204-205
    p = _lib.BootParams()
    p.alpha = float(alpha)
    p.bootstraps = int(bootstraps)
    p.smooth_upto = _smooth_upto(smoothed)
    p.noise_seed = int(noise_seed)          # pcg_state / pcg_inc stay zero: numpy's PCG64(0xA5EED)
    return p


def effective_bootstraps(alpha, bootstraps, warn=True) -> int:
    """segmetrics.py:206-214 (the reference warns once per segment; once per table here)."""
    if bootstraps <= 2 / alpha:
        new_boots = int(np.ceil(2 / alpha))
        if warn:
            logging.warning("%d bootstraps not enough to estimate CI alpha level %f; increasing to %d",
                            bootstraps, alpha, new_boots)
        return new_boots
    return int(bootstraps)


def bootstrap_ci(values, weights, seg_lo, seg_hi, alpha=0.05, bootstraps=100, smoothed=False,
                 noise_seed=NOISE_SEED, device_ptrs=None, stream=0):
    """confidence_interval_bootstrap (segmetrics.py:176-245) for every row range [seg_lo, seg_hi) of the
    bin columns.  Host arrays in, (ci_lo, ci_hi) numpy arrays out; with ``device_ptrs=(d_values,
    d_weights)`` the columns are resident float64 CUDA buffers and the result is a pair of CUDA tensors."""
    lo = _lib.i64(seg_lo)
    hi = _lib.i64(seg_hi)
    if len(lo) != len(hi):
        raise ValueError("seg_lo / seg_hi differ in length")
    p = make_params(alpha, bootstraps, smoothed, noise_seed)
    ns = len(lo)
    if device_ptrs is None:
        v = _lib.f64(values)
        w = _lib.f64(weights)
        if len(v) != len(w):
            raise ValueError("values / weights differ in length")
        out_lo = np.full(ns, np.nan)
        out_hi = np.full(ns, np.nan)
        if ns:
            _lib.call("cnvk_bootstrap_ci", _lib.ptr(v), _lib.ptr(w), len(v), _lib.ptr(lo), _lib.ptr(hi), ns,
                      ctypes.addressof(p), _lib.ptr(out_lo), _lib.ptr(out_hi))
        return out_lo, out_hi
    import torch

    from . import _dev

    d_lo = _dev.empty(max(ns, 1), torch.float64)
    d_hi = _dev.empty(max(ns, 1), torch.float64)
    if ns:
        _lib.call("cnvk_bootstrap_ci_dev", int(device_ptrs[0]), int(device_ptrs[1]), _lib.ptr(lo), _lib.ptr(hi), ns,
                  ctypes.addressof(p), d_lo.data_ptr(), d_hi.data_ptr(), stream)
    return d_lo[:ns], d_hi[:ns]


class BinIndex:
    """Per-chromosome row runs and coordinate columns of a bin table, checked once (ascending start and
    end within each chromosome) so that many segment tables can be matched against it cheaply."""

    def __init__(self, bins):
        from .genome import chrom_runs

        self.n = len(bins)
        self.start = bins["start"].to_numpy()
        self.end = bins["end"].to_numpy()
        self.runs = {}
        if not self.n:
            return
        try:
            names, run_off, _ = chrom_runs(bins["chromosome"].to_numpy())
        except ValueError as exc:
            raise NotImplementedError("bins of one chromosome must be contiguous rows (sort the .cnr first)") from exc
        for r, name in enumerate(names):
            a, b = int(run_off[r]), int(run_off[r + 1])
            st, en = self.start[a:b], self.end[a:b]
            if len(st) > 1 and np.any(st[1:] < st[:-1]):
                row = int(np.flatnonzero(st[1:] < st[:-1])[0]) + 1
                raise ValueError(
                    "Genomic intervals must be sorted by start position to be searched: "
                    f"start={st[row]} follows start={st[row - 1]}, at row {row} of {len(st)}. "
                    "To fix, sort the rows by start -- GenomicArray.sort, or sort_values.")   # intersect.py:270-310
            if len(en) > 1 and np.any(en[1:] < en[:-1]):
                raise NotImplementedError("nested bins (end not ascending) are outside the GPU path")
            self.runs[name] = (a, b)

    def ranges(self, segs):
        """(lo, hi) row positions of the bins each segment selects in 'outer' mode."""
        s_chrom = segs["chromosome"].to_numpy()
        s_start = segs["start"].to_numpy()
        s_end = segs["end"].to_numpy()
        lo = np.zeros(len(segs), dtype=np.int64)
        hi = np.zeros(len(segs), dtype=np.int64)
        if not self.n or not len(segs):
            return lo, hi
        seg_names, seg_codes = np.unique(s_chrom.astype(object), return_inverse=True)
        for code, name in enumerate(seg_names):
            run = self.runs.get(name)
            if run is None:
                continue
            a, b = run
            sel = np.flatnonzero(seg_codes == code)
            lo[sel] = a + np.searchsorted(self.end[a:b], s_start[sel], "right")
            hi[sel] = a + np.searchsorted(self.start[a:b], s_end[sel], "left")
        return lo, np.maximum(hi, lo)


def segment_bin_ranges(bins, segs):
    """Rows of the bin table each segment selects: ``cnarr.iter_ranges_of(segarr, "log2", "outer", True)``
    (skgenome/gary.py:582-630 -> intersect.py:169-342).  Per chromosome, in 'outer' mode, the selection
    is the run [searchsorted(end, seg_start, "right"), searchsorted(start, seg_end)).  Returns (lo, hi)
    as positions into ``bins``; hi == lo where nothing is selected (unknown chromosome included)."""
    if not len(bins) or not len(segs):
        return np.zeros(len(segs), dtype=np.int64), np.zeros(len(segs), dtype=np.int64)
    return BinIndex(bins).ranges(segs)


def do_segmetrics(cnarr, segarr, location_stats=(), spread_stats=(), interval_stats=(), alpha=0.05,
                  bootstraps=100, smoothed=10, skip_low=False):
    """Compute the bootstrap confidence interval of every segment's mean (columns ``ci_lo``, ``ci_hi``).

    Mirrors cnvlib/segmetrics.py:24-121 for ``interval_stats=["ci"]``.
    """
    extra = [s for s in interval_stats if s != "ci"]
    if location_stats or spread_stats or extra:
        raise NotImplementedError(
            "only interval_stats=['ci'] runs on the GPU path; cnvkit_b200.install routes the other "
            f"statistics ({list(location_stats) + list(spread_stats) + extra}) to cnvlib")
    segarr = segarr.copy()
    if "ci" not in interval_stats:
        return segarr
    if not 0 < alpha < 1:
        raise ValueError(f"alpha must be between 0 and 1; got {alpha}")
    bins = cnarr.data
    keep = np.ones(len(bins), dtype=bool)
    if skip_low:    # CopyNumArray.drop_low_coverage (cnary.py:315-335)
        l2 = bins["log2"].to_numpy(dtype=np.float64)
        keep &= ~((l2 < params.NULL_LOG2_COVERAGE - params.MIN_REF_COVERAGE) | np.isnan(l2))
        if "depth" in bins.columns:
            dp = bins["depth"].to_numpy(dtype=np.float64)
            keep &= ~((dp == 0) | np.isnan(dp))
        bins = bins[keep]
    lo, hi = segment_bin_ranges(bins, segarr.data)
    values = bins["log2"].to_numpy(dtype=np.float64)
    weights = bins["weight"].to_numpy(dtype=np.float64)
    valid = ~np.isnan(weights)        # calc_intervals drops NaN-weight bins (segmetrics.py:168-172)
    if not valid.all():
        rank = np.concatenate([[0], np.cumsum(valid)]).astype(np.int64)
        lo, hi = rank[lo], rank[hi]
        values, weights = values[valid], weights[valid]
    if len(segarr):
        effective_bootstraps(alpha, bootstraps, warn=bool(((hi - lo) >= 1).any()))
    ci_lo, ci_hi = bootstrap_ci(values, weights, lo, hi, alpha, bootstraps, smoothed)
    segarr["ci_lo"] = ci_lo
    segarr["ci_hi"] = ci_hi
    return segarr

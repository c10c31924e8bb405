"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

CPU restatement of the *composed* exome hot path on plain columns, built from the pinned units
of oracle/np_oracle.py: do_fix (cnvlib/fix.py:20-215, 218-292) and the numeric part of
do_segmentation (cnvlib/segmentation/__init__.py:222-369) for 'cbs' and 'haar'.  Used by the
C2-scale parity tests and as bench.py's `cpu_baseline` for the exome leg (kind "port").

Rows are expected in genomic order with each chromosome contiguous; autosomes are the
chromosomes not named chrX/chrY/X/Y (skgenome/gary.py:268-307 without PAR handling).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import np_oracle as O

INSERT_SIZE = 250           # cnvlib/params.py:14
NULL_LOG2 = -20.0           # cnvlib/params.py NULL_LOG2_COVERAGE
MIN_REF = -5.0              # cnvlib/params.py MIN_REF_COVERAGE


def rolling_median_pandas(x, width):
    """cnvlib/smoothing.py:77-87 with the reference's own library call (pandas skiplist median):
    the fast CPU form, numerically identical to np_oracle.rolling_median."""
    x = np.asarray(x, dtype=float)
    if len(x) < 2:
        return x
    wing = O.width2wing(width, len(x))
    padded = O.pad_mirror(x, wing)
    r = pd.Series(padded).rolling(2 * wing + 1, 1, center=True).median()
    return np.asarray(r[wing:-wing])


def center_by_window(log2, key, fraction):
    """np_oracle.center_by_window with the pandas median (cnvlib/fix.py:373-427)."""
    n = len(log2)
    shuffle = np.random.default_rng(0xA5EED).permutation(n)
    order = np.argsort(np.asarray(key, dtype=float)[shuffle], kind="mergesort")
    src = shuffle[order]
    xs = log2[src]
    out = log2.copy()
    out[src] = xs - rolling_median_pandas(xs, fraction)
    return out


def autosome_mask(chromosome):
    c = np.asarray(chromosome, dtype=object)
    return ~np.isin(c, ["chrX", "chrY", "X", "Y"])


def _adjust_group(chrom, start, end, log2, depth, ref, do_gc, do_edge, do_rmask, skip_low):
    """load_adjust_coverages (fix.py:218-292) for one bin set; returns (keep mask, corrected log2[keep])."""
    bad = O.mask_bad_bins(ref["log2"], ref["spread"], ref.get("depth"), ref.get("gc")) | np.isnan(log2)
    keep = ~bad
    chrom, start, end, log2, depth = chrom[keep], start[keep], end[keep], log2[keep].copy(), depth[keep]
    n = len(log2)
    if n == 0:
        return keep, log2
    log2 += O.center_all_shift(log2, chrom, autosome_mask(chrom), depth, skip_low=skip_low)
    if (log2 > NULL_LOG2 - MIN_REF).sum() <= n // 2:
        return keep, log2
    frac = max(0.01, n ** -0.5)
    if do_gc and "gc" in ref:
        log2 = center_by_window(log2, ref["gc"][keep], frac)
    if do_edge:
        log2 = center_by_window(log2, O.edge_bias(chrom, start, end, INSERT_SIZE), frac)
    if do_rmask and "rmask" in ref:
        log2 = center_by_window(log2, ref["rmask"][keep], frac)
    return keep, log2


def do_fix(bins, sample, ref):
    """fix.do_fix on columns.  bins: chromosome,start,end,is_target (+gc, rmask inside `ref`);
    sample: log2, depth; ref: log2, depth, spread, gc, rmask -- all over the same rows.
    Returns dict(keep, log2, weight) where log2/weight cover the kept rows in genomic order."""
    chrom = np.asarray(bins["chromosome"], dtype=object)
    start, end = np.asarray(bins["start"]), np.asarray(bins["end"])
    is_t = np.asarray(bins["is_target"], dtype=bool)
    n = len(start)
    keep = np.zeros(n, dtype=bool)
    log2 = np.full(n, np.nan)
    for grp, (gc, edge, rmask, skip_low) in ((is_t, (True, True, False, True)), (~is_t, (True, False, True, False))):
        idx = np.flatnonzero(grp)
        if not len(idx):
            continue
        sub_ref = {k: np.asarray(v)[idx] for k, v in ref.items()}
        k, l2 = _adjust_group(chrom[idx], start[idx], end[idx], np.asarray(sample["log2"], dtype=float)[idx],
                              np.asarray(sample["depth"], dtype=float)[idx], sub_ref, gc, edge, rmask, skip_low)
        keep[idx[k]] = True
        log2[idx[k]] = l2
    kk = np.flatnonzero(keep)
    c, s, e = chrom[kk], start[kk], end[kk]
    depth = np.asarray(sample["depth"], dtype=float)[kk]
    ratio = log2[kk] - np.asarray(ref["log2"])[kk]
    weight = O.apply_weights(ratio, c, s, e, ~is_t[kk], np.asarray(ref["log2"])[kk], np.asarray(ref["spread"])[kk], depth)
    ratio = ratio + O.center_all_shift(ratio, c, autosome_mask(c), depth, skip_low=True)
    return dict(keep=keep, log2=ratio, weight=weight)


def arm_offsets(chrom, start, end, min_gap_size=1e5, min_arm_bins=50):
    """Row offsets of GenomicArray.by_arm (skgenome/gary.py:309-355)."""
    off = O.group_runs(chrom)
    cuts = [0]
    for lo, hi in zip(off[:-1], off[1:]):
        n = hi - lo
        margin = max(min_arm_bins, round(0.1 * n))
        if n > 2 * margin + 1:
            gaps = start[lo + margin + 1: hi - margin] - end[lo + margin: hi - margin - 1]
            k = int(np.argmax(gaps))
            if gaps[k] >= min_gap_size:
                cuts.append(int(lo + k + margin + 1))
        cuts.append(int(hi))
    return np.asarray(cuts, dtype=np.int64)


def segment_columns(chrom, start, end, log2, weight, method, threshold=1e-4, skip_outliers=10, nthreads=0,
                    prune=False):
    """Numeric core of do_segmentation: per-arm filters, %.6g quantisation (cbs), the segmenter.
    Returns (seg_arm, seg_lo, seg_hi, seg_mean) with bin ranges into the *filtered* rows, plus the
    filtered row indices."""
    off = arm_offsets(chrom, start, end)
    keep = np.isfinite(log2)
    if skip_outliers:
        for lo, hi in zip(off[:-1], off[1:]):
            idx = np.flatnonzero(keep[lo:hi]) + lo
            if len(idx):
                keep[idx[O.outlier_mask(log2[idx], 50, 0.95, skip_outliers)]] = False
    keep &= ~(weight == 0)
    fidx = np.flatnonzero(keep)
    arm_id = np.repeat(np.arange(len(off) - 1), np.diff(off))[fidx]
    foff = np.concatenate([[0], np.cumsum(np.bincount(arm_id, minlength=len(off) - 1))]).astype(np.int64)
    x, w = log2[fidx], weight[fidx]
    if method == "cbs":
        from . import cbs as OC

        xq = np.char.mod("%.6g", x).astype(float)
        wq = np.char.mod("%.6g", w).astype(float)
        res = OC.segment(xq, wq, foff, OC.make_params(alpha=threshold, prune=prune), nthreads=nthreads)[:4]
    else:
        from . import haar as OH

        # segment_haar re-splits the filtered arm tables with by_arm() again (haar.py:80-82)
        units = [0]
        for lo, hi in zip(foff[:-1], foff[1:]):
            if hi > lo:
                units.extend((arm_offsets(chrom[fidx[lo:hi]], start[fidx[lo:hi]], end[fidx[lo:hi]])[1:] + lo).tolist())
        res = OH.segment_arms(x, w, np.asarray(units, dtype=np.int64), threshold)
    return res, fidx

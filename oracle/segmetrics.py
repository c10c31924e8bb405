"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

CPU restatement (numpy) of the bootstrap confidence interval of the segment mean:
cnvlib/segmetrics.py:176-245 confidence_interval_bootstrap, :248-289 _smooth_samples_by_weight,
:292-351 _bca_correct_alpha, and the bin selection of do_segmetrics (:93, GenomicArray.iter_ranges_of
-> skgenome/intersect.py:169-342, mode "outer").  Pinned by tests/golden/segmetrics.npz, which
oracle/make_golden.py writes from the unmodified reference.

What is restated and what is not:
* the resampled indices are the reference's own call, default_rng(0xA5EED).integers(0, k, (B, k));
  `pcg64_lemire_indices` is an independent restatement of that stream (PCG64 XSL-RR, 32-bit halves,
  Lemire rejection) used to pin the CUDA generator's arithmetic;
* replicate means are np.average, percentiles np.percentile -- the reference's library calls;
* the jackknife of the BCa correction uses the closed form (S - v_i w_i) / (W - w_i) instead of the
  reference's O(k^2) concatenations: equal to rounding (tolerance 1e-9 in the tests);
* the reference draws the smoothing noise from an UNSEEDED generator, so smoothed intervals are not
  reproducible there; here (and in the CUDA path) the noise is a counter-based normal,
  `boot_normal(key, r * k + j)`.  Smoothed results are compared with the reference statistically only.
"""
from __future__ import annotations

import math

import numpy as np

M64 = (1 << 64) - 1
M128 = (1 << 128) - 1
PCG_MULT = 0x2360ED051FC65DA44385DF649FCCF645
# numpy.random.PCG64(0xA5EED).state["state"]
PCG_STATE_A5EED = 0x52637B81F59D8CC600B11ABC6A6A0481
PCG_INC_A5EED = 0x392A8421E3A06208808513643339BEB7
NOISE_SEED = 0xC0FFEE


def pcg64_raw32(count, state=PCG_STATE_A5EED, inc=PCG_INC_A5EED):
    """First `count` 32-bit draws of numpy's PCG64: each 64-bit output gives its low half, then its high."""
    out = np.empty(count + 1, dtype=np.uint64)
    s = state
    for t in range((count + 1) // 2):
        s = (s * PCG_MULT + inc) & M128
        hi, lo = s >> 64, s & M64
        x = hi ^ lo
        rot = s >> 122
        o = ((x >> rot) | (x << ((64 - rot) & 63))) & M64
        out[2 * t] = o & 0xFFFFFFFF
        out[2 * t + 1] = o >> 32
    return out[:count]


def pcg64_lemire_indices(k, count, state=PCG_STATE_A5EED, inc=PCG_INC_A5EED):
    """Generator(PCG64).integers(0, k, count) restated: keep a draw d iff low32(d*k) >= 2^32 mod k;
    the value is high32(d*k).  Returns (indices, raw draws consumed)."""
    thr = ((1 << 32) - k) % k
    need = count + 64
    while True:
        raw = pcg64_raw32(need, state, inc)
        m = raw * np.uint64(k)
        keep = (m & np.uint64(0xFFFFFFFF)) >= np.uint64(thr)
        if int(keep.sum()) >= count:
            pos = np.flatnonzero(keep)[:count]
            return (m[pos] >> np.uint64(32)).astype(np.int64), int(pos[-1]) + 1 if count else 0
        need *= 2


def _mix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def noise_key(seg_index, noise_seed=NOISE_SEED):
    with np.errstate(over="ignore"):
        inner = _mix64(np.uint64(seg_index) + np.uint64(0x632BE59BD9B4E019))
        return _mix64(np.uint64(noise_seed) ^ inner)


def boot_normal(key, c):
    """Standard normal number c of stream `key`: Box-Muller on two hashed uniforms (bootstrap.cu)."""
    c = np.asarray(c, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h1 = _mix64(np.uint64(key) + (c + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15))
        h2 = _mix64(h1 ^ np.uint64(0xD1B54A32D192ED03))
    u1 = ((h1 >> np.uint64(11)).astype(np.float64) + 0.5) * 2.0 ** -53
    u2 = (h2 >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def _ndtri(p):
    from scipy import special

    return special.ndtri(p)


def _ndtr(x):
    from scipy import special

    return special.ndtr(x)


def bca_correct_alpha(values, weights, bootstrap_dist, alphas):
    """segmetrics.py:292-351 with the closed-form jackknife."""
    n_boots = len(bootstrap_dist)
    svw, sw = (values * weights).sum(), weights.sum()
    orig_mean = np.average(values, weights=weights)
    below = int((bootstrap_dist < orig_mean).sum())
    prop = below / n_boots
    if prop == 0 or prop == 1:
        return alphas
    z0 = _ndtri(prop)
    zalpha = _ndtri(alphas)
    u = (svw - values * weights) / (sw - weights)
    uu = u.mean() - u
    uu_var = (uu ** 2).sum()
    if uu_var < 1e-10:
        return alphas
    acc = (uu ** 3).sum() / (6 * uu_var ** 1.5)
    denom = 1 - acc * (z0 + zalpha)
    if (denom <= 0).any():
        return alphas
    new = _ndtr(z0 + (z0 + zalpha) / denom)
    if not (0 < new[0] < 1 and 0 < new[1] < 1):
        return alphas
    return new


def effective_bootstraps(alpha, bootstraps):
    if not 0 < alpha < 1:
        raise ValueError(f"alpha must be between 0 and 1; got {alpha}")
    if bootstraps <= 2 / alpha:
        return int(np.ceil(2 / alpha))
    return bootstraps


def confidence_interval_bootstrap(values, weights, alpha, bootstraps=100, smoothed=False, seg_index=0,
                                  noise_seed=NOISE_SEED, return_dist=False):
    """segmetrics.py:176-245.  `smoothed`: bool, or int threshold (smooth when k <= smoothed)."""
    values = np.asarray(values, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    bootstraps = effective_bootstraps(alpha, bootstraps)
    k = len(values)
    if k < 2:
        return np.repeat(values[0], 2)
    use_smoothing = smoothed if isinstance(smoothed, bool) else k <= smoothed
    idx = np.random.default_rng(0xA5EED).integers(0, k, size=(bootstraps, k))
    key = noise_key(seg_index, noise_seed)
    bw = k ** (-1 / 4)
    dist = np.empty(bootstraps)
    for r in range(bootstraps):
        v, w = np.take(values, idx[r]), np.take(weights, idx[r])
        if use_smoothing:
            c = np.uint64(r) * np.uint64(k) + np.arange(k, dtype=np.uint64)
            v = v + (bw * np.sqrt(1 - w)) * boot_normal(key, c)
        dist[r] = np.average(v, weights=w)
    alphas = np.array([alpha / 2, 1 - alpha / 2])
    if not use_smoothing:
        alphas = bca_correct_alpha(values, weights, dist, alphas)
    ci = np.percentile(dist, list(100 * alphas))
    return (ci, dist) if return_dist else ci


def ci_table(values, weights, seg_lo, seg_hi, alpha=0.05, bootstraps=100, smoothed=10, noise_seed=NOISE_SEED):
    """calc_intervals (segmetrics.py:158-173) over row ranges [lo, hi) of the bin columns."""
    lo_out = np.full(len(seg_lo), np.nan)
    hi_out = np.full(len(seg_lo), np.nan)
    for s, (lo, hi) in enumerate(zip(seg_lo, seg_hi)):
        if hi > lo:
            lo_out[s], hi_out[s] = confidence_interval_bootstrap(
                values[lo:hi], weights[lo:hi], alpha, bootstraps, smoothed, seg_index=s, noise_seed=noise_seed)
    return lo_out, hi_out


def bin_ranges(bin_chrom, bin_start, bin_end, seg_chrom, seg_start, seg_end):
    """Rows of the bin table each segment selects in mode "outer" (intersect.py:169-342, _irange_simple):
    per chromosome, lo = searchsorted(bin_end, seg_start, "right"), hi = searchsorted(bin_start, seg_end).
    Bins of one chromosome must be contiguous rows with ascending start and end."""
    bin_chrom = np.asarray(bin_chrom, dtype=object)
    bin_start, bin_end = np.asarray(bin_start), np.asarray(bin_end)
    lo = np.zeros(len(seg_start), dtype=np.int64)
    hi = np.zeros(len(seg_start), dtype=np.int64)
    seg_chrom = np.asarray(seg_chrom, dtype=object)
    for chrom in dict.fromkeys(seg_chrom.tolist()):
        rows = np.flatnonzero(bin_chrom == chrom)
        sel = np.flatnonzero(seg_chrom == chrom)
        if not len(rows):
            continue
        base = rows[0]
        assert rows[-1] - base + 1 == len(rows), "chromosome rows not contiguous"
        lo[sel] = base + np.searchsorted(bin_end[rows], np.asarray(seg_start)[sel], "right")
        hi[sel] = base + np.searchsorted(bin_start[rows], np.asarray(seg_end)[sel], "left")
    hi = np.maximum(hi, lo)
    return lo, hi


def standard_error_interval(values, weights, alpha):
    """Normal-theory interval of the weighted mean -- a yardstick for the smoothed bootstrap's width."""
    w = weights / weights.sum()
    m = (values * w).sum()
    se = math.sqrt(((values - m) ** 2 * w ** 2).sum())
    z = _ndtri(1 - alpha / 2)
    return m - z * se, m + z * se

"""Seeded synthetic workloads of the shapes BASELINE.json names.

There is no network and no patient data: benchmarks and full-size property tests run on these
generators.  Values are rounded to 6 significant digits like every .cnr the reference writes
(skgenome/tabio/__init__.py:188-190) and like the TSV handed to DNAcopy
(cnvlib/segmentation/__init__.py:299-301).
"""
from __future__ import annotations

import numpy as np

# hg38 primary assembly lengths (Mb), chr1..22, X, Y
HG38_MB = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09, 133.28,
           114.36, 107.04, 101.99, 90.34, 83.26, 80.37, 58.62, 64.44, 46.71, 50.82, 156.04, 57.23]
CHROM_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]
STEPS = np.array([-1.0, -0.4, 0.3, 0.58, 1.0])


def round_sig6(x: np.ndarray) -> np.ndarray:
    """float('%.6g' % v) for every element (exact, via the C formatter)."""
    return np.char.mod("%.6g", x).astype(np.float64)


def chrom_bin_counts(total_bins: int) -> np.ndarray:
    frac = np.asarray(HG38_MB) / np.sum(HG38_MB)
    counts = np.floor(frac * total_bins).astype(np.int64)
    counts[0] += total_bins - counts.sum()
    return counts


def wgs_sample(total_bins: int = 3_000_000, n_events: int = 100, seed: int = 0xC3, sigma: float = 0.25,
               bin_size: int = 1000, quantise: bool = True):
    """C3: one WGS sample, contiguous 1 kb bins over 24 chromosomes, a 3 Mb centromere gap at
    40 % of each chromosome (=> 48 arms under GenomicArray.by_arm), log2 = N(0, sigma) plus
    `n_events` planted aberrant segments (2 change-points each, step from STEPS, geometric
    lengths >= 5 bins), weights ~ U(0.5, 1).

    Returns dict(chromosome_id, start, end, log2, weight, arm_offsets, truth) where truth holds
    the planted (lo, hi, step) in global bin coordinates.
    """
    rng = np.random.default_rng(seed)
    counts = chrom_bin_counts(total_bins)
    chrom_id = np.repeat(np.arange(len(counts)), counts)
    start = np.empty(total_bins, dtype=np.int64)
    arm_offsets = [0]
    pos = 0
    for c, n in enumerate(counts):
        p_bins = int(round(0.4 * n))
        s = np.arange(n, dtype=np.int64) * bin_size
        s[p_bins:] += 3_000_000
        start[pos : pos + n] = s
        arm_offsets += [pos + p_bins, pos + n]
        pos += n
    end = start + bin_size
    log2 = rng.normal(0.0, sigma, total_bins)
    truth = []
    arm_offsets = np.asarray(arm_offsets, dtype=np.int64)
    for _ in range(n_events):
        a = int(rng.integers(0, len(arm_offsets) - 1))
        lo_a, hi_a = int(arm_offsets[a]), int(arm_offsets[a + 1])
        length = 5 + int(rng.geometric(1.0 / 2000.0))
        length = min(length, (hi_a - lo_a) // 2)
        lo = int(rng.integers(lo_a + 5, hi_a - length - 5))
        step = float(rng.choice(STEPS))
        log2[lo : lo + length] += step
        truth.append((lo, lo + length, step))
    weight = rng.uniform(0.5, 1.0, total_bins)
    if quantise:
        log2 = round_sig6(log2)
        weight = round_sig6(weight)
    return dict(chromosome_id=chrom_id, start=start, end=end, log2=log2, weight=weight, arm_offsets=arm_offsets,
                truth=truth)

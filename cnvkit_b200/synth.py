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


# ---------------------------------------------------------------------------------------- exome (C2 / C4 / C5)
def exome_bins(n_targets: int = 150_000, n_antitargets: int = 50_000, seed: int = 0xC2):
    """C2 bin set: targets (size ~ N(267, 60) clipped to [80, 1000] bp, log-normal gaps with ~30 %
    closer than 250 bp so edge gains are exercised) interleaved with antitargets (20-150 kb,
    gene 'Antitarget') over 24 chromosomes with counts proportional to hg38, one 3 Mb gap at
    40 % of each chromosome (=> 2 arms per chromosome), gc ~ Beta(8, 8) clipped to [0.2, 0.8]
    (~5 % outside the [0.3, 0.7] mask window), rmask ~ Beta(2, 5).  Rows are in genomic order.

    Returns a dict of equal-length columns: chromosome (object), start, end (int64), gene (object),
    is_target (bool), gc, rmask (float64).
    """
    rng = np.random.default_rng(seed)
    n = n_targets + n_antitargets
    counts = chrom_bin_counts(n)
    is_target = np.zeros(n, dtype=bool)
    is_target[rng.permutation(n)[:n_targets]] = True
    size = np.where(is_target, np.clip(rng.normal(267, 60, n), 80, 1000), rng.uniform(20e3, 150e3, n)).astype(np.int64)
    close = rng.random(n) < 0.3
    gap = np.where(is_target, np.where(close, rng.uniform(0, 250, n), rng.lognormal(np.log(2500.0), 1.0, n)),
                   rng.uniform(0, 500, n)).astype(np.int64)
    start = np.empty(n, dtype=np.int64)
    chrom = np.empty(n, dtype=object)
    pos = 0
    for c, m in enumerate(counts):
        g = gap[pos:pos + m].copy()
        g[0] = 10_000
        g[int(round(0.4 * m))] += 3_000_000
        s = np.cumsum(g + np.concatenate([[0], size[pos:pos + m - 1]]))
        start[pos:pos + m] = s
        chrom[pos:pos + m] = CHROM_NAMES[c]
        pos += m
    end = start + size
    gene = np.where(is_target, np.char.add("G", (np.arange(n) // 10).astype(str)), "Antitarget").astype(object)
    gc = np.clip(rng.beta(8, 8, n), 0.2, 0.8)
    rmask = rng.beta(2, 5, n)
    return dict(chromosome=chrom, start=start, end=end, gene=gene, is_target=is_target, gc=gc, rmask=rmask)


def exome_reference(bins, seed: int = 0xC4):
    """Pooled-reference columns on `bins`: log2 ~ N(0, 0.4) with 1 % of bins at -6 (masked by
    fix.mask_bad_bins), spread ~ |N(0.15, 0.05)| with 0.5 % above 1, depth = 100 * 2^log2."""
    rng = np.random.default_rng(seed)
    n = len(bins["start"])
    log2 = rng.normal(0, 0.4, n)
    log2[rng.random(n) < 0.01] = -6.0
    spread = np.abs(rng.normal(0.15, 0.05, n))
    spread[rng.random(n) < 0.005] = 1.5
    return dict(log2=round_sig6(log2), depth=round_sig6(100.0 * np.exp2(log2)), spread=round_sig6(spread))


def exome_sample(bins, ref, seed: int = 1, n_events=None):
    """One tumour on `bins`: log2 = reference + quadratic GC bias (amplitude 0.3) + an edge-density
    term for small targets + N(0, 0.2) + 6-12 planted copy-number segments (step from STEPS,
    20-5000 bins); depth = 100 * 2^log2.  Returns dict(log2, depth, truth)."""
    rng = np.random.default_rng([0xC2, seed])
    n = len(bins["start"])
    size = (bins["end"] - bins["start"]).astype(np.float64)
    gc_bias = 0.3 * (1.0 - ((bins["gc"] - 0.5) / 0.3) ** 2) * rng.uniform(0.5, 1.5)
    edge = np.where(bins["is_target"], -0.2 * np.minimum(1.0, 250.0 / (2.0 * size)), 0.0)
    log2 = ref["log2"] + gc_bias + edge + rng.normal(0, 0.2, n)
    truth = []
    for _ in range(int(rng.integers(6, 13)) if n_events is None else n_events):
        length = int(min(rng.integers(20, 5001), n // 4))
        lo = int(rng.integers(0, n - length))
        step = float(rng.choice(STEPS))
        log2[lo:lo + length] += step
        truth.append((lo, lo + length, step))
    log2 = round_sig6(log2)
    return dict(log2=log2, depth=round_sig6(100.0 * np.exp2(log2)), truth=truth)


def exome_tables(bins, ref, sample, sample_id="tumor"):
    """(target_raw, antitarget_raw, reference) CopyNumArray tables the way `cnvkit.py coverage` and
    `cnvkit.py reference` write them."""
    import pandas as pd

    from .cnary import CopyNumArray

    base = pd.DataFrame({"chromosome": bins["chromosome"], "start": bins["start"], "end": bins["end"],
                         "gene": bins["gene"]})
    cov = base.assign(log2=sample["log2"], depth=sample["depth"])
    t = bins["is_target"]
    target = CopyNumArray(cov[t].reset_index(drop=True), {"sample_id": sample_id})
    anti = CopyNumArray(cov[~t].reset_index(drop=True), {"sample_id": sample_id})
    reference = CopyNumArray(base.assign(log2=ref["log2"], depth=ref["depth"], gc=bins["gc"], rmask=bins["rmask"],
                                         spread=ref["spread"]), {"sample_id": "reference"})
    return target, anti, reference


def exome_normal(bins, bin_effect, seed: int, male: bool = False):
    """One normal sample on `bins` for the reference build (C4): log2 = bin effect + per-sample GC /
    RepeatMasker biases of random amplitude + N(0, 0.15); chrX at -1 and chrY present for males."""
    rng = np.random.default_rng([0xC4, seed])
    n = len(bins["start"])
    gc_b = rng.uniform(0.1, 0.4) * (1.0 - ((bins["gc"] - 0.5) / 0.3) ** 2)
    rm_b = rng.uniform(-0.2, 0.2) * bins["rmask"]
    log2 = bin_effect + gc_b + rm_b + rng.normal(0, 0.15, n)
    chrom = bins["chromosome"]
    if male:
        log2[chrom == "chrX"] -= 1.0
        log2[chrom == "chrY"] -= 1.0
    else:
        log2[chrom == "chrY"] = -18.0
    log2 = round_sig6(log2)
    return dict(log2=log2, depth=round_sig6(100.0 * np.exp2(log2)))

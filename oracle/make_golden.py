#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generate tests/golden/*.npz from the reference.

Run in the authoring container (where /root/reference is mounted):

    python oracle/make_golden.py

Every vector written here is produced by the *unmodified* reference code
(imported through oracle/refshim.py) on the reference's own test fixtures
(test/picard/*.csv, test/formats/*.cnr|cnn) or on seeded synthetic inputs.
The files are small, committed, and are what pins the oracle restatements in
oracle/ (tests -m "not gpu") and the CUDA path (tests -m gpu) on the GPU box,
where /root/reference does not exist.
"""
from __future__ import annotations

import logging
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import refshim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
NUMERIC = ("start", "end", "log2", "depth", "weight", "gc", "rmask", "spread", "probes")


def table_arrays(cna, prefix):
    """Flatten a reference CopyNumArray into {prefix_col: ndarray}."""
    out = {}
    df = cna.data
    for col in df.columns:
        v = df[col].to_numpy()
        if col in ("chromosome", "gene"):
            v = np.asarray(v, dtype=str)
        out[f"{prefix}__{col}"] = v
    out[f"{prefix}__columns"] = np.asarray(list(df.columns), dtype=str)
    return out


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB, {len(arrays)} arrays)")


# ---------------------------------------------------------------------------
def gen_smoothing(R):
    """smoothing.rolling_median / rolling_quantile / savgol known answers."""
    rng = np.random.default_rng(20261019)
    arrays = {}
    cases = []
    for i, (n, width) in enumerate(
        [(2, 0.5), (5, 0.3), (57, 0.1), (100, 3), (1000, 0.05), (1000, 51), (4097, 0.02),
         (18759, 0.0073), (18759, 0.1), (30000, 0.01), (30000, 0.5)]
    ):
        x = rng.normal(0, 1, n)
        if i % 3 == 1:  # heavy ties
            x = np.round(x, 1)
        y = R.smoothing.rolling_median(x, width)
        arrays[f"rm{i}_x"] = x
        arrays[f"rm{i}_y"] = y
        cases.append((n, float(width)))
    arrays["rm_cases"] = np.asarray(cases)
    # NaN-bearing input (pandas skips NaN, min_periods=1, even counts averaged)
    x = rng.normal(0, 1, 500)
    x[rng.integers(0, 500, 60)] = np.nan
    arrays["rm_nan_x"] = x
    arrays["rm_nan_y"] = R.smoothing.rolling_median(x, 0.05)
    # rolling_quantile as used by drop_outliers: width 50, q .95
    for i, n in enumerate((51, 73, 500, 4000)):
        x = np.abs(rng.normal(0, 1, n))
        arrays[f"rq{i}_x"] = x
        arrays[f"rq{i}_y"] = R.smoothing.rolling_quantile(x, 50, 0.95)
    # savgol trend as used by drop_outliers (total_width=50)
    for i, n in enumerate((51, 73, 500, 4000)):
        x = rng.normal(0, 1, n) + np.sin(np.arange(n) / 30.0)
        arrays[f"sg{i}_x"] = x
        arrays[f"sg{i}_y"] = R.smoothing.savgol(x, 50)
        arrays[f"sg{i}_mask"] = np.asarray(R.smoothing.rolling_outlier_quantile(x, 50, 0.95, 3))
    save("smoothing", **arrays)


# ---------------------------------------------------------------------------
def gen_regression(R):
    """import-picard -> do_reference -> do_fix -> do_segmentation('haar').

    Mirrors test/test_regression.py:91-126 and also records the intermediate
    center_by_window / rolling_median / edge-bias calls of do_fix.
    """
    picard = os.path.join(R.root, "test", "picard")
    normals = ("p2-5_5", "p2-9_5", "p2-20_5")
    sample = "p2-9_2"
    arrays = {}
    with tempfile.TemporaryDirectory() as tmp:

        def to_cnn(sid, kind):
            path = os.path.join(tmp, f"{sid}.{kind}coverage.cnn")
            garr = R.importers.do_import_picard(os.path.join(picard, f"{sid}.{kind}coverage.csv"))
            R.tabio.write(garr, path)
            return path

        tfn = [to_cnn(s, "target") for s in normals]
        afn = [to_cnn(s, "antitarget") for s in normals]
        for s, t, a in zip(normals, tfn, afn):
            arrays.update(table_arrays(R.read_cna(t), f"normal_{s}_target"))
            arrays.update(table_arrays(R.read_cna(a), f"normal_{s}_antitarget"))

        # capture the matrices that feed summarize_info
        captured = {}
        orig_summarize = R.reference.summarize_info

        def spy_summarize(all_logr, all_depths):
            captured["all_logr"] = np.array(all_logr)
            captured["all_depths"] = np.array(all_depths)
            return orig_summarize(all_logr, all_depths)

        orig_combine = R.reference.combine_probes

        def spy_combine(filenames, antitarget_fnames, fa_fname, is_haploid_x, diploid_parx_genome, sexes, *a, **k):
            captured["sexes"] = dict(sexes)
            return orig_combine(filenames, antitarget_fnames, fa_fname, is_haploid_x, diploid_parx_genome, sexes,
                                *a, **k)

        R.reference.summarize_info = spy_summarize
        R.reference.combine_probes = spy_combine
        try:
            ref = R.reference.do_reference(tfn, afn, is_haploid_x_reference=True)
        finally:
            R.reference.summarize_info = orig_summarize
            R.reference.combine_probes = orig_combine
        arrays["ref_sex_ids"] = np.asarray(list(captured["sexes"].keys()), dtype=str)
        arrays["ref_sex_is_female"] = np.asarray([bool(v) for v in captured["sexes"].values()])
        arrays["ref_all_logr"] = captured["all_logr"]
        arrays["ref_all_depths"] = captured["all_depths"]
        stats = orig_summarize(captured["all_logr"], captured["all_depths"])
        for k, v in stats.items():
            arrays[f"ref_summary_{k}"] = np.asarray(v)
        arrays.update(table_arrays(ref, "reference"))
        tgt = R.read_cna(to_cnn(sample, "target"))
        anti = R.read_cna(to_cnn(sample, "antitarget"))
    arrays.update(table_arrays(tgt, "sample_target"))
    arrays.update(table_arrays(anti, "sample_antitarget"))

    # spy on center_by_window + edge bias inside do_fix
    cbw_calls = []
    orig_cbw = R.fix.center_by_window
    orig_edge = R.fix.get_edge_bias

    def spy_cbw(cnarr, fraction, sort_key, bias_smoother="median"):
        before = cnarr["log2"].to_numpy().copy()
        key = np.asarray(sort_key, dtype=float).copy()
        out = orig_cbw(cnarr, fraction, sort_key, bias_smoother=bias_smoother)
        cbw_calls.append((before, key, float(fraction), out["log2"].to_numpy().copy()))
        return out

    edge_calls = []

    def spy_edge(cnarr, margin):
        out = orig_edge(cnarr, margin)
        edge_calls.append(
            (
                np.asarray(cnarr.chromosome, dtype=str),
                cnarr.start.to_numpy().copy(),
                cnarr.end.to_numpy().copy(),
                np.asarray(out, dtype=float).copy(),
            )
        )
        return out

    R.fix.center_by_window = spy_cbw
    R.fix.get_edge_bias = spy_edge
    try:
        cnr = R.fix.do_fix(tgt, anti, ref)
    finally:
        R.fix.center_by_window = orig_cbw
        R.fix.get_edge_bias = orig_edge
    for i, (b, k, f, o) in enumerate(cbw_calls):
        arrays[f"cbw{i}_before"] = b
        arrays[f"cbw{i}_key"] = k
        arrays[f"cbw{i}_fraction"] = np.asarray(f)
        arrays[f"cbw{i}_after"] = o
    arrays["cbw_ncalls"] = np.asarray(len(cbw_calls))
    for i, (c, s, e, o) in enumerate(edge_calls):
        arrays[f"edge{i}_chromosome"] = c
        arrays[f"edge{i}_start"] = s
        arrays[f"edge{i}_end"] = e
        arrays[f"edge{i}_bias"] = o
    arrays["edge_ncalls"] = np.asarray(len(edge_calls))
    arrays.update(table_arrays(cnr, "cnr"))
    cns = R.segmentation.do_segmentation(cnr, "haar")
    arrays.update(table_arrays(cns, "cns_haar"))
    # no-outlier-filter variant isolates haar + transfer_fields from drop_outliers
    cns2 = R.segmentation.do_segmentation(cnr, "haar", skip_outliers=0)
    arrays.update(table_arrays(cns2, "cns_haar_nooutlier"))
    save("regression", **arrays)


# ---------------------------------------------------------------------------
def gen_haar(R):
    """Per-arm HaarSeg internals on test/formats/p2-20_1.cnr + KATs."""
    cnr = R.read_cna(os.path.join(R.root, "test", "formats", "p2-20_1.cnr"))
    arrays = table_arrays(cnr, "cnr")
    offsets = [0]
    arm_chrom = []
    for chrom, arm in cnr.by_arm():
        offsets.append(offsets[-1] + len(arm))
        arm_chrom.append(chrom)
    arrays["arm_offsets"] = np.asarray(offsets)
    arrays["arm_chrom"] = np.asarray(arm_chrom, dtype=str)
    H = R.haar
    log2 = cnr["log2"].to_numpy()
    wt = cnr["weight"].to_numpy()
    bp_w, bp_u, n_w, n_u = [], [], [0], [0]
    mean_w, mean_u = [], []
    for a in range(len(arm_chrom)):
        lo, hi = offsets[a], offsets[a + 1]
        rw = H.haarSeg(log2[lo:hi], 1e-4, weights=wt[lo:hi])
        ru = H.haarSeg(log2[lo:hi], 1e-4)
        bp_w.extend(rw["start"].tolist())
        mean_w.extend(rw["mean"].tolist())
        n_w.append(len(bp_w))
        bp_u.extend(ru["start"].tolist())
        mean_u.extend(ru["mean"].tolist())
        n_u.append(len(bp_u))
    arrays["w_seg_start"] = np.asarray(bp_w)
    arrays["w_seg_mean"] = np.asarray(mean_w)
    arrays["w_seg_offsets"] = np.asarray(n_w)
    arrays["u_seg_start"] = np.asarray(bp_u)
    arrays["u_seg_mean"] = np.asarray(mean_u)
    arrays["u_seg_offsets"] = np.asarray(n_u)
    # a looser q so that more breakpoints/levels are exercised
    bp_q, n_q = [], [0]
    for a in range(len(arm_chrom)):
        lo, hi = offsets[a], offsets[a + 1]
        r = H.haarSeg(log2[lo:hi], 0.05, weights=wt[lo:hi])
        bp_q.extend(r["start"].tolist())
        n_q.append(len(bp_q))
    arrays["q05_seg_start"] = np.asarray(bp_q)
    arrays["q05_seg_offsets"] = np.asarray(n_q)
    # conv + peaks of the largest arm, each level
    sizes = np.diff(offsets)
    a = int(np.argmax(sizes))
    lo, hi = offsets[a], offsets[a + 1]
    arrays["big_arm"] = np.asarray(a)
    for level in range(1, 6):
        h = 2**level
        cw_ = H.HaarConv(log2[lo:hi], wt[lo:hi], h)
        cu_ = H.HaarConv(log2[lo:hi], None, h)
        arrays[f"conv_w_L{level}"] = cw_
        arrays[f"conv_u_L{level}"] = cu_
        arrays[f"peaks_w_L{level}"] = H.FindLocalPeaks(cw_)
        arrays[f"peaks_u_L{level}"] = H.FindLocalPeaks(cu_)
    arrays["conv_u_L0"] = H.HaarConv(log2[lo:hi], None, 1)
    # full driver outputs
    arrays.update(table_arrays(R.segmentation.do_segmentation(cnr, "haar"), "cns_haar"))
    arrays.update(
        table_arrays(R.segmentation.do_segmentation(cnr, "haar", threshold=0.01, skip_low=True), "cns_haar_t01")
    )
    # drop_outliers mask per arm (width 50, factor 10 and a stricter factor 3)
    masks10, masks3 = [], []
    for _chrom, arm in cnr.by_arm():
        for factor, dest in ((10, masks10), (3, masks3)):
            parts = [
                np.asarray(R.smoothing.rolling_outlier_quantile(sub["log2"], 50, 0.95, factor))
                for _c, sub in arm.by_chromosome()
            ]
            dest.append(np.concatenate(parts))
    arrays["outlier_mask_f10"] = np.concatenate(masks10)
    arrays["outlier_mask_f3"] = np.concatenate(masks3)
    # KATs of test/test_haar.py:219-230 (planted steps)
    rng = np.random.default_rng(7)
    sig = np.concatenate([np.zeros(100), np.full(100, -2.0), np.full(100, -2.3)]) + rng.normal(0, 0.05, 300)
    r = H.haarSeg(sig, 1e-4)
    arrays["kat_signal"] = sig
    arrays["kat_start"] = r["start"]
    arrays["kat_mean"] = r["mean"]
    # FDRThres edge cases
    x = rng.normal(0, 1, 400)
    x[:10] *= 8
    arrays["fdr_x"] = x
    arrays["fdr_T"] = np.asarray([H.FDRThres(x, q, 1.0) for q in (1e-4, 1e-2, 0.2, 0.5)], dtype=float)
    save("haar", **arrays)


# ---------------------------------------------------------------------------
def gen_fix_units(R):
    """Unit-level vectors: center_all, mask_bad_bins, apply_weights, biweight."""
    rng = np.random.default_rng(99)
    arrays = {}
    ref = R.read_cna(os.path.join(R.root, "test", "formats", "reference-tr.cnn"))
    arrays.update(table_arrays(ref, "ref_tr"))
    arrays["ref_tr_badmask"] = np.asarray(R.fix.mask_bad_bins(ref))
    # biweight location / midvariance on assorted vectors
    vecs = [rng.normal(0, 1, n) for n in (2, 3, 4, 7, 8, 9, 33, 257, 1000)]
    vecs.append(np.r_[rng.normal(0, 1, 50), 40.0, -35.0])
    vecs.append(np.r_[np.zeros(30), rng.normal(0, 1, 10)])  # MAD == 0
    vecs.append(np.full(9, 3.25))
    for i, v in enumerate(vecs):
        arrays[f"bw{i}_x"] = v
        arrays[f"bw{i}_loc"] = np.asarray(R.descriptives.biweight_location(v))
        arrays[f"bw{i}_var"] = np.asarray(R.descriptives.biweight_midvariance(v))
        arrays[f"bw{i}_var_init0"] = np.asarray(R.descriptives.biweight_midvariance(v, initial=0.1))
    arrays["bw_n"] = np.asarray(len(vecs))
    # center_all on a real .cnr, skip_low on/off
    cnr = R.read_cna(os.path.join(R.root, "test", "formats", "p2-20_2.cnr"))
    arrays.update(table_arrays(cnr, "cnr2"))
    for skip in (False, True):
        c = cnr.copy()
        c.data["log2"] = c.data["log2"] + 0.37
        c.center_all(skip_low=skip)
        arrays[f"cnr2_centered_skip{int(skip)}"] = c["log2"].to_numpy()
    save("fix_units", **arrays)


def gen_exome_mini(R):
    """Synthetic exome panel (cnvkit_b200.synth, 9 000 targets + 3 000 antitargets, C2's construction at
    1/16 scale) through the UNMODIFIED reference: do_fix, then do_segmentation('haar') -- pins the fix
    path on a bin set with close-packed targets (edge gains), masked bins and antitargets, which the
    real-data regression fixture exercises less."""
    import pandas as pd

    from cnvkit_b200 import synth

    bins = synth.exome_bins(n_targets=9000, n_antitargets=3000, seed=7)
    ref = synth.exome_reference(bins, seed=8)
    arrays = {}
    t = bins["is_target"]
    base = pd.DataFrame({"chromosome": bins["chromosome"], "start": bins["start"], "end": bins["end"],
                         "gene": bins["gene"]})
    reference = R.cnary.CopyNumArray(base.assign(log2=ref["log2"], depth=ref["depth"], gc=bins["gc"],
                                                 rmask=bins["rmask"], spread=ref["spread"]), {"sample_id": "reference"})
    for k in range(2):
        smp = synth.exome_sample(bins, ref, seed=k, n_events=4)
        cov = base.assign(log2=smp["log2"], depth=smp["depth"])
        tgt = R.cnary.CopyNumArray(cov[t].reset_index(drop=True), {"sample_id": f"S{k}"})
        anti = R.cnary.CopyNumArray(cov[~t].reset_index(drop=True), {"sample_id": f"S{k}"})
        cnr = R.fix.do_fix(tgt, anti, reference)
        cns = R.segmentation.do_segmentation(cnr, "haar")
        arrays[f"s{k}_in_log2_checksum"] = np.asarray([float(np.sum(smp["log2"])), float(np.sum(smp["depth"]))])
        arrays.update(table_arrays(cnr, f"s{k}_cnr"))
        arrays.update(table_arrays(cns, f"s{k}_cns_haar"))
    arrays["ref_checksum"] = np.asarray([float(np.sum(ref["log2"])), float(np.sum(ref["spread"])),
                                         float(np.sum(bins["gc"])), float(np.sum(bins["start"]))])
    save("exome_mini", **arrays)


def gen_segmetrics(R):
    """segmetrics.confidence_interval_bootstrap / do_segmetrics(interval_stats=['ci']) known answers:
    seeded synthetic segments of many sizes, and test/formats/amplicon.cnr + .cns (the tables of
    test_analysis.py:270-311)."""
    from cnvlib import segmetrics  # the unmodified reference

    rng = np.random.default_rng(0x5E6)
    arrays = {}
    cases = []
    sizes = [2, 3, 4, 5, 7, 8, 9, 11, 16, 31, 64, 100, 127, 128, 129, 137, 255, 256, 257, 300, 520, 1000, 2049, 5000]
    for i, k in enumerate(sizes):
        level = rng.choice([-1.0, -0.4, 0.0, 0.3, 0.58])
        v = level + rng.normal(0, 0.25, k)
        if i % 5 == 0:
            v[rng.integers(0, k)] += 3.0         # an outlier makes the jackknife skewed
        w = np.clip(rng.uniform(0.3, 1.0, k), 1e-4, 1)
        arrays[f"unit{i}_values"] = v
        arrays[f"unit{i}_weights"] = w
        for alpha, boots in ((0.05, 100), (0.5, 100), (0.05, 30), (0.01, 100)):
            ci = segmetrics.confidence_interval_bootstrap(v, w, alpha, boots, False)
            arrays[f"unit{i}_ci_a{alpha}_b{boots}"] = np.asarray(ci, dtype=float)
        cases.append(k)
    arrays["unit_sizes"] = np.asarray(cases)
    # one smoothed realisation per size (the reference's noise is unseeded: statistical comparison only)
    for i, k in enumerate(sizes):
        reps = np.array([segmetrics.confidence_interval_bootstrap(arrays[f"unit{i}_values"], arrays[f"unit{i}_weights"],
                                                                  0.5, 100, True) for _ in range(8)])
        arrays[f"unit{i}_ci_smoothed_a0.5_reps"] = reps

    cnr = R.read_cna(os.path.join(R.root, "test", "formats", "amplicon.cnr"))
    cns = R.read_cna(os.path.join(R.root, "test", "formats", "amplicon.cns"))
    arrays.update(table_arrays(cnr, "cnr"))
    arrays.update(table_arrays(cns, "cns"))
    sel = [ser.index.to_numpy() for ser in cnr.iter_ranges_of(cns, "log2", "outer", True)]
    arrays["sel_lo"] = np.asarray([s[0] if len(s) else 0 for s in sel])
    arrays["sel_hi"] = np.asarray([s[-1] + 1 if len(s) else 0 for s in sel])
    assert all(len(s) == 0 or np.array_equal(s, np.arange(s[0], s[-1] + 1)) for s in sel)
    for tag, kw in (("default", dict()), ("bca", dict(smoothed=False)), ("bca_a0.5_b50", dict(smoothed=False, alpha=0.5, bootstraps=50)),
                    ("batch", dict(alpha=0.5, smoothed=True))):
        sm = segmetrics.do_segmetrics(cnr, cns, interval_stats=["ci"], **kw)
        arrays[f"amplicon_{tag}_ci_lo"] = sm["ci_lo"].to_numpy()
        arrays[f"amplicon_{tag}_ci_hi"] = sm["ci_hi"].to_numpy()
    save("segmetrics", **arrays)


def synth_fasta(rng, lengths, width=60):
    """A small soft-masked genome: random ACGT with lowercase repeat runs, N gaps and a few IUPAC codes."""
    lines = []
    for name, L in lengths.items():
        seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L, p=[0.29, 0.21, 0.21, 0.29])
        gc_rich = rng.integers(0, max(L - 5000, 1), size=max(L // 20000, 1))
        for g in gc_rich:
            seq[g:g + 4000] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=len(seq[g:g + 4000]), p=[0.15, 0.35, 0.35, 0.15])
        for _ in range(max(L // 3000, 1)):           # repeats -> lowercase
            a = rng.integers(0, L)
            b = min(L, a + int(rng.geometric(1 / 400)))
            seq[a:b] |= 0x20
        for _ in range(max(L // 40000, 1)):          # assembly gaps
            a = rng.integers(0, L)
            seq[a:a + int(rng.integers(10, 3000))] = ord("N")
        for a in rng.integers(0, L, size=max(L // 5000, 1)):
            seq[a] = rng.choice(np.frombuffer(b"RYKMSWnry", dtype=np.uint8))
        text = seq.tobytes().decode()
        lines.append(f">{name} synthetic")
        lines.extend(text[i:i + width] for i in range(0, L, width))
    return "\n".join(lines) + "\n"


def gen_fasta(R):
    """reference.get_fasta_stats / calculate_gc_lo known answers and a do_reference(fa_fname=...) run on a
    synthetic soft-masked genome (the pyfaidx calls go through oracle/refshim_stubs/pyfaidx.py)."""
    import pandas as pd

    rng = np.random.default_rng(0xFA57A)
    arrays = {}
    seqs = ["", "N" * 50, "ACGT", "acgt", "AAAA", "ggcc", "ACGTNNNNacgtRYKM", "nnnnACGTTGCAacgtnn\n", "GATTACA" * 37]
    seqs += ["".join(rng.choice(list("ACGTacgtNn"), size=int(k))) for k in (1, 2, 15, 16, 17, 31, 33, 64, 100, 1000, 4097)]
    arrays["kat_seqs"] = np.asarray(seqs, dtype=str)
    arrays["kat_gc_lo"] = np.asarray([R.reference.calculate_gc_lo(q) for q in seqs], dtype=float)

    lengths = {"chr1": 150_011, "chr2": 90_000, "chrX": 61_234, "chrY": 20_001}
    text = synth_fasta(rng, lengths)
    arrays["fasta_bytes"] = np.frombuffer(text.encode(), dtype=np.uint8)
    rows = []
    for name, L in lengths.items():
        pos = 200
        while pos < L - 2000:                        # targets
            size = int(np.clip(rng.normal(267, 60), 80, 1000))
            rows.append((name, pos, pos + size, f"G{len(rows) // 7}"))
            pos += size + int(rng.lognormal(6.5, 1.0))
        rows.append((name, L - 150, L + 500, "Edge"))            # runs past the end of the sequence
    tgt = pd.DataFrame(rows, columns=["chromosome", "start", "end", "gene"])
    anti_rows = []
    for name, L in lengths.items():
        for a in range(0, L, 15_000):
            anti_rows.append((name, a, min(a + 14_000, L), "Antitarget"))
    anti = pd.DataFrame(anti_rows, columns=["chromosome", "start", "end", "gene"])
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "genome.fa")
        with open(fa, "w") as handle:
            handle.write(text)
        for tag, bins in (("tgt", tgt), ("anti", anti)):
            probe = R.cnary.CopyNumArray(bins.assign(log2=0.0, depth=1.0), {"sample_id": "probe"})
            gc, rm = R.reference.get_fasta_stats(probe, fa)
            arrays[f"{tag}_gc"], arrays[f"{tag}_rmask"] = gc, rm
            arrays.update(table_arrays(probe, tag))
        # three normals through the unmodified do_reference with the FASTA
        tnames, anames = [], []
        for k in range(3):
            for tag, bins, names, sd in (("targetcoverage", tgt, tnames, 0.25), ("antitargetcoverage", anti, anames, 0.12)):
                gcv = arrays["tgt_gc" if bins is tgt else "anti_gc"]
                log2 = rng.normal(0, sd, len(bins)) + 1.5 * (gcv - 0.45) * (k - 1) + rng.normal(0, 0.5, 1)
                log2[bins["chromosome"].to_numpy() == "chrY"] -= 6.0          # female samples
                cna = R.cnary.CopyNumArray(bins.assign(log2=log2, depth=np.exp2(log2) * 80), {"sample_id": f"N{k}"})
                path = os.path.join(tmp, f"N{k}.{tag}.cnn")
                R.tabio.write(cna, path)
                names.append(path)
                arrays[f"N{k}_{tag}_log2"] = cna["log2"].to_numpy()
                arrays[f"N{k}_{tag}_depth"] = cna["depth"].to_numpy()
        ref = R.reference.do_reference(tnames, anames, fa, female_samples=True)
        arrays.update(table_arrays(ref, "ref"))
        # flat reference from BED files (do_reference_flat, reference.py:26-75)
        t_bed, a_bed = os.path.join(tmp, "targets.bed"), os.path.join(tmp, "antitargets.bed")
        with open(t_bed, "w") as handle:
            handle.write('track name="panel" description="synthetic"\n')
            for c, a, b, gname in tgt.itertuples(index=False):
                handle.write(f"{c}\t{a}\t{b}\t{gname}\t0\t+\n")
        with open(a_bed, "w") as handle:
            for c, a, b, _g in anti.itertuples(index=False):
                handle.write(f"{c}\t{a}\t{b}\n")               # BED3: gene becomes "-"
        arrays["targets_bed"] = np.frombuffer(open(t_bed, "rb").read(), dtype=np.uint8)
        arrays["antitargets_bed"] = np.frombuffer(open(a_bed, "rb").read(), dtype=np.uint8)
        for tag, kw in (("flat_hx", dict(is_haploid_x_reference=True)), ("flat", dict())):
            flat = R.reference.do_reference_flat(t_bed, a_bed, fa, **kw)
            arrays.update(table_arrays(flat, tag))
        arrays.update(table_arrays(R.reference.do_reference_flat(t_bed), "flat_nofa"))
    save("fasta", **arrays)


def gen_sex(R):
    """CopyNumArray.compare_sex_chromosomes (cnary.py:484-632) and reference._reconcile_sex_guesses
    (reference.py:335-382) known answers on synthetic coverage tables."""
    import pandas as pd

    rng = np.random.default_rng(0x5E)
    arrays = {}
    chroms = ["chr1", "chr2", "chr3", "chrX", "chrY"]
    sizes = [400, 300, 251, 180, 41]
    cases = []
    for i, (x_shift, y_shift, drop_y, haploid, parx) in enumerate([
            (0.0, -18.0, False, False, None), (-1.0, 0.0, False, False, None), (-0.45, -9.0, False, False, None),
            (-0.55, -10.5, False, False, None), (0.0, 0.0, True, False, None), (-1.0, 0.0, True, False, None),
            (1.0, -19.0, False, True, None), (0.0, 0.2, False, True, None), (-1.0, -0.3, False, False, "grch38"),
            (-0.9, -19.5, False, False, None), (0.1, 0.3, False, False, None)]):
        rows = []
        for c, m in zip(chroms, sizes):
            if c == "chrY" and drop_y:
                continue
            pos = np.sort(rng.choice(np.arange(10_000, 150_000_000, 1000), size=m, replace=False))
            if c in ("chrX", "chrY"):
                pos[:6] = np.arange(6) * 20_000 + 100_000          # a few bins inside PAR1
            shift = x_shift if c == "chrX" else y_shift if c == "chrY" else 0.0
            for p in pos:
                rows.append((c, int(p), int(p) + 500, "G", shift + rng.normal(0, 0.3)))
        df = pd.DataFrame(rows, columns=["chromosome", "start", "end", "gene", "log2"])
        df["depth"] = np.exp2(df["log2"]) * 50
        cna = R.cnary.CopyNumArray(df, {"sample_id": f"S{i}"})
        is_xy, stats = cna.compare_sex_chromosomes(haploid, parx)
        arrays.update(table_arrays(cna, f"case{i}"))
        arrays[f"case{i}_is_xy"] = np.asarray(-1 if is_xy is None else int(is_xy))
        arrays[f"case{i}_stats"] = np.asarray([stats.get(k, np.nan) for k in ("chrx_ratio", "chry_ratio", "chrx_male_lr", "chry_male_lr")], dtype=float)
        arrays[f"case{i}_haploid"] = np.asarray(int(haploid))
        arrays[f"case{i}_parx"] = np.asarray(parx or "")
        cases.append(i)
    arrays["n_cases"] = np.asarray(len(cases))
    # reconcile: (target guess, antitarget guess) -> final is_xx
    recon_in = [((True, 0.2), (True, 0.3)), ((True, 0.9), (False, 5.0)), ((False, 4.0), (True, 0.8)), ((False, 1.2), (True, 0.1)),
                ((True, 0.5), (False, 2.0)), ((True, float("nan")), (False, 3.0)), ((False, 2.0), (True, float("inf"))),
                ((True, 0.7), None), (None, (False, 2.5))]
    outs = []
    for k, (t, a) in enumerate(recon_in):
        tg = {f"s{k}": (np.bool_(t[0]), t[1])} if t else {}
        ag = {f"s{k}": (np.bool_(a[0]), a[1])} if a else {}
        outs.append(int(bool(R.reference._reconcile_sex_guesses(tg, ag)[f"s{k}"])))
    arrays["recon_t"] = np.asarray([[-1, np.nan] if t is None else [int(t[0]), t[1]] for t, _ in recon_in], dtype=float)
    arrays["recon_a"] = np.asarray([[-1, np.nan] if a is None else [int(a[0]), a[1]] for _, a in recon_in], dtype=float)
    arrays["recon_out"] = np.asarray(outs)
    save("sex", **arrays)


def synth_variants(cnr, seed=0xBAF):
    """Heterozygous SNPs on the bins of `cnr` (about every second bin, depth ~ Poisson(60)); two chromosomes carry a
    copy-neutral LOH stretch (allele fraction 0.2) that the depth signal does not show."""
    import pandas as pd

    rng = np.random.default_rng(seed)
    d = cnr.data
    pick = rng.random(len(d)) < 0.5
    chrom = d["chromosome"].to_numpy()[pick]
    pos = ((d["start"].to_numpy() + d["end"].to_numpy()) // 2)[pick]
    depth = rng.poisson(60, pick.sum()).astype(np.int64) + 8
    frac = np.full(pick.sum(), 0.5)
    chroms = list(dict.fromkeys(chrom.tolist()))
    for c in chroms[2:4]:
        idx = np.flatnonzero(chrom == c)
        frac[idx[len(idx) // 3: 2 * len(idx) // 3]] = 0.2
    alt = rng.binomial(depth, frac)
    return pd.DataFrame({"chromosome": chrom, "start": pos, "end": pos + 1, "ref": "A", "alt": "G", "zygosity": 0.5,
                         "depth": depth, "alt_count": alt, "alt_freq": alt / depth})


def gen_haar_baf(R):
    """do_segmentation(cnr, 'haar', variants=...) of the UNMODIFIED reference (haar.py:188-225 depth + BAF breakpoint
    union; segmentation/__init__.py:359 per-segment BAF) on test/formats/p2-20_1.cnr with synthetic het SNPs, once
    with allele counts and once with allele frequencies only."""
    from cnvlib.vary import VariantArray

    cnr = R.read_cna(os.path.join(R.root, "test", "formats", "p2-20_1.cnr"))
    vdf = synth_variants(cnr)
    arrays = {f"var__{c}": (vdf[c].to_numpy().astype(str) if vdf[c].dtype == object or str(vdf[c].dtype) in ("str", "string")
                            else vdf[c].to_numpy()) for c in vdf.columns}
    arrays["var__columns"] = np.asarray(list(vdf.columns), dtype=str)
    cns_counts = R.segmentation.do_segmentation(cnr, "haar", variants=VariantArray(vdf.copy()))
    arrays.update(table_arrays(cns_counts, "cns_counts"))
    vfreq = vdf.drop(columns=["alt_count", "depth"])
    cns_freq = R.segmentation.do_segmentation(cnr, "haar", variants=VariantArray(vfreq.copy()))
    arrays.update(table_arrays(cns_freq, "cns_freq"))
    plain = R.segmentation.do_segmentation(cnr, "haar")
    arrays["n_plain"] = np.asarray(len(plain))
    print("haar_baf: segments plain", len(plain), "counts", len(cns_counts), "freq", len(cns_freq))
    save("haar_baf", **arrays)


def main():
    logging.basicConfig(level=logging.ERROR)
    os.makedirs(OUT, exist_ok=True)
    R = refshim.load()
    which = sys.argv[1:] or ["smoothing", "regression", "haar", "fix_units", "exome_mini", "segmetrics", "fasta", "sex"]
    for w in which:
        globals()["gen_" + w](R)


if __name__ == "__main__":
    main()

"""Pooled reference build on the GPU -- host mirror of cnvlib/reference.py (numeric core).

``combine_probes`` / ``do_reference`` keep the reference's contract: they take ``.cnn`` *file
paths*, return the reference table (chromosome, start, end, gene, log2, depth, [gc], spread) and
raise the same errors.  On the device: per-normal ``bias_correct_logr`` (centre, sex-chromosome
shift, GC / RepeatMasker / edge centring by window at fraction 0.1) and ``summarize_info`` (Tukey
biweight location and midvariance of every bin across the S+1 samples, one warp per bin).

``fa_fname`` adds the GC / RepeatMasker columns by counting letters of the FASTA file on the device
(``cnvkit_b200.fasta``, reference.py:859-895).  Outside the hot path and therefore *not* reimplemented
(SURVEY.md section 2): sample clustering (``do_cluster``).  With ``female_samples=None`` each sample's sex
is inferred from its coverage as the reference does (medians on the device, reference.py:284-382).
"""
from __future__ import annotations

import ctypes
import logging
import os

import numpy as np
import pandas as pd
import torch

from . import _dev, _lib, cnary, genome, params, smoothing, tabio


def fbase(fname: str) -> str:
    """Strip directory and extensions (cnvlib/core.py:118-140)."""
    base = os.path.basename(fname)
    if base.endswith(".gz"):
        base = base[:-3]
    for ext in (".antitargetcoverage.cnn", ".targetcoverage.cnn", ".antitargetcoverage.csv", ".targetcoverage.csv",
                ".recal.bam", ".deduplicated.realign.bam"):
        if base.endswith(ext):
            return base[: -len(ext)]
    return base.rsplit(".", 1)[0]


def _read(fname):
    arr = cnary.read(fname, sample_id=fbase(fname))
    return arr


def read_bed(bed_fname):
    """skgenome/tabio/bedio.py:17-57 read_bed: chromosome, start, end, [gene] of a UCSC BED file
    (0-based), up to the second "track" line; the gene is "-" where the file has no name column."""
    rows = []
    with open(bed_fname) as handle:
        first = True
        for line in handle:
            if first:
                first = False
                if line.startswith("browser "):
                    first = True
                    continue
                if line.startswith("track"):
                    continue
            elif line.startswith("track"):
                break
            fields = line.split("\t", 6)
            if len(fields) < 3:
                raise ValueError(f"Bad line in {bed_fname}: {line!r}")
            gene = fields[3].rstrip() if len(fields) >= 4 else "-"
            rows.append((fields[0], int(fields[1]), int(fields[2]), gene))
    return pd.DataFrame.from_records(rows, columns=["chromosome", "start", "end", "gene"])


def bed2probes(bed_fname):
    """Create a neutral-coverage CopyNumArray from a file of regions (cnvlib/reference.py:78-86).
    The reference sniffs the format (tabio.read_auto); here BED files and CNVkit's own tab format are read."""
    with open(bed_fname) as handle:
        head = handle.readline()
    if head.startswith("chromosome\tstart\tend"):
        table = tabio.read(bed_fname).data.loc[:, ["chromosome", "start", "end", "gene"]]
    elif str(bed_fname).endswith(".bed") or head.startswith(("track", "browser ")):
        table = read_bed(bed_fname)
    else:
        raise NotImplementedError(f"{bed_fname}: only BED and CNVkit tab region files are read on this path")
    table = table.assign(log2=0.0, spread=0.0)
    probes = cnary.CopyNumArray(table, {"sample_id": fbase(bed_fname)})
    probes.sort()
    return probes


def do_reference_flat(targets, antitargets=None, fa_fname=None, is_haploid_x_reference=False,
                      diploid_parx_genome=None):
    """Compile a neutral-coverage reference from the given intervals (cnvlib/reference.py:26-75): flat
    log2 (-1 on the sex chromosomes a haploid-X reference expects at half coverage), depth = 2^log2, and
    GC / RepeatMasker content counted from the FASTA on the device."""
    ref_probes = bed2probes(targets)
    if antitargets:
        ref_probes.add(bed2probes(antitargets))
    d = ref_probes.data
    chrom = d["chromosome"].to_numpy()
    start = d["start"].to_numpy(dtype=np.int64)
    end = d["end"].to_numpy(dtype=np.int64)
    is_x, is_y, is_y_all = _sex_chrom_rows(chrom, start, end, diploid_parx_genome)
    flat = np.zeros(len(d))                                  # expect_flat_log2 (cnary.py:635-664)
    flat[(is_x | is_y) if is_haploid_x_reference else is_y_all] = -1.0
    ref_probes["log2"] = flat
    ref_probes["depth"] = np.exp2(flat)
    if fa_fname:
        from . import fasta

        gc, rmask = fasta.get_fasta_stats(ref_probes, fa_fname)
        ref_probes["gc"] = gc
        ref_probes["rmask"] = rmask
    else:
        logging.info("No FASTA reference genome provided; skipping GC, RM calculations")
    ref_probes.sort_columns()
    return ref_probes


def do_reference(
    target_fnames,
    antitarget_fnames=None,
    fa_fname=None,
    is_haploid_x_reference=False,
    diploid_parx_genome=None,
    female_samples=None,
    do_gc=True,
    do_edge=True,
    do_rmask=True,
    do_cluster=False,
    min_cluster_size=4,
    cluster_method="hierarchical",
    bias_smoother="median",
    sexes=None,
):
    """Compile a pooled coverage reference from normal samples (cnvlib/reference.py:88-281).

    ``sexes`` ({sample_id: is_female}) is this package's hook for callers that ran the reference's
    own sex inference; with ``female_samples`` given it is derived exactly as the reference does.
    """
    if antitarget_fnames and len(target_fnames) != len(antitarget_fnames):
        raise ValueError(
            f"Unequal number of target and antitarget files given: targets = {len(target_fnames)!r}, "
            f"antitargets = {len(antitarget_fnames)!r}")
    if do_cluster:
        raise NotImplementedError("do_cluster=True is outside the GPU hot path")
    if bias_smoother != "median":
        raise NotImplementedError("only bias_smoother='median' runs on the GPU path")
    if not fa_fname:
        logging.info("No FASTA reference genome provided; skipping GC, RM calculations")
    if sexes is None:
        if female_samples is None:
            # reference.py:226-239: infer from coverage, targets reconciled with antitargets
            logging.info("Sample sex not provided; inferring from samples. ")
            t_guesses = _guess_sexes(target_fnames, False, diploid_parx_genome)
            if antitarget_fnames:
                a_guesses = _guess_sexes(antitarget_fnames, False, diploid_parx_genome)
                sexes = _reconcile_sex_guesses(t_guesses, a_guesses)
            else:
                sexes = {sid: is_xx for sid, (is_xx, _lr) in t_guesses.items()}
        else:
            logging.info(f"Provided sample sex is {'female' if female_samples else 'male'}. ")
            sexes = {fbase(f): female_samples for f in target_fnames}
    ref = combine_probes(target_fnames, antitarget_fnames, fa_fname, is_haploid_x_reference, diploid_parx_genome,
                         sexes, do_gc, do_edge, do_rmask, do_cluster, min_cluster_size)
    return ref


def compare_sex_chromosomes(cnarr, is_haploid_x_reference=False, diploid_parx_genome=None):
    """CopyNumArray.compare_sex_chromosomes (cnvlib/cnary.py:484-632) for a coverage table without a
    weight column: medians of log2 over the autosomes, chrX and chrY (exact radix-select medians on the
    device), "maleness" ratios against the expected positions, AND-gate of the two.  Returns
    (is_male or None, stats)."""
    if not len(cnarr):
        return None, {}
    d = cnarr.data
    if "weight" in d.columns:
        raise NotImplementedError("weighted sex-chromosome medians (a 'weight' column) stay with cnvlib")
    chrom = d["chromosome"].to_numpy()
    start = d["start"].to_numpy(dtype=np.int64)
    end = d["end"].to_numpy(dtype=np.int64)
    log2 = d["log2"].to_numpy(dtype=np.float64)
    if np.isnan(log2).any():
        raise NotImplementedError("NaN log2 values in a coverage table")
    is_x, is_y, _ = _sex_chrom_rows(chrom, start, end, diploid_parx_genome)
    if not is_x.any():
        return None, {}
    auto = genome.autosome_row_mask(chrom, start, end, diploid_parx_genome)
    n = len(log2)
    d_x = _dev.up(log2)
    d_off = _dev.up(np.array([0, n], dtype=np.int64))
    meds = []
    for mask in (auto, is_x, is_y):
        if not mask.any():
            meds.append(np.nan)
            continue
        d_med, d_cnt = _dev.empty(1, torch.float64), _dev.empty(1, torch.int32)
        d_mask = _dev.up(mask.astype(np.uint8))      # bound to a name: must outlive the enqueued kernel
        _lib.call("cnvk_segmented_median_dev", _dev.ptr(d_x), _dev.ptr(d_mask), _dev.ptr(d_off), 1,
                  _dev.ptr(d_med), _dev.ptr(d_cnt), _dev.stream())
        meds.append(float(d_med.cpu()[0]))
    auto_med, chrx_med, chry_med = meds

    def maleness(observed_med, female_shift, male_shift):
        ratio = observed_med - auto_med
        return abs(ratio + female_shift) / max(abs(ratio + male_shift), 1e-6)

    female_x_shift, male_x_shift = (-1, 0) if is_haploid_x_reference else (0, +1)
    chrx_male_lr = maleness(chrx_med, female_x_shift, male_x_shift)
    if is_y.any():
        chry_male_lr = maleness(chry_med, -params.NULL_LOG2_COVERAGE, 0.0)
        is_male = chrx_male_lr > 1.0 and chry_male_lr > 1.0
        chry_ratio = chry_med - auto_med
    else:
        chry_male_lr = np.nan
        chry_ratio = np.nan
        is_male = chrx_male_lr > 1.0
    return bool(is_male), {"chrx_ratio": chrx_med - auto_med, "chry_ratio": chry_ratio,
                           "chrx_male_lr": chrx_male_lr, "chry_male_lr": chry_male_lr}


def _guess_sexes(cnn_fnames, is_haploid_x, diploid_parx_genome, verbose=True):
    """cnvlib/reference.py:284-321: {sample_id: (is_xx, chrx_male_lr)} where sex is determinable."""
    guesses = {}
    for fname in cnn_fnames:
        cnarr = _read(fname)
        if not len(cnarr):
            continue
        is_xy, stats = compare_sex_chromosomes(cnarr, is_haploid_x, diploid_parx_genome)
        if is_xy is None:
            continue
        if verbose:
            x_label, y_label = genome.infer_sex_chrom_labels(list(dict.fromkeys(cnarr.data["chromosome"].tolist())))
            logging.info("Relative log2 coverage of %s=%.3g, %s=%.3g (maleness chrX=%.3g, chrY=%.3g) --> assuming %s",
                         x_label or "chrX", stats["chrx_ratio"], y_label or "chrY", stats["chry_ratio"],
                         stats["chrx_male_lr"], stats["chry_male_lr"], "male" if is_xy else "female")
        guesses[cnarr.sample_id] = (not is_xy, stats["chrx_male_lr"])
    return guesses


def _chrx_maleness_confidence(chrx_male_lr):
    """cnvlib/reference.py:324-332."""
    if not (chrx_male_lr > 0.0 and np.isfinite(chrx_male_lr)):
        return 0.0
    return abs(float(np.log(chrx_male_lr)))


def _reconcile_sex_guesses(target_guesses, antitarget_guesses):
    """cnvlib/reference.py:335-382: the target call stands unless the antitargets' chrX signal is strictly
    more decisive, in which case only their chrX call is adopted."""
    sexes = {sid: is_xx for sid, (is_xx, _lr) in target_guesses.items()}
    for sid, (a_is_xx, a_lr) in antitarget_guesses.items():
        if sid not in target_guesses:
            sexes[sid] = a_is_xx
            continue
        t_is_xx, t_lr = target_guesses[sid]
        if t_is_xx == a_is_xx:
            continue
        if _chrx_maleness_confidence(a_lr) > _chrx_maleness_confidence(t_lr):
            sexes[sid] = a_lr <= 1.0
            preferred = "antitargets"
        else:
            sexes[sid] = t_is_xx
            preferred = "targets"
        logging.warning("Sample %s chromosomal X/Y ploidy looks like %s in targets but %s in antitargets; "
                        "preferring %s (more decisive chrX coverage)", sid, "female" if t_is_xx else "male",
                        "female" if a_is_xx else "male", preferred)
    return sexes


def combine_probes(filenames, antitarget_fnames, fa_fname, is_haploid_x, diploid_parx_genome, sexes, fix_gc,
                   fix_edge, fix_rmask, do_cluster=False, min_cluster_size=4):
    """cnvlib/reference.py:385-494."""
    ref_df, d_logr, d_depths = load_sample_block(filenames, fa_fname, is_haploid_x, diploid_parx_genome, sexes, True,
                                                 fix_gc, fix_edge, False)
    if antitarget_fnames:
        anti_df, a_logr, a_depths = load_sample_block(antitarget_fnames, fa_fname, is_haploid_x, diploid_parx_genome,
                                                      sexes, False, fix_gc, False, fix_rmask)
        if not anti_df.empty:
            ref_df = pd.concat([ref_df, anti_df], ignore_index=True)
        d_logr = torch.cat([d_logr, a_logr], dim=1).contiguous()
        d_depths = torch.cat([d_depths, a_depths], dim=1).contiguous()
    stats = summarize_info_device(d_logr, d_depths)
    ref_df = ref_df.assign(**stats)
    ref = cnary.CopyNumArray(ref_df, {"sample_id": "reference"})
    ref.sort()
    ref.sort_columns()
    return ref


def _sex_chrom_rows(chrom, start, end, diploid_parx_genome):
    """chr_x_filter / chr_y_filter (cnvlib/cnary.py:89-153): X and Y rows, minus PAR when a build is named."""
    x_label, y_label = genome.infer_sex_chrom_labels(list(dict.fromkeys(chrom.tolist())))
    is_x = (chrom == x_label) if x_label is not None else np.zeros(len(chrom), dtype=bool)
    is_y = (chrom == y_label) if y_label is not None else np.zeros(len(chrom), dtype=bool)
    is_y_all = is_y.copy()
    if diploid_parx_genome is not None:
        regs = genome.PAR_REGIONS[diploid_parx_genome.lower()]
        for label, mask, keys in ((x_label, is_x, ("PAR1X", "PAR2X")), (y_label, is_y, ("PAR1Y", "PAR2Y"))):
            if label is None:
                continue
            (a1, b1), (a2, b2) = regs[keys[0]], regs[keys[1]]
            par = (chrom == label) & (((start >= a1) & (end <= b1)) | ((start >= a2) & (end <= b2)))
            mask &= ~par
    return is_x, is_y, is_y_all


def load_sample_block(filenames, fa_fname, is_haploid_x, diploid_parx_genome, sexes, skip_low, fix_gc, fix_edge,
                      fix_rmask):
    """cnvlib/reference.py:497-622; the matrices stay on the device:
    returns (ref_df, logr[S+1, n] CUDA float64 with row 0 = flat pseudo-sample, depths[S, n])."""
    filenames = sorted(filenames, key=fbase)
    logging.info("Loading %s", filenames[0])
    cnarr1 = _read(filenames[0])
    if len(cnarr1) == 0:
        cols = ["chromosome", "start", "end", "gene", "log2", "depth"]
        if "gc" in cnarr1 or fa_fname:
            cols.append("gc")
        if fa_fname:
            cols.append("rmask")
        cols.append("spread")
        dev = _dev.device()
        return (pd.DataFrame.from_records([], columns=cols),
                torch.zeros((len(filenames) + 1, 0), dtype=torch.float64, device=dev),
                torch.zeros((len(filenames), 0), dtype=torch.float64, device=dev))
    d1 = cnarr1.data
    n = len(d1)
    chrom = d1["chromosome"].to_numpy()
    start = d1["start"].to_numpy(dtype=np.int64)
    end = d1["end"].to_numpy(dtype=np.int64)
    ref_columns = {"chromosome": d1["chromosome"], "start": d1["start"], "end": d1["end"], "gene": d1["gene"]}
    gc_col = rm_col = None
    if fa_fname and (fix_rmask or fix_gc):         # reference.py:552-557
        from . import fasta

        gc_vals, rm_vals = fasta.get_fasta_stats(cnarr1, fa_fname)
        if fix_gc:
            gc_col = ref_columns["gc"] = gc_vals
        if fix_rmask:
            rm_col = ref_columns["rmask"] = rm_vals
    elif "gc" in cnarr1 and fix_gc:                # reuse .cnn GC values (import-picard)
        ref_columns["gc"] = d1["gc"]
        gc_col = d1["gc"].to_numpy(dtype=np.float64)
    names, offsets, chrom_id = genome.chrom_runs(chrom)
    is_x, is_y, is_y_all = _sex_chrom_rows(chrom, start, end, diploid_parx_genome)
    # expect_flat_log2 (cnary.py:635-664)
    flat = np.zeros(n)
    flat[(is_x | is_y) if is_haploid_x else is_y_all] = -1.0
    auto = genome.autosome_row_mask(chrom, start, end, diploid_parx_genome)

    d_off, d_cid = _dev.up(offsets), _dev.up(chrom_id)
    d_start, d_end = _dev.up(start), _dev.up(end)
    d_auto = _dev.up(auto.astype(np.uint8))
    d_flat = _dev.up(flat)
    d_xy = _dev.up((is_x.astype(np.uint8) | (is_y.astype(np.uint8) << 1)))
    d_gc = _dev.up(np.asarray(gc_col, dtype=np.float64)) if gc_col is not None else None
    d_rm = _dev.up(np.asarray(rm_col, dtype=np.float64)) if rm_col is not None else None
    d_edge = None
    if fix_edge:
        d_edge = _dev.empty(n, torch.float64)
        _lib.call("cnvk_edge_bias_dev", _dev.ptr(d_start), _dev.ptr(d_end), _dev.ptr(d_cid), n, params.INSERT_SIZE,
                  _dev.ptr(d_edge), _dev.stream())
    wing = smoothing._width2wing(0.1, n) if n >= 2 else 1
    d_perm = _dev.shuffle_permutation(n)

    key_cols = d1.loc[:, ("chromosome", "start", "end", "gene")].values
    # Fast path for files 2..S: when the first file is already in sorted order with no dropped rows,
    # the other files need only their numeric columns -- the native reader returns them together with
    # a hash of the coordinate/gene fields, which replaces the row-by-row key comparison
    # (reference.py:594-598) without building a table.
    fast_hash = None
    num_cols = ("log2", "depth") if "depth" in cnarr1 else ("log2",)
    if len(filenames) > 1:
        try:
            raw, h0 = tabio.read_columns(filenames[0], columns=num_cols + ("start", "end"))
            if (len(raw["log2"]) == n and np.array_equal(raw["log2"], d1["log2"].to_numpy(dtype=np.float64))
                    and np.array_equal(raw["start"], start.astype(np.float64))
                    and np.array_equal(raw["end"], end.astype(np.float64))):
                fast_hash = h0
        except (KeyError, _lib.CnvkError):
            fast_hash = None
    rows_logr = [d_flat.clone()]
    rows_depth = []

    def _correct(d_log2, d_depth, sample_id):
        known = sexes.get(sample_id)
        if known is None:
            logging.warning("Cannot determine sex for sample %s; treating chrX as diploid (female)",
                            sample_id or "")
        is_xx = True if known is None else bool(known)
        skipped = ctypes.c_int32(0)
        _lib.call("cnvk_bias_correct_logr_dev", _dev.ptr(d_log2), _dev.ptr(d_depth), _dev.ptr(d_auto), n,
                  _dev.ptr(d_off), len(names), _dev.ptr(d_gc), _dev.ptr(d_rm), _dev.ptr(d_edge), int(bool(skip_low)), wing,
                  _dev.ptr(d_perm), _dev.ptr(d_flat), _dev.ptr(d_xy), int(is_xx), ctypes.addressof(skipped),
                  _dev.stream())
        if skipped.value:
            logging.warning("WARNING: most bins have no or very low coverage; check that the right BED file was used")
        rows_logr.append(d_log2)

    for i, fname in enumerate(filenames):
        sample_id = fbase(fname)
        if i == 0:
            log2 = d1["log2"].to_numpy(dtype=np.float64)
            depth = d1["depth"].to_numpy(dtype=np.float64) if "depth" in cnarr1 else None
        else:
            logging.info("Loading %s", fname)
            raw = None
            if fast_hash is not None:
                try:
                    raw, h = tabio.read_columns(fname, columns=num_cols)
                except (KeyError, _lib.CnvkError):
                    raw, h = None, None
                if raw is not None and (h != fast_hash or len(raw["log2"]) != n or np.isnan(raw["log2"]).any()):
                    raw = None
            if raw is not None:
                log2, depth = raw["log2"], raw.get("depth")
                d_log2 = _dev.up(log2)
                d_depth = _dev.up(depth) if depth is not None else None
                rows_depth.append(d_depth if d_depth is not None else torch.exp2(d_log2))
                _correct(d_log2, d_depth, sample_id)
                continue
            # the byte-level hash differs (formatting, row order, a missing column, rows without log2): compare the
            # PARSED keys as the reference does (reference.py:594-598) before declaring a mismatch
            cn = _read(fname)
            if not np.array_equal(key_cols, cn.data.loc[:, ("chromosome", "start", "end", "gene")].values):
                raise RuntimeError(f"{fname} bins do not match those in {filenames[0]}")
            log2 = cn.data["log2"].to_numpy(dtype=np.float64)
            depth = cn.data["depth"].to_numpy(dtype=np.float64) if "depth" in cn else None
        d_log2 = _dev.up(log2)
        d_depth = _dev.up(depth) if depth is not None else None
        rows_depth.append(d_depth if d_depth is not None else torch.exp2(d_log2))
        _correct(d_log2, d_depth, sample_id)
    ref_df = pd.DataFrame.from_dict(ref_columns)
    return ref_df, torch.stack(rows_logr).contiguous(), torch.stack(rows_depth).contiguous()


def summarize_info_device(d_logr, d_depths):
    """cnvlib/reference.py:717-741 on device matrices [samples, bins]."""
    ncols = d_logr.shape[1]
    logging.info("Calculating average bin coverages")
    d_loc, d_var, d_dloc = (_dev.empty(max(ncols, 1), torch.float64) for _ in range(3))
    if ncols:
        _lib.call("cnvk_biweight_columns_dev", _dev.ptr(d_logr), d_logr.shape[0], ncols, 3, _dev.ptr(d_loc),
                  _dev.ptr(d_var), _dev.stream())
        _lib.call("cnvk_biweight_columns_dev", _dev.ptr(d_depths), d_depths.shape[0], ncols, 1, _dev.ptr(d_dloc), 0,
                  _dev.stream())
    logging.info("Calculating bin spreads")
    return {"log2": d_loc[:ncols].cpu().numpy(), "depth": d_dloc[:ncols].cpu().numpy(),
            "spread": d_var[:ncols].cpu().numpy()}


def summarize_info(all_logr, all_depths):
    """numpy in / numpy out seam of cnvlib/reference.py:717-741."""
    d_logr = _dev.up(np.ascontiguousarray(all_logr, dtype=np.float64))
    d_dep = _dev.up(np.ascontiguousarray(all_depths, dtype=np.float64))
    return summarize_info_device(d_logr, d_dep)

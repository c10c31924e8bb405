"""do_fix on the GPU -- host mirror of cnvlib/fix.py.

Strings (chromosome, gene), bin matching and row filtering stay on the host; every numeric step
-- mask_bad_bins, center_all, center_by_window (GC / edge / RepeatMasker), the log2 ratio,
apply_weights and the final centring -- runs in csrc/fix.cu behind the C ABI.  Same signature,
return type, error text and log messages as the reference.
"""
from __future__ import annotations

import ctypes
import logging

import numpy as np
import pandas as pd
import torch

from . import _dev, _lib, genome, params, smoothing


def do_fix(
    target_raw,
    antitarget_raw,
    reference,
    diploid_parx_genome=None,
    do_gc=True,
    do_edge=True,
    do_rmask=True,
    do_cluster=False,
    smoothing_window_fraction=None,
    bias_smoother="median",
):
    """Normalise a sample against a pooled reference and correct biases (cnvlib/fix.py:20-215).

    Accepts and returns CopyNumArray objects (this package's or the reference's own class).
    """
    if bias_smoother != "median":
        if bias_smoother == "loess":
            raise NotImplementedError("bias_smoother='loess' (statsmodels LOWESS) is outside the GPU hot path")
        raise ValueError(f"Unknown bias_smoother {bias_smoother!r}; expected 'median' or 'loess'")
    if do_cluster:
        raise NotImplementedError("do_cluster=True (clustered references) is outside the GPU hot path")
    logging.info("Processing target: %s", target_raw.sample_id)
    cnarr, ref_matched = load_adjust_coverages(
        target_raw, reference, True, do_gc, do_edge, False, diploid_parx_genome,
        smoothing_window_fraction=smoothing_window_fraction)
    logging.info("Processing antitarget: %s", antitarget_raw.sample_id)
    anti_cnarr, ref_anti = load_adjust_coverages(
        antitarget_raw, reference, False, do_gc, False, do_rmask, diploid_parx_genome,
        smoothing_window_fraction=smoothing_window_fraction)
    if len(anti_cnarr):
        # fix.py:171-174: concat + genomic (stable) sort of both tables
        cdf = pd.concat([cnarr.data, anti_cnarr.data], ignore_index=True)
        rdf = pd.concat([ref_matched.data, ref_anti.data], ignore_index=True)
        order = genome.genomic_order(cdf["chromosome"].to_numpy(), cdf["start"].to_numpy(), cdf["end"].to_numpy())
        cnarr = cnarr.as_dataframe(cdf.iloc[order].reset_index(drop=True))
        rorder = genome.genomic_order(rdf["chromosome"].to_numpy(), rdf["start"].to_numpy(), rdf["end"].to_numpy())
        ref_matched = ref_matched.as_dataframe(rdf.iloc[rorder].reset_index(drop=True))
    weights, log2 = _finish(cnarr, ref_matched, diploid_parx_genome)
    out = cnarr.data.assign(log2=log2, weight=weights)
    return cnarr.as_dataframe(out)


def load_adjust_coverages(cnarr, ref_cnarr, skip_low, fix_gc, fix_edge, fix_rmask, diploid_parx_genome,
                          smoothing_window_fraction=None):
    """Filter one bin set against the reference and bias-correct it (cnvlib/fix.py:218-292)."""
    if "gc" in cnarr:
        cnarr = cnarr.keep_columns(("chromosome", "start", "end", "gene", "log2", "depth"))
    if not len(cnarr):
        return cnarr, ref_cnarr[:0]
    ref_matched = match_ref_to_sample(ref_cnarr, cnarr)

    dev_ref = {c: _dev.up(ref_matched.data[c].to_numpy(dtype=np.float64))
               for c in ("log2", "spread", "depth", "gc", "rmask") if c in ref_matched}
    n0 = len(cnarr)
    d_log2 = _dev.up(cnarr.data["log2"].to_numpy(dtype=np.float64))
    d_keep = _dev.empty(n0, torch.uint8)
    _lib.call("cnvk_fix_mask_dev", _dev.ptr(d_log2), _dev.ptr(dev_ref["log2"]), _dev.ptr(dev_ref["spread"]),
              _dev.ptr(dev_ref.get("depth")), _dev.ptr(dev_ref.get("gc")), n0, _dev.ptr(d_keep), _dev.stream())
    keep = d_keep.cpu().numpy().astype(bool)
    logging.info("Keeping %d of %d bins", int(keep.sum()), n0)
    cnarr = cnarr[keep]
    ref_matched = ref_matched[keep]
    n = len(cnarr)
    if n == 0:
        return cnarr, ref_matched
    # genomic order is required for the per-chromosome runs (read_cna sorts; do it for API callers)
    chrom = cnarr.data["chromosome"].to_numpy()
    kept_rows = np.flatnonzero(keep)            # rows of the uploaded (pre-filter) columns, in table order
    try:
        names, offsets, chrom_id = genome.chrom_runs(chrom)
    except ValueError:
        order = genome.genomic_order(chrom, cnarr.data["start"].to_numpy(), cnarr.data["end"].to_numpy())
        cnarr = cnarr.as_dataframe(cnarr.data.iloc[order])
        ref_matched = ref_matched.as_dataframe(ref_matched.data.iloc[order])
        kept_rows = kept_rows[order]            # the device columns follow the same re-ordering
        chrom = cnarr.data["chromosome"].to_numpy()
        names, offsets, chrom_id = genome.chrom_runs(chrom)
    start = cnarr.data["start"].to_numpy(dtype=np.int64)
    end = cnarr.data["end"].to_numpy(dtype=np.int64)
    auto = genome.autosome_row_mask(chrom, start, end, diploid_parx_genome)

    kidx = torch.from_numpy(np.ascontiguousarray(kept_rows)).to(d_log2.device)
    d_log2 = d_log2.index_select(0, kidx).contiguous()
    d_depth = _dev.up(cnarr.data["depth"].to_numpy(dtype=np.float64)) if "depth" in cnarr else None
    d_gc = dev_ref["gc"].index_select(0, kidx).contiguous() if (fix_gc and "gc" in dev_ref) else None
    d_rmask = dev_ref["rmask"].index_select(0, kidx).contiguous() if (fix_rmask and "rmask" in dev_ref) else None
    if fix_gc and d_gc is None:
        logging.warning("WARNING: Skipping correction for GC bias")
    if fix_rmask and d_rmask is None:
        logging.warning("WARNING: Skipping correction for RepeatMasker bias")
    frac = smoothing_window_fraction
    if frac is None:
        frac = max(0.01, n ** -0.5)
    wing = smoothing._width2wing(frac, n) if n >= 2 else 1
    skipped = ctypes.c_int32(0)
    d_off, d_cid = _dev.up(offsets), _dev.up(chrom_id)
    d_start, d_end, d_auto = _dev.up(start), _dev.up(end), _dev.up(auto.astype(np.uint8))
    if d_gc is not None:
        logging.info("Correcting for GC bias...")
    if fix_edge:
        logging.info("Correcting for density bias...")
    if d_rmask is not None:
        logging.info("Correcting for RepeatMasker bias...")
    _lib.call("cnvk_fix_group_dev", _dev.ptr(d_log2), _dev.ptr(d_depth), _dev.ptr(d_auto), n, _dev.ptr(d_off),
              len(names), _dev.ptr(d_cid), _dev.ptr(d_start), _dev.ptr(d_end), _dev.ptr(d_gc), _dev.ptr(d_rmask), 0,
              int(bool(fix_edge)), int(bool(skip_low)), 1, wing, _dev.ptr(_dev.shuffle_permutation(n)),
              params.INSERT_SIZE, b"ger", ctypes.addressof(skipped), _dev.stream())
    if skipped.value:
        logging.warning("WARNING: most bins have no or very low coverage; check that the right BED file was used")
    out = cnarr.data.assign(log2=d_log2.cpu().numpy()).reset_index(drop=True)
    ref_out = ref_matched.data.reset_index(drop=True)
    return cnarr.as_dataframe(out), ref_matched.as_dataframe(ref_out)


def match_ref_to_sample(ref_cnarr, samp_cnarr):
    """Reference rows re-ordered to the sample's bins by exact (chromosome, start, end)
    (cnvlib/fix.py:324-370), with the reference's error messages."""
    fast = _match_fast(ref_cnarr, samp_cnarr)
    if fast is not None:
        matched = ref_cnarr.data.iloc[fast].set_index(samp_cnarr.data.index)
        return ref_cnarr.as_dataframe(matched)
    samp_keys = pd.MultiIndex.from_arrays(
        [samp_cnarr.data["chromosome"].to_numpy(), samp_cnarr.data["start"].to_numpy(),
         samp_cnarr.data["end"].to_numpy()])
    ref_keys = pd.MultiIndex.from_arrays(
        [ref_cnarr.data["chromosome"].to_numpy(), ref_cnarr.data["start"].to_numpy(),
         ref_cnarr.data["end"].to_numpy()])
    for keys, name in ((samp_keys, "sample"), (ref_keys, "reference")):
        dupes = keys.duplicated()
        if dupes.any():
            raise ValueError(
                "Duplicated genomic coordinates in {} set. Total duplicated regions: {}, starting with:\n{}.\n"
                "These usually come from duplicate or overlapping rows in the bait BED. Rebuild the targets "
                "with 'cnvkit.py target', which merges overlapping baits into disjoint bins, then redo "
                "'coverage' and 'reference' with the rebuilt BED.".format(
                    name, int(dupes.sum()), "\n".join(map(str, keys[dupes][:10]))))
    pos = ref_keys.get_indexer(samp_keys)
    num_missing = int((pos < 0).sum())
    if num_missing > 0:
        if num_missing == len(samp_keys):
            raise ValueError(
                f"Reference shares no bins with {samp_cnarr.sample_id}. Is the reference argument actually a "
                "reference .cnn built with 'cnvkit.py reference' (and not a coverage file or the wrong panel)?")
        raise ValueError(
            f"Reference is missing {num_missing} of {len(samp_keys)} bins found in {samp_cnarr.sample_id}; the "
            "reference was likely built for a different target panel.")
    matched = ref_cnarr.data.iloc[pos].set_index(samp_cnarr.data.index)
    return ref_cnarr.as_dataframe(matched)


def _match_fast(ref_cnarr, samp_cnarr):
    """Positions of the sample's bins in the reference by exact (chromosome, start, end), or None when
    anything is unusual (duplicates, missing bins, coordinates outside 32 bits) -- the caller then takes
    the MultiIndex path, which also words the errors.  Chromosome names are factorised jointly and
    (code, start) packed into one int64 that is sorted once and searched."""
    rs, re_ = ref_cnarr.data["start"].to_numpy(), ref_cnarr.data["end"].to_numpy()
    ss, se = samp_cnarr.data["start"].to_numpy(), samp_cnarr.data["end"].to_numpy()
    nr, ns = len(rs), len(ss)
    if nr == 0 or ns == 0:
        return None
    if not (np.issubdtype(rs.dtype, np.integer) and np.issubdtype(ss.dtype, np.integer)):
        return None
    if min(rs.min(), ss.min()) < 0 or max(rs.max(), ss.max()) >= (1 << 32):
        return None
    both = np.concatenate([ref_cnarr.data["chromosome"].to_numpy().astype(object),
                           samp_cnarr.data["chromosome"].to_numpy().astype(object)])
    codes, uniq = pd.factorize(both, sort=False)
    if len(uniq) >= (1 << 30) or (codes < 0).any():
        return None
    kr = (codes[:nr].astype(np.int64) << 32) | rs.astype(np.int64)
    ks = (codes[nr:].astype(np.int64) << 32) | ss.astype(np.int64)
    order = np.lexsort((re_, kr))
    kr_s, re_s = kr[order], np.asarray(re_)[order]
    if nr > 1 and ((kr_s[1:] == kr_s[:-1]) & (re_s[1:] == re_s[:-1])).any():
        return None                                   # duplicated reference bins
    so = np.lexsort((se, ks))
    if ns > 1 and ((ks[so][1:] == ks[so][:-1]) & (np.asarray(se)[so][1:] == np.asarray(se)[so][:-1])).any():
        return None                                   # duplicated sample bins
    lo = np.searchsorted(kr_s, ks, "left")
    lo_c = np.minimum(lo, nr - 1)
    hit = (lo < nr) & (kr_s[lo_c] == ks) & (re_s[lo_c] == se)
    if not hit.all():
        # several reference bins may share (chromosome, start): look further along the run
        miss = np.flatnonzero(~hit)
        for m in miss:
            j = lo[m]
            while j < nr and kr_s[j] == ks[m] and re_s[j] != se[m]:
                j += 1
            if j < nr and kr_s[j] == ks[m] and re_s[j] == se[m]:
                lo_c[m] = j
            else:
                return None                           # missing bins
    return order[lo_c]


def _finish(cnarr, ref_matched, diploid_parx_genome):
    """fix.py:212-214 on device: ratio, apply_weights, center_all(skip_low=True)."""
    n = len(cnarr)
    if n == 0:
        return np.zeros(0), np.zeros(0)
    chrom = cnarr.data["chromosome"].to_numpy()
    names, offsets, chrom_id = genome.chrom_runs(chrom)
    start = cnarr.data["start"].to_numpy(dtype=np.int64)
    end = cnarr.data["end"].to_numpy(dtype=np.int64)
    auto = genome.autosome_row_mask(chrom, start, end, diploid_parx_genome)
    is_anti = cnarr.data["gene"].isin(params.ANTITARGET_ALIASES).to_numpy()
    d_log2 = _dev.up(cnarr.data["log2"].to_numpy(dtype=np.float64))
    d_depth = _dev.up(cnarr.data["depth"].to_numpy(dtype=np.float64)) if "depth" in cnarr else None
    d_rl = _dev.up(ref_matched.data["log2"].to_numpy(dtype=np.float64))
    d_rs = _dev.up(ref_matched.data["spread"].to_numpy(dtype=np.float64))
    d_w = _dev.empty(n, torch.float64)
    d_info = _dev.empty(8, torch.float64)
    # keep every staged tensor referenced until the call has been enqueued
    d_start, d_end, d_cid = _dev.up(start), _dev.up(end), _dev.up(chrom_id)
    d_anti, d_auto, d_off = _dev.up(is_anti.astype(np.uint8)), _dev.up(auto.astype(np.uint8)), _dev.up(offsets)
    _lib.call("cnvk_fix_finish_dev", _dev.ptr(d_log2), _dev.ptr(d_depth), _dev.ptr(d_rl), _dev.ptr(d_rs),
              _dev.ptr(d_start), _dev.ptr(d_end), _dev.ptr(d_cid), _dev.ptr(d_anti), _dev.ptr(d_auto), n,
              _dev.ptr(d_off), len(names), _dev.ptr(d_w), _dev.ptr(d_info), _dev.stream())
    info = d_info.cpu().numpy()
    if is_anti.any():
        n_anti, n_ok = info[5], info[6]
        frac_low = 1 - (n_ok / n_anti) if n_anti else 0.0
        if frac_low > 0.5:
            logging.warning(
                f"WARNING: Most antitarget bins ({100 * frac_low:.2f}%, {int(n_anti - n_ok):d}/{int(n_anti):d})"
                " have low or no coverage; is this amplicon/WGS?")
        tgt_var, anti_var = info[1] ** 2, info[3] ** 2
        var_ratio = max(tgt_var, 0.01) / max(anti_var, 0.01)
        if var_ratio > 1:
            logging.info("Targets are %.2f x more variable than antitargets", var_ratio)
        else:
            logging.info("Antitargets are %.2f x more variable than targets", 1.0 / var_ratio)
    return d_w.cpu().numpy(), d_log2.cpu().numpy()


# ---- inner seams, numpy in / numpy out -------------------------------
def center_by_window(log2, fraction, sort_key):
    """cnvlib/fix.py:373-427 on plain arrays: returns the bias-corrected log2 in the input order."""
    x = _lib.f64(np.asarray(log2, dtype=float)).copy()
    key = _lib.f64(np.asarray(sort_key, dtype=float))
    n = len(x)
    if n == 0:
        return x
    perm = np.random.default_rng(0xA5EED).permutation(n).astype(np.uint32)
    wing = smoothing._width2wing(fraction, n) if n >= 2 else 1
    _lib.call("cnvk_center_by_window", _lib.ptr(x), _lib.ptr(key), _lib.ptr(perm), n, wing)
    return x


def get_edge_bias(chromosomes, start, end, margin=params.INSERT_SIZE):
    """cnvlib/fix.py:430-459 on plain columns (rows genomically sorted)."""
    names, offsets, chrom_id = genome.chrom_runs(np.asarray(chromosomes, dtype=object))
    n = len(chrom_id)
    out = _dev.empty(n, torch.float64)
    d_start, d_end = _dev.up(np.asarray(start, dtype=np.int64)), _dev.up(np.asarray(end, dtype=np.int64))
    d_cid = _dev.up(chrom_id)
    _lib.call("cnvk_edge_bias_dev", _dev.ptr(d_start), _dev.ptr(d_end), _dev.ptr(d_cid), n, int(margin),
              _dev.ptr(out), _dev.stream())
    return out.cpu().numpy()


def center_all(log2, chromosomes, depth=None, start=None, end=None, skip_low=False, diploid_parx_genome=None):
    """CopyNumArray.center_all (cnvlib/cnary.py:253-313) on plain columns; returns shifted log2."""
    chrom = np.asarray(chromosomes, dtype=object)
    names, offsets, _ = genome.chrom_runs(chrom)
    auto = genome.autosome_row_mask(chrom, start, end, diploid_parx_genome)
    d_log2 = _dev.up(np.asarray(log2, dtype=np.float64))
    d_depth = _dev.up(np.asarray(depth, dtype=np.float64)) if depth is not None else None
    d_auto, d_off = _dev.up(auto.astype(np.uint8)), _dev.up(offsets)
    _lib.call("cnvk_center_all_dev", _dev.ptr(d_log2), _dev.ptr(d_depth), _dev.ptr(d_auto), len(chrom),
              _dev.ptr(d_off), len(names), int(bool(skip_low)), 0, _dev.stream())
    return d_log2.cpu().numpy()

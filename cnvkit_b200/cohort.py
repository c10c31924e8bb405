"""Cohort fix + segment on one GPU -- the numeric core of cnvlib/batch.py:524-555 (batch_run_sample:
do_fix followed by do_segmentation) for many samples that share one bin set and one pooled
reference.

Everything do_fix / do_segmentation derive from the *bins and the reference alone* -- reference
rows matched to the sample's bins (fix.py:324-370), mask_bad_bins (fix.py:295-321), chromosome
runs, autosome mask, edge-bias column, shuffle permutation, the target/antitarget merge order
(fix.py:171-174), arm boundaries (gary.py:309-355), gene codes for transfer_fields -- is computed once
per `Panel` and kept in HBM.  A sample then costs two host->device column copies, the device calls
of csrc/fix.cu and csrc/segutil.cu, and one share of a CBS / Haar call that segments the arms of
many samples together (the recursion is level-synchronous over the whole batch, so the small arms
of an exome panel fill the GPU only when samples are batched).

Results are identical to calling cnvkit_b200.fix.do_fix and
cnvkit_b200.segmentation.do_segmentation sample by sample (tests/test_cohort_gpu.py).  Samples the
shared plan does not cover (NaN log2 in a coverage table, non-finite ratios) take that general path.
"""
from __future__ import annotations

import ctypes
import logging

import numpy as np
import pandas as pd
import torch

from . import _dev, _lib, fix, genome, params, segmentation, segmetrics, smoothing, tabio
from .cnary import arm_ranges
from .segmentation import cbs as _cbs
from .segmentation import haar as _haar


class _Group:
    """One bin set (targets or antitargets) of load_adjust_coverages after reference masking."""

    def __init__(self, table, ref_matched, skip_low, do_gc, do_edge, do_rmask, diploid_parx_genome):
        self.n0 = len(table)
        self.skip_low = skip_low
        self.do_edge = do_edge
        rd = ref_matched.data
        self.table = table
        self.ref_matched = ref_matched
        if self.n0 == 0:
            self.n = 0
            self.keep = np.zeros(0, dtype=bool)
            return
        cols = {c: _dev.up(rd[c].to_numpy(dtype=np.float64)) for c in ("log2", "spread", "depth", "gc", "rmask")
                if c in rd.columns}
        d_keep = _dev.empty(self.n0, torch.uint8)
        # sample-independent part of the mask (the sample-NaN test is checked per sample)
        _lib.call("cnvk_fix_mask_dev", 0, _dev.ptr(cols["log2"]), _dev.ptr(cols["spread"]), _dev.ptr(cols.get("depth")),
                  _dev.ptr(cols.get("gc")), self.n0, _dev.ptr(d_keep), _dev.stream())
        self.keep = d_keep.cpu().numpy().astype(bool)
        self.d_keep_idx = torch.from_numpy(np.flatnonzero(self.keep)).to(_dev.device())
        kept = table.data[self.keep]
        self.kept_table = kept.reset_index(drop=True)
        self.kept_ref = rd[self.keep].reset_index(drop=True)
        self.n = n = len(kept)
        if n == 0:
            return
        chrom = kept["chromosome"].to_numpy()
        names, offsets, chrom_id = genome.chrom_runs(chrom)
        start = kept["start"].to_numpy(dtype=np.int64)
        end = kept["end"].to_numpy(dtype=np.int64)
        auto = genome.autosome_row_mask(chrom, start, end, diploid_parx_genome)
        self.nchrom = len(names)
        self.d_off, self.d_cid = _dev.up(offsets), _dev.up(chrom_id)
        self.d_start, self.d_end = _dev.up(start), _dev.up(end)
        self.d_auto = _dev.up(auto.astype(np.uint8))
        self.d_gc = cols["gc"].index_select(0, self.d_keep_idx).contiguous() if (do_gc and "gc" in cols) else None
        self.d_rmask = (cols["rmask"].index_select(0, self.d_keep_idx).contiguous()
                        if (do_rmask and "rmask" in cols) else None)
        self.wing = smoothing._width2wing(max(0.01, n ** -0.5), n) if n >= 2 else 1
        self.d_perm = _dev.shuffle_permutation(n)
        self.d_edge = None
        if do_edge:
            self.d_edge = _dev.empty(n, torch.float64)
            _lib.call("cnvk_edge_bias_dev", _dev.ptr(self.d_start), _dev.ptr(self.d_end), _dev.ptr(self.d_cid), n,
                      params.INSERT_SIZE, _dev.ptr(self.d_edge), _dev.stream())
        # visiting orders of center_by_window for GC, edge, RepeatMasker ("ger", fix.py:265-289): they
        # depend on the bins only, so the per-sample work is one rolling median per trait, no sort
        self.d_orders = []
        for key in (self.d_gc, self.d_edge, self.d_rmask):
            if key is None:
                self.d_orders.append(None)
                continue
            d_ord = _dev.empty(n, torch.int32)
            _lib.call("cnvk_trait_order_dev", _dev.ptr(key), _dev.ptr(self.d_perm), n, _dev.ptr(d_ord), _dev.stream())
            self.d_orders.append(d_ord)

    def adjust(self, log2, depth):
        """Bias-corrected log2 of the kept rows (device tensor), or None if the sample needs the
        general path (NaN log2 would change the kept set, fix.py:244)."""
        if self.n0 == 0 or self.n == 0:
            return _dev.empty(0, torch.float64), _dev.empty(0, torch.float64)
        if torch.is_tensor(log2):                 # columns already resident in HBM
            if bool(torch.isnan(log2).any()):
                return None
            d_all, d_dall = log2, depth
        else:
            x = np.ascontiguousarray(log2, dtype=np.float64)
            if np.isnan(x).any():
                return None
            d_all, d_dall = _dev.up(x), _dev.up(np.ascontiguousarray(depth, dtype=np.float64))
        d_log2 = d_all.index_select(0, self.d_keep_idx).contiguous()
        d_depth = d_dall.index_select(0, self.d_keep_idx).contiguous()
        skipped = ctypes.c_int32(0)
        o = self.d_orders
        _lib.call("cnvk_fix_group_ordered_dev", _dev.ptr(d_log2), _dev.ptr(d_depth), _dev.ptr(self.d_auto), self.n,
                  _dev.ptr(self.d_off), self.nchrom, _dev.ptr(o[0]), _dev.ptr(o[1]), _dev.ptr(o[2]),
                  int(bool(self.skip_low)), 1, self.wing, ctypes.addressof(skipped), _dev.stream())
        if skipped.value:
            logging.warning("WARNING: most bins have no or very low coverage; check that the right BED file was used")
        return d_log2, d_depth


BATCH_CI = dict(alpha=0.5, bootstraps=100, smoothed=True)     # cnvlib/batch.py:559-566


def _ci_options(ci):
    """None / False -> no intervals; True -> the batch command's settings; dict -> do_segmetrics keywords."""
    if not ci:
        return None
    opts = dict(BATCH_CI)
    if isinstance(ci, dict):
        unknown = set(ci) - {"alpha", "bootstraps", "smoothed"}
        if unknown:
            raise ValueError(f"unknown confidence-interval options: {sorted(unknown)}")
        opts = dict(alpha=0.05, bootstraps=100, smoothed=10)   # do_segmetrics' own defaults
        opts.update(ci)
    return opts


class Panel:
    """Sample-independent plan for fix + segment over a shared bin set.

    Parameters mirror do_fix (cnvlib/fix.py:20-31): `target_bins` / `antitarget_bins` are any
    coverage tables of the panel (only their coordinates and gene names are used), `reference` the
    pooled reference.
    """

    def __init__(self, target_bins, antitarget_bins, reference, diploid_parx_genome=None, do_gc=True, do_edge=True,
                 do_rmask=True):
        self.diploid_parx_genome = diploid_parx_genome
        self.opts = (do_gc, do_edge, do_rmask)
        self.meta_cls = target_bins.__class__
        keepc = ("chromosome", "start", "end", "gene", "log2", "depth")
        tb = target_bins.keep_columns(keepc) if "gc" in target_bins else target_bins
        ab = antitarget_bins.keep_columns(keepc) if "gc" in antitarget_bins else antitarget_bins
        self._proto_t, self._proto_a = tb, ab
        self.t = _Group(tb, fix.match_ref_to_sample(reference, tb) if len(tb) else reference[:0], True, do_gc, do_edge,
                        False, diploid_parx_genome)
        self.a = _Group(ab, fix.match_ref_to_sample(reference, ab) if len(ab) else reference[:0], False, do_gc, False,
                        do_rmask, diploid_parx_genome)
        # fix.py:171-174: concat + genomic sort of the two kept tables (and of the matched reference rows)
        parts = [g.kept_table for g in (self.t, self.a) if g.n0 and g.n]
        rparts = [g.kept_ref for g in (self.t, self.a) if g.n0 and g.n]
        cdf = pd.concat(parts, ignore_index=True)
        rdf = pd.concat(rparts, ignore_index=True)
        if self.a.n0 and self.a.n:
            order = genome.genomic_order(cdf["chromosome"].to_numpy(), cdf["start"].to_numpy(), cdf["end"].to_numpy())
        else:
            order = np.arange(len(cdf))
        self.order = order
        self.d_order = torch.from_numpy(order.astype(np.int64)).to(_dev.device())
        self.bins = cdf.iloc[order].reset_index(drop=True)            # chromosome,start,end,gene (+ log2/depth unused)
        rdf = rdf.iloc[order].reset_index(drop=True)
        self.n = n = len(self.bins)
        chrom = self.bins["chromosome"].to_numpy()
        self.chrom = chrom
        names, offsets, chrom_id = genome.chrom_runs(chrom)
        self.start = self.bins["start"].to_numpy(dtype=np.int64)
        self.end = self.bins["end"].to_numpy(dtype=np.int64)
        auto = genome.autosome_row_mask(chrom, self.start, self.end, diploid_parx_genome)
        self.is_anti = self.bins["gene"].isin(params.ANTITARGET_ALIASES).to_numpy()
        self.nchrom = len(names)
        self.d_rl = _dev.up(rdf["log2"].to_numpy(dtype=np.float64))
        self.d_rs = _dev.up(rdf["spread"].to_numpy(dtype=np.float64))
        self.d_start, self.d_end, self.d_cid = _dev.up(self.start), _dev.up(self.end), _dev.up(chrom_id)
        self.d_anti, self.d_auto, self.d_off = (_dev.up(self.is_anti.astype(np.uint8)), _dev.up(auto.astype(np.uint8)),
                                                _dev.up(offsets))
        # segmentation: arms of the .cnr table, gene codes for transfer_fields
        self.arms = arm_ranges(self.bins)
        self.arm_off = np.array([a[1] for a in self.arms] + [self.arms[-1][2]], dtype=np.int64)
        self.arm_id = np.repeat(np.arange(len(self.arms)), np.diff(self.arm_off))
        self.d_arm_id = torch.from_numpy(self.arm_id).to(_dev.device())
        self.gene_codes = segmentation.factorize_genes(self.bins["gene"].to_numpy())
        self._bin_index = None           # segmetrics.BinIndex of the .cnr rows, built on first use
        self._pool = None                # worker threads of run(workers > 1), kept across calls
        import threading
        self._tls = threading.local()    # per-thread pinned staging block of the batch path
        self._pool_size = 0
        self._desc = None                # cnvk_panel descriptor of the native batch entry points, built on first use
        # transfer_fields' simple row ranges need ascending bin ends within every arm (intersect.py:323-346)
        step = np.diff(self.end)
        inner = np.ones(len(step), dtype=bool)
        inner[self.arm_off[1:-1] - 1] = False
        self.nested_bins = bool(np.any(step[inner] < 0))

    # ------------------------------------------------------------------ fix
    def fix(self, target, antitarget):
        """do_fix for one sample given its two coverage tables (CopyNumArray or dict with
        'log2'/'depth' columns in the panel's bin order).  Returns (d_log2, d_weight, d_depth) device
        columns over the panel's .cnr rows, or None when the sample needs the general path."""
        tl, td = _cols(target)
        al, ad = _cols(antitarget)
        if len(tl) != self.t.n0 or len(al) != self.a.n0:
            raise ValueError("sample tables do not match the panel's bin set")
        rt = self.t.adjust(tl, td)
        ra = self.a.adjust(al, ad)
        if rt is None or ra is None:
            return None
        parts = [r[0] for r in (rt, ra) if len(r[0])]
        dparts = [r[1] for r in (rt, ra) if len(r[1])]
        d_log2 = torch.cat(parts).index_select(0, self.d_order).contiguous()
        d_depth = torch.cat(dparts).index_select(0, self.d_order).contiguous()
        d_w = _dev.empty(self.n, torch.float64)
        _lib.call("cnvk_fix_finish_dev", _dev.ptr(d_log2), _dev.ptr(d_depth), _dev.ptr(self.d_rl), _dev.ptr(self.d_rs),
                  _dev.ptr(self.d_start), _dev.ptr(self.d_end), _dev.ptr(self.d_cid), _dev.ptr(self.d_anti),
                  _dev.ptr(self.d_auto), self.n, _dev.ptr(self.d_off), self.nchrom, _dev.ptr(d_w), 0, _dev.stream())
        return d_log2, d_w, d_depth

    def cnr_table(self, d_log2, d_w, d_depth, sample_id=None):
        """The .cnr CopyNumArray do_fix returns (host copy of the three device columns)."""
        df = self.bins[["chromosome", "start", "end", "gene"]].assign(
            log2=d_log2.cpu().numpy(), depth=d_depth.cpu().numpy(), weight=d_w.cpu().numpy())
        return self.meta_cls(df, {"sample_id": sample_id})

    # ------------------------------------------------------------------ segment
    def segment(self, samples, method="cbs", threshold=None, skip_outliers=10, sample_ids=None, ci=None):
        """do_segmentation(method) for a list of (d_log2, d_weight, d_depth) device triples over the
        panel's .cnr rows; all samples are segmented in one batched device call.  Returns one .cns
        CopyNumArray per sample.  With `ci` (True = the batch command's alpha 0.5 / smoothed bootstrap, or a
        dict of do_segmetrics keywords) every table also gets the `ci_lo` / `ci_hi` columns of
        do_segmetrics(cnr, cns, interval_stats=["ci"], ...) (cnvlib/batch.py:559-566)."""
        ci = _ci_options(ci)
        if method not in ("cbs", "haar"):
            raise ValueError("cohort segmentation implements 'cbs' and 'haar'")
        if threshold is None:
            threshold = 0.0001
        n_arms = len(self.arms)
        per = []
        xs, ws, offs = [], [], [0]
        unit_arm_all = []
        for d_log2, d_w, d_depth in samples:
            d_mask = None
            if method == "cbs":
                # outlier mask, weight gate, %.6g quantisation, the R script's row filters and the compaction in one
                # native call (one synchronisation; cnvk_cbs_input_filter_dev)
                if skip_outliers:
                    d_mask = _dev.empty(self.n, torch.uint8)
                    _lib.call("cnvk_outlier_mask_dev", _dev.ptr(d_log2), _lib.ptr(self.arm_off), n_arms, 0.95,
                              float(skip_outliers), _dev.ptr(d_mask), _dev.stream())
                xq, wq = _dev.empty(self.n, torch.float64), _dev.empty(self.n, torch.float64)
                d_src = _dev.empty(self.n, torch.int64)
                units = np.zeros(n_arms + 1, dtype=np.int64)
                flags = np.zeros(1, dtype=np.uint32)
                _lib.call("cnvk_cbs_input_filter_dev", _dev.ptr(d_log2), _dev.ptr(d_w), _dev.ptr(d_mask),
                          _dev.ptr(self.d_arm_id), _lib.ptr(self.arm_off), n_arms, self.n, _dev.ptr(xq), _dev.ptr(wq),
                          _dev.ptr(d_src), _lib.ptr(units), _lib.ptr(flags), _dev.stream())
                if flags[0] & 1:                  # non-finite ratios: the general path handles the per-arm drops
                    per.append(None)
                    continue
                if not flags[0] & 2:
                    nk = int(units[-1])
                    per.append(dict(fidx=None, d_src=d_src[:nk], units=units, unit_arm=np.arange(n_arms), base=offs[-1],
                                    unit0=len(unit_arm_all)))
                    unit_arm_all.extend(range(n_arms))
                    xs.append(xq[:nk])
                    ws.append(wq[:nk])
                    offs.extend((units[1:] + offs[-1]).tolist() if len(units) > 1 else [])
                    continue
                # a value outside the device quantiser's range (|v| < 1e-16 or > 1e21): the step-by-step path below
            if not bool(torch.isfinite(d_log2).all()):
                per.append(None)
                continue
            keep = torch.ones(self.n, dtype=torch.bool, device=d_log2.device)
            if skip_outliers:
                if d_mask is None:
                    d_mask = _dev.empty(self.n, torch.uint8)
                    _lib.call("cnvk_outlier_mask_dev", _dev.ptr(d_log2), _lib.ptr(self.arm_off), n_arms, 0.95,
                              float(skip_outliers), _dev.ptr(d_mask), _dev.stream())
                keep &= d_mask == 0
            keep &= d_w != 0                      # NaN weights compare True and survive, as in the reference
            d_fidx = torch.nonzero(keep).flatten()
            fidx = d_fidx.cpu().numpy()
            f_arm = self.arm_id[fidx]
            x = d_log2.index_select(0, d_fidx)
            w = d_w.index_select(0, d_fidx)
            if method == "cbs":
                x, w, sub = self._quantise(x, w, f_arm, n_arms)
                if sub is not None:
                    fidx, f_arm = fidx[sub], f_arm[sub]
                units = segmentation._arm_offsets(f_arm, n_arms)
                unit_arm = np.arange(n_arms)
            else:
                units, unit_arm = self._haar_units(fidx, f_arm, n_arms)
            per.append(dict(fidx=fidx, units=units, unit_arm=unit_arm, base=offs[-1], unit0=len(unit_arm_all)))
            unit_arm_all.extend(unit_arm.tolist())
            xs.append(x)
            ws.append(w)
            offs.extend((units[1:] + offs[-1]).tolist() if len(units) > 1 else [])
        results = [None] * len(samples)
        if xs:
            X = (xs[0] if len(xs) == 1 else torch.cat(xs)).contiguous()
            W = (ws[0] if len(ws) == 1 else torch.cat(ws)).contiguous()
            off = np.asarray(offs, dtype=np.int64)
            if method == "cbs":
                u, lo, hi, mean, _st = _cbs.segment_arms(None, None, off, alpha=float(threshold),
                                                         device_ptrs=(X.data_ptr(), W.data_ptr()), stream=_dev.stream())
                mean = np.array([float("%.15g" % m) for m in mean]) if len(mean) else mean
            else:
                u, lo, hi, mean = _haar.segment_arms(None, None, off, float(threshold),
                                                     device_ptrs=(X.data_ptr(), W.data_ptr()), stream=_dev.stream())
            for k, info in enumerate(per):
                if info is None:
                    continue
                nu = len(info["unit_arm"])
                sel = (u >= info["unit0"]) & (u < info["unit0"] + nu)
                s_unit = u[sel] - info["unit0"]
                s_lo, s_hi = lo[sel] - info["base"], hi[sel] - info["base"]
                results[k] = self._segments_table(samples[k], info, s_unit, s_lo, s_hi, mean[sel],
                                                  None if sample_ids is None else sample_ids[k], ci)
        for k, info in enumerate(per):
            if info is None:  # non-finite ratios: the general path handles the per-arm drops
                cnr = self.cnr_table(*samples[k], sample_id=None if sample_ids is None else sample_ids[k])
                results[k] = segmentation.do_segmentation(cnr, method, self.diploid_parx_genome, threshold=threshold,
                                                          skip_outliers=skip_outliers)
                if ci:
                    results[k] = segmetrics.do_segmetrics(cnr, results[k], interval_stats=["ci"], **ci)
        return results

    def _quantise(self, x, w, f_arm=None, n_arms=0):
        """%.6g of both columns (segmentation/__init__.py:299-301) and the R row filters (cbs.py:22-35)."""
        n = len(x)
        xq, wq = _dev.empty(n, torch.float64), _dev.empty(n, torch.float64)
        unhandled = np.zeros(1, dtype=np.uint64)
        cols = [(x.contiguous(), xq)]
        w_nan = torch.isnan(w)
        any_nan = bool(w_nan.any())
        if not any_nan or not bool(w_nan.all()):
            cols.append((w.contiguous(), wq))
            if any_nan and f_arm is not None:
                # the R script runs per arm: an arm without a single non-NA weight gets weight 1.0 (cbs.py:22-28)
                ok_np = (~w_nan).cpu().numpy()
                has_w = np.bincount(f_arm, weights=ok_np.astype(np.float64), minlength=n_arms) > 0
                no_w = ~has_w[f_arm]
                if no_w.any():
                    w = torch.where(torch.from_numpy(no_w).to(w.device), torch.ones_like(w), w)
                    cols[-1] = (w.contiguous(), wq)
        else:
            wq.fill_(1.0)                         # cbs.py:25-28: no usable weights -> weight 1 per bin
        for src, dst in cols:
            _lib.call("cnvk_round_sig_dev", _dev.ptr(src), n, 6, _dev.ptr(dst), _lib.ptr(unhandled), _dev.stream())
            if unhandled[0]:
                h = src.cpu().numpy()
                bad = (np.abs(h) < 1e-16) | (np.abs(h) > 1e21)
                fixed = dst.cpu().numpy()
                fixed[bad] = [float("%.6g" % v) for v in h[bad]]
                dst.copy_(torch.from_numpy(fixed))
        ok = ~torch.isnan(wq) & (wq > 0) & ~torch.isnan(xq)
        if bool(ok.all()):
            return xq, wq, None
        d_sub = torch.nonzero(ok).flatten()
        return xq.index_select(0, d_sub), wq.index_select(0, d_sub), d_sub.cpu().numpy()

    def _haar_units(self, fidx, f_arm, n_arms):
        """segment_haar re-splits every filtered arm with by_arm() again (haar.py:80-82)."""
        off = segmentation._arm_offsets(f_arm, n_arms)
        units, unit_arm = [0], []
        if len(fidx) == self.n:        # nothing filtered: the re-split is a property of the panel
            if not hasattr(self, "_haar_full"):
                self._haar_full = self._haar_units_general(fidx, off, n_arms)
            return self._haar_full
        return self._haar_units_general(fidx, off, n_arms)

    def _haar_units_general(self, fidx, off, n_arms):
        units, unit_arm = [0], []
        for a in range(n_arms):
            lo, hi = int(off[a]), int(off[a + 1])
            if hi <= lo:
                continue
            rows = fidx[lo:hi]
            sub = pd.DataFrame({"chromosome": self.chrom[rows], "start": self.start[rows], "end": self.end[rows]})
            for _c, _slo, shi in arm_ranges(sub):
                units.append(lo + shi)
                unit_arm.append(a)
        return np.asarray(units, dtype=np.int64), np.asarray(unit_arm, dtype=np.int64)

    def _segments_table(self, sample, info, s_unit, s_lo, s_hi, s_mean, sample_id, ci=None):
        d_log2, d_w, d_depth = sample
        seg_arm = info["unit_arm"][s_unit]
        if info["fidx"] is None:
            # source rows of the segment boundaries only: one small gather instead of the whole index column
            want = torch.from_numpy(np.concatenate([s_lo, s_hi - 1]).astype(np.int64)).to(d_log2.device)
            rows = info["d_src"].index_select(0, want).cpu().numpy()
            r_lo, r_hi = rows[:len(s_lo)], rows[len(s_lo):]
        else:
            src = info["fidx"]
            r_lo, r_hi = src[s_lo], src[s_hi - 1]
        seg_chrom = self.chrom[r_lo]
        seg_start = self.start[r_lo].copy()
        seg_end = self.end[r_hi].copy()
        probes = (s_hi - s_lo).astype(np.int64)
        segtab = segmentation._transfer_fields(seg_arm, seg_chrom, seg_start, seg_end, s_mean, probes, self.arms,
                                               self.bins, None, d_w, d_depth, self.start, self.end,
                                               gene_codes=self.gene_codes)
        out = self.meta_cls(segtab, {"sample_id": sample_id})
        # rows are produced arm by arm in the panel's genomic order and _transfer_fields lays the columns out in
        # sort_columns() order, so the reference's sort() / sort_columns() (:189-190) are no-ops here
        if list(segtab.columns[:5]) != ["chromosome", "start", "end", "gene", "log2"]:
            out.sort_columns()
        if ci:
            out = self._add_ci(out, sample, ci)
        return out

    # ------------------------------------------------------------------ native batch path
    def _descriptor(self):
        """cnvk_panel (include/cnvkit_b200.h) over this panel's device-resident columns."""
        if self._desc is None:
            if int(_lib.load().cnvk_panel_sizeof()) != ctypes.sizeof(_lib.PanelDesc):
                raise _lib.CnvkError("cnvk_panel layout differs between _lib.PanelDesc and the built library")
            D = _lib.PanelDesc()
            for k, g in enumerate((self.t, self.a)):
                G = D.grp[k]
                if not (g.n0 and g.n):
                    G.n0 = G.n = 0
                    continue
                G.n0, G.n = g.n0, g.n
                G.d_keep_idx, G.d_auto, G.d_chrom_off = g.d_keep_idx.data_ptr(), g.d_auto.data_ptr(), g.d_off.data_ptr()
                G.nchrom, G.skip_low, G.wing = g.nchrom, int(bool(g.skip_low)), int(g.wing)
                for j, o in enumerate(g.d_orders):
                    G.d_order[j] = o.data_ptr() if o is not None else None
            D.n = self.n
            D.d_order = self.d_order.data_ptr()
            D.d_ref_log2, D.d_ref_spread = self.d_rl.data_ptr(), self.d_rs.data_ptr()
            D.d_start, D.d_end, D.d_chrom_id = self.d_start.data_ptr(), self.d_end.data_ptr(), self.d_cid.data_ptr()
            D.d_is_anti, D.d_auto, D.d_chrom_off = self.d_anti.data_ptr(), self.d_auto.data_ptr(), self.d_off.data_ptr()
            D.nchrom, D.n_arms = self.nchrom, len(self.arms)
            D.d_arm_id = self.d_arm_id.data_ptr()
            D.arm_off = self.arm_off.ctypes.data
            self._desc = D
        return self._desc

    def _native_ok(self, method):
        return (method == "cbs" and not self.nested_bins and self.n > 0 and self.d_cid.dtype == torch.int32
                and all(o is None or o.dtype == torch.int32 for g in (self.t, self.a) if g.n0 and g.n for o in g.d_orders))

    def _prepare_native(self, work, want_cbs, skip_outliers=10):
        """cnvk_panel_prepare_dev for a batch: returns the device blocks (.cnr columns [ns, 3, n] = log2, depth,
        weight; CBS inputs [ns, 2, n]; source rows [ns, n]), their pointer tables, the survivors' arm offsets and the
        per-sample status bits."""
        ns, n, n_arms = len(work), self.n, len(self.arms)
        D = self._descriptor()
        hold, in_ptrs = [], np.zeros(4 * ns, dtype=np.uint64)
        dev = _dev.device()
        all_cols = []
        for tgt, anti in work:
            cols = _cols(tgt) + _cols(anti)
            if len(cols[0]) != self.t.n0 or len(cols[2]) != self.a.n0:
                raise ValueError("sample tables do not match the panel's bin set")
            all_cols.append(cols)
        if all(not torch.is_tensor(c) for cols in all_cols for c in cols):
            # host columns: packed into this thread's pinned staging block and moved with ONE asynchronous copy per
            # batch (four pageable copies per sample cost ~0.3 ms of the ~4 ms a sample takes at batch 1); the block is
            # free again when the prepare call below returns (it synchronises)
            per = 2 * (self.t.n0 + self.a.n0)
            stage = getattr(self._tls, "stage", None)
            if stage is None or stage.numel() < ns * per:
                stage = self._tls.stage = torch.empty(max(ns * per, 1), dtype=torch.float64).pin_memory()
            host = stage.numpy()
            pos, offs = 0, []
            for cols in all_cols:
                for c in cols:
                    host[pos: pos + len(c)] = c
                    offs.append(pos)
                    pos += len(c)
            d_in = stage[:pos].to(dev, non_blocking=True)
            hold.append(d_in)
            in_ptrs[:] = np.uint64(d_in.data_ptr()) + np.uint64(8) * np.asarray(offs, dtype=np.uint64)
        else:
            for k, cols in enumerate(all_cols):
                for j, c in enumerate(cols):
                    t = c if torch.is_tensor(c) else _dev.up(c)
                    if t.dtype != torch.float64 or not t.is_contiguous():
                        t = t.to(torch.float64).contiguous()
                    hold.append(t)
                    in_ptrs[4 * k + j] = t.data_ptr()
        d_out = torch.empty((ns, 3, n), dtype=torch.float64, device=dev)
        step = np.uint64(8 * n)
        out_ptrs = np.uint64(d_out.data_ptr()) + step * np.arange(3 * ns, dtype=np.uint64)
        status = np.zeros(ns, dtype=np.uint32)
        d_cbs = d_src = cbs_ptrs = src_ptrs = units = None
        if want_cbs:
            d_cbs = torch.empty((ns, 2, n), dtype=torch.float64, device=dev)
            d_src = torch.empty((ns, n), dtype=torch.int64, device=dev)
            cbs_ptrs = np.uint64(d_cbs.data_ptr()) + step * np.arange(2 * ns, dtype=np.uint64)
            src_ptrs = np.uint64(d_src.data_ptr()) + step * np.arange(ns, dtype=np.uint64)
            units = np.zeros((ns, n_arms + 1), dtype=np.int64)
        _lib.call("cnvk_panel_prepare_dev", ctypes.addressof(D), ns, _lib.ptr(in_ptrs), _lib.ptr(out_ptrs),
                  int(bool(want_cbs)), float(skip_outliers or 0.0), _lib.ptr(cbs_ptrs), _lib.ptr(src_ptrs), _lib.ptr(units),
                  _lib.ptr(status), _dev.stream())
        del hold
        return d_out, d_cbs, d_src, out_ptrs, src_ptrs, units, status

    def _chunk_native(self, work, sids, threshold, keep_cnr, ci, skip_outliers=10):
        """fix + CBS + transfer_fields for a batch of samples through the batch entry points: one synchronisation for
        the fix / filter stage, the CBS level loop, one for the transfer stage; the per-sample Python work is the
        gene join and the DataFrame.  Returns one (.cnr or None, .cns) per sample, None where the sample must take
        the step-by-step path (NaN input, the low-coverage guard, values the device quantiser declines)."""
        ns, n, n_arms = len(work), self.n, len(self.arms)
        D = self._descriptor()
        d_out, d_cbs, d_src, out_ptrs, src_ptrs, units, status = self._prepare_native(work, True, skip_outliers)
        good = [k for k in range(ns) if status[k] == 0]
        results = [None] * ns
        if not good:
            return results
        # ---- CBS over the survivors of every good sample (sample-major units)
        nk = units[:, -1]
        if len(good) == 1:
            k = good[0]
            X, W = d_cbs[k, 0, : nk[k]], d_cbs[k, 1, : nk[k]]
        else:
            X = torch.cat([d_cbs[k, 0, : nk[k]] for k in good])
            W = torch.cat([d_cbs[k, 1, : nk[k]] for k in good])
        bases = np.concatenate([[0], np.cumsum(nk[good])]).astype(np.int64)
        off = np.concatenate([units[k, :-1] + bases[i] for i, k in enumerate(good)] + [bases[-1:]]).astype(np.int64)
        u, lo, hi, mean, _st = _cbs.segment_arms(None, None, off, alpha=float(0.0001 if threshold is None else threshold),
                                                 device_ptrs=(X.data_ptr(), W.data_ptr()), stream=_dev.stream())
        nseg = len(u)
        if nseg:
            mean = np.array([float("%.15g" % m) for m in mean])
        u = u.astype(np.int64)
        smp = u // n_arms                     # position in `good`
        seg = np.empty((4, nseg), dtype=np.int64)
        seg[0] = np.asarray(good, dtype=np.int64)[smp]
        seg[1] = u - smp * n_arms
        seg[2] = lo - bases[smp]
        seg[3] = hi - bases[smp]
        bounds = np.empty((6, nseg), dtype=np.int64)
        sums = np.empty((2, nseg), dtype=np.float64)
        dep_ptrs = np.ascontiguousarray(out_ptrs[1::3])
        w_ptrs = np.ascontiguousarray(out_ptrs[2::3])
        _lib.call("cnvk_panel_transfer_dev", ctypes.addressof(D), ns, nseg, _lib.ptr(seg), _lib.ptr(src_ptrs),
                  _lib.ptr(dep_ptrs), _lib.ptr(w_ptrs), _lib.ptr(bounds), _lib.ptr(sums), _dev.stream())
        cut = np.searchsorted(smp, np.arange(len(good) + 1))
        for i, k in enumerate(good):
            a, b = int(cut[i]), int(cut[i + 1])
            genes = segmentation._join_genes(None, bounds[4, a:b], bounds[5, a:b], self.gene_codes)
            segtab = pd.DataFrame({
                "chromosome": self.chrom[bounds[0, a:b]], "start": bounds[2, a:b], "end": bounds[3, a:b], "gene": genes,
                "log2": mean[a:b], "depth": sums[1, a:b], "probes": seg[3, a:b] - seg[2, a:b], "weight": sums[0, a:b],
            }, copy=False)     # the columns are fresh per batch and the row ranges of its tables are disjoint
            cns = self.meta_cls(segtab, {"sample_id": sids[k]})
            trip = (d_out[k, 0], d_out[k, 2], d_out[k, 1])               # (log2, weight, depth) as Panel.fix returns
            if ci:
                cns = self._add_ci(cns, trip, ci)
            results[k] = (self.cnr_table(*trip, sample_id=sids[k]) if keep_cnr else None, cns)
        return results

    def _add_ci(self, out, trip, ci):
        """segmetrics.py:93: the bins each (stretched) segment overlaps, filtered bins included."""
        d_log2, d_w, _d_depth = trip
        if self._bin_index is None:
            self._bin_index = segmetrics.BinIndex(self.bins)
        b_lo, b_hi = self._bin_index.ranges(out.data)
        c_lo, c_hi = segmetrics.bootstrap_ci(None, None, b_lo, b_hi, device_ptrs=(d_log2.data_ptr(), d_w.data_ptr()),
                                             stream=_dev.stream(), **ci)
        out["ci_lo"] = c_lo.cpu().numpy()
        out["ci_hi"] = c_hi.cpu().numpy()
        return out

    # ------------------------------------------------------------------ driver
    def run(self, samples, method="cbs", threshold=None, batch=16, sample_ids=None, keep_cnr=False, workers=1,
            ci=None):
        """fix + segment for an iterable of (target, antitarget) coverage tables.  Returns a list of
        (.cnr or None, .cns) per sample, in input order; `batch` consecutive samples share one
        segmentation call.  With `workers` > 1 the batches are processed by that many host threads,
        each on its own CUDA stream, so that the host side of one batch (table assembly, gene names)
        overlaps the device side of another.  `ci`: see `segment`."""
        ci = _ci_options(ci)
        samples = list(samples)
        n = len(samples)
        ids = [sample_ids[k] if sample_ids is not None else getattr(samples[k][0], "sample_id", None)
               for k in range(n)]
        chunks = _chunk_plan(n, max(int(batch), 1), int(workers))

        native = self._native_ok(method)

        def do_chunk(idx):
            out, pend, pend_k = {}, [], []
            if native:
                fits = [k for k in idx if not (hasattr(samples[k][0], "data") and (
                    len(samples[k][0]) != self.t.n0 or len(samples[k][1]) != self.a.n0))]
                if fits:
                    res = self._chunk_native([samples[k] for k in fits], [ids[k] for k in fits], threshold, keep_cnr, ci)
                    for k, r in zip(fits, res):
                        if r is not None:
                            out[k] = r
                idx = [k for k in idx if k not in out]
            for k in idx:
                tgt, anti = samples[k]
                if hasattr(tgt, "data") and (len(tgt) != self.t.n0 or len(anti) != self.a.n0):
                    trip = None   # a table whose rows differ from the panel's (e.g. rows dropped on reading)
                else:
                    trip = self.fix(tgt, anti)
                if trip is None:  # general path for this sample
                    t_tab, a_tab = self._as_tables(tgt, anti, ids[k])
                    cnr = fix.do_fix(t_tab, a_tab, self._reference_table(), self.diploid_parx_genome, *self.opts)
                    cns = segmentation.do_segmentation(cnr, method, self.diploid_parx_genome, threshold=threshold)
                    if ci:
                        cns = segmetrics.do_segmetrics(cnr, cns, interval_stats=["ci"], **ci)
                    out[k] = (cnr if keep_cnr else None, cns)
                    continue
                pend.append(trip)
                pend_k.append(k)
            if pend:
                cns = self.segment(pend, method, threshold, sample_ids=[ids[k] for k in pend_k], ci=ci)
                for trip, k, seg in zip(pend, pend_k, cns):
                    out[k] = (self.cnr_table(*trip, sample_id=ids[k]) if keep_cnr else None, seg)
            return out

        results = {}
        if workers <= 1 or len(chunks) <= 1:
            for idx in chunks:
                results.update(do_chunk(idx))
        else:
            # Worker threads (one CUDA stream each) live as long as the panel.  The library keeps per-thread state --
            # a pinned staging arena, a device scratch slab, occupancy queries -- and torch's caching allocator pools
            # blocks per stream; with fresh threads and streams in every call, the first chunk of each worker took
            # 55-155 ms instead of ~40 (scripts/cohort_step_probe.py).
            pool, tls = self._executor(min(int(workers), len(chunks)))

            def task(idx):
                stream = tls.stream
                with torch.cuda.stream(stream):
                    part = do_chunk(idx)
                    stream.synchronize()
                return part

            torch.cuda.current_stream().synchronize()
            futures = [pool.submit(task, idx) for idx in chunks]
            first_error = None
            for f in futures:
                try:
                    results.update(f.result())
                except Exception as exc:  # surface in the caller, after every worker has stopped
                    first_error = first_error or exc
            if first_error is not None:
                raise first_error
        return [results[k] for k in range(n)]

    def _executor(self, nthreads):
        """Thread pool of run(workers > 1): `nthreads` persistent workers, each bound to the device with its own stream."""
        from concurrent.futures import ThreadPoolExecutor
        import threading

        if self._pool is None or self._pool_size < nthreads:
            if self._pool is not None:
                self._pool.shutdown(wait=True)
            dev = _dev.device()
            tls = threading.local()

            def init():
                torch.cuda.set_device(dev)
                tls.stream = torch.cuda.Stream(device=dev)

            self._pool = ThreadPoolExecutor(max_workers=nthreads, initializer=init, thread_name_prefix="cnvk-panel")
            self._pool_tls, self._pool_size = tls, nthreads
        return self._pool, self._pool_tls

    def close(self):
        """Stop the worker threads of run(workers > 1)."""
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None

    def __del__(self):
        try:
            if self._pool is not None:
                self._pool.shutdown(wait=False)
        except Exception:
            pass

    def _as_tables(self, tgt, anti, sid):
        def one(proto, s):
            if hasattr(s, "data"):
                return s
            return proto.as_dataframe(proto.data.assign(log2=np.asarray(s["log2"], dtype=float),
                                                        depth=np.asarray(s["depth"], dtype=float)))
        return one(self._proto_t, tgt), one(self._proto_a, anti)

    def _reference_table(self):
        rt = pd.concat([g.ref_matched.data for g in (self.t, self.a) if g.n0], ignore_index=True)
        ref = self.t.ref_matched.as_dataframe(rt)
        ref.sort()
        return ref


def run_files(target_fnames, antitarget_fnames, reference, out_dir, method="cbs", threshold=None, batch=8, workers=4,
              io_threads=8, diploid_parx_genome=None, do_gc=True, do_edge=True, do_rmask=True, write_cnr=True, ci=None):
    """File-to-file cohort run: the fix -> segment core of `cnvkit.py batch` (cnvlib/batch.py:524-555) for
    samples that share one panel.  Reads every sample's `.targetcoverage.cnn` / `.antitargetcoverage.cnn`
    with the native columns-only reader (bin sets verified by hash against the first sample), runs the
    shared-panel plan on the GPU, and writes `<sample>.cnr` / `<sample>.cns` into `out_dir` with the native
    `%.6g` writer.  Files whose bins differ from the first sample's (or that hold NaN coverage) go through
    the general per-sample path.  `ci=True` adds the batch command's bootstrap `ci_lo` / `ci_hi` columns
    (batch.py:559-566) to every `.cns`.  Returns the list of (cnr_path or None, cns_path)."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    from .cnary import CopyNumArray
    from .reference import fbase

    if len(target_fnames) != len(antitarget_fnames):
        raise ValueError("Unequal number of target and antitarget files given")
    if isinstance(reference, (str, os.PathLike)):
        reference = tabio.read(reference)
    os.makedirs(out_dir, exist_ok=True)
    n = len(target_fnames)
    if n == 0:
        return []
    t0 = tabio.read(target_fnames[0], sample_id=fbase(target_fnames[0]))
    a0 = tabio.read(antitarget_fnames[0], sample_id=fbase(antitarget_fnames[0]))
    panel = Panel(t0, a0, reference, diploid_parx_genome, do_gc, do_edge, do_rmask)
    _, ht = tabio.read_columns(target_fnames[0])
    _, ha = tabio.read_columns(antitarget_fnames[0]) if len(a0) else (None, None)
    # the columns-only reader returns rows in file order: usable when that is the table's sorted order
    sorted_t = _file_is_sorted(target_fnames[0], t0)
    sorted_a = _file_is_sorted(antitarget_fnames[0], a0) if len(a0) else True

    def load(k):
        tf, af = target_fnames[k], antitarget_fnames[k]
        try:
            if not (sorted_t and sorted_a):
                raise KeyError("unsorted panel")
            tc, h1 = tabio.read_columns(tf)
            if len(a0):
                ac, h2 = tabio.read_columns(af)
            else:
                ac, h2 = {"log2": np.zeros(0), "depth": np.zeros(0)}, None
            if h1 != ht or h2 != ha or len(tc["log2"]) != len(t0) or len(ac["log2"]) != len(a0):
                raise KeyError("different bins")
            return tc, ac
        except (KeyError, _lib.CnvkError):
            return (tabio.read(tf, sample_id=fbase(tf)), tabio.read(af, sample_id=fbase(af)))

    ids = [fbase(f) for f in target_fnames]
    outputs = [None] * n
    with ThreadPoolExecutor(max_workers=max(1, io_threads)) as pool:
        loaded = list(pool.map(load, range(n)))
        results = panel.run(loaded, method=method, threshold=threshold, batch=batch, sample_ids=ids, keep_cnr=write_cnr,
                            workers=workers, ci=ci)

        def store(k):
            cnr, cns = results[k]
            cnr_path = None
            if write_cnr and cnr is not None:
                cnr_path = os.path.join(out_dir, ids[k] + ".cnr")
                tabio.write(cnr, cnr_path)
            cns_path = os.path.join(out_dir, ids[k] + ".cns")
            tabio.write(cns, cns_path)
            return cnr_path, cns_path

        outputs = list(pool.map(store, range(n)))
    return outputs


def _chunk_plan(n, batch, workers):
    """Consecutive index ranges of `batch` samples.  (Tapering the sizes towards the end so that the workers finish
    together was measured and is slower: 646 vs 664 samples/s on a 128-sample share, 688 vs 779 on 1 024 -- small
    chunks lose more in per-call cost than the tail costs.)"""
    return [list(range(i, min(i + batch, n))) for i in range(0, n, batch)]


def _file_is_sorted(path, table):
    """True when the rows of `path` are already in the order tabio.read gives the table (genomic order,
    no dropped rows), so that columns read in file order line up with the panel's rows."""
    try:
        cols, _ = tabio.read_columns(path, columns=("start", "end"))
    except (KeyError, _lib.CnvkError):
        return False
    return (len(cols["start"]) == len(table)
            and np.array_equal(cols["start"], table.data["start"].to_numpy(dtype=np.float64))
            and np.array_equal(cols["end"], table.data["end"].to_numpy(dtype=np.float64)))


def _cols(s):
    if hasattr(s, "data"):
        d = s.data
        return d["log2"].to_numpy(dtype=np.float64), d["depth"].to_numpy(dtype=np.float64)
    if torch.is_tensor(s["log2"]):
        return s["log2"], s["depth"]
    return np.asarray(s["log2"], dtype=np.float64), np.asarray(s["depth"], dtype=np.float64)

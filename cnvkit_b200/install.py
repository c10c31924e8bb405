"""Drop the GPU hot path in beneath an imported CNVkit (``cnvlib``).

    import cnvlib, cnvkit_b200.install
    cnvkit_b200.install.install()        # cnvkit.py fix / reference / segment now run on the B200
    ...
    cnvkit_b200.install.uninstall()

The reference calls its seams through module attributes (``fix.do_fix`` at
cnvlib/commands.py:1177 and cnvlib/batch.py:524, ``reference.do_reference`` at commands.py:1017 and
batch.py:411, ``segmentation.do_segmentation`` at commands.py:1290 and batch.py:547,
``smoothing.rolling_median`` at fix.py:417, ``segmetrics.do_segmetrics`` at batch.py:559 and
commands.py's segmetrics command), so rebinding those attributes -- plus the public
aliases commands.py:979/1128/1277 creates at import time -- is the whole integration.
Options outside the GPU path (``method='hmm*'/'none'``, ``bias_smoother='loess'``, ``do_cluster``, ``variants`` with
CBS) are forwarded to the original callables by explicit predicates on the arguments; nothing else falls
back -- an in-scope call that fails raises.
"""
from __future__ import annotations

import importlib
import sys

_saved = {}


def _resolve(name):
    mod = sys.modules.get(name)
    if mod is None:
        mod = importlib.import_module(name)
    return mod


def _route(ours, theirs, needs_reference):
    """Call ``ours`` unless the arguments ask for something only the reference implements."""

    def seam(*args, **kwargs):
        if needs_reference(args, kwargs):       # an OPTION the GPU path does not cover (explicit predicates below)
            return theirs(*args, **kwargs)
        return ours(*args, **kwargs)            # in-scope inputs never fall back: errors propagate

    seam.__name__ = getattr(theirs, "__name__", "seam")
    seam.__doc__ = getattr(ours, "__doc__", None)
    seam.__wrapped__ = theirs
    return seam


def _arg(args, kwargs, pos, name, default):
    if name in kwargs:
        return kwargs[name]
    return args[pos] if len(args) > pos else default


def _fix_needs_reference(args, kwargs):
    # do_fix(target_raw, antitarget_raw, reference, diploid_parx_genome, do_gc, do_edge, do_rmask,
    #        do_cluster, smoothing_window_fraction, bias_smoother)   cnvlib/fix.py:20-31
    return bool(_arg(args, kwargs, 7, "do_cluster", False)) or _arg(args, kwargs, 9, "bias_smoother", "median") == "loess"


def _reference_needs_reference(args, kwargs):
    # do_reference(target_fnames, antitarget_fnames, fa_fname, is_haploid_x_reference, diploid_parx_genome,
    #              female_samples, do_gc, do_edge, do_rmask, do_cluster, min_cluster_size, cluster_method,
    #              bias_smoother)                                   cnvlib/reference.py:88-102
    return bool(_arg(args, kwargs, 9, "do_cluster", False)) or _arg(args, kwargs, 12, "bias_smoother", "median") == "loess"


def _segment_needs_reference(args, kwargs):
    # do_segmentation(cnarr, method, diploid_parx_genome, threshold, variants, skip_low, skip_outliers,
    #                 min_weight, save_dataframe, rscript_path, processes, smooth_cbs)
    #                                                               cnvlib/segmentation/__init__.py:31-44
    method = _arg(args, kwargs, 1, "method", None)
    variants = _arg(args, kwargs, 4, "variants", None)
    has_variants = variants is not None and len(variants) > 0
    return method not in ("cbs", "haar") or (has_variants and method != "haar")   # haar takes the BAF union itself


def _segmetrics_seam(ours, theirs):
    """do_segmetrics(cnarr, segarr, location_stats, spread_stats, interval_stats, alpha, bootstraps,
    smoothed, skip_low)  cnvlib/segmetrics.py:24-34: the bootstrap 'ci' columns come from the GPU, every
    other statistic from the reference (called without 'ci'), in the reference's column order."""

    def seam(*args, **kwargs):
        interval = list(_arg(args, kwargs, 4, "interval_stats", ()))
        if "ci" not in interval:
            return theirs(*args, **kwargs)
        names = ("cnarr", "segarr", "location_stats", "spread_stats", "interval_stats", "alpha", "bootstraps",
                 "smoothed", "skip_low")
        kw = dict(zip(names, args))
        kw.update(kwargs)
        rest = [s for s in interval if s != "ci"]
        ci = ours(kw["cnarr"], kw["segarr"], interval_stats=["ci"], alpha=kw.get("alpha", 0.05),
                  bootstraps=kw.get("bootstraps", 100), smoothed=kw.get("smoothed", 10),
                  skip_low=kw.get("skip_low", False))
        if not (kw.get("location_stats") or kw.get("spread_stats") or rest):
            return ci
        out = theirs(**dict(kw, interval_stats=rest))
        # the reference writes ci_lo/ci_hi before pi_lo/pi_hi (segmetrics.py:111-119)
        pi = {c: out.data.pop(c) for c in ("pi_lo", "pi_hi") if c in out.data.columns}
        out["ci_lo"] = ci["ci_lo"].to_numpy()
        out["ci_hi"] = ci["ci_hi"].to_numpy()
        for c, v in pi.items():
            out[c] = v.to_numpy()
        return out

    seam.__name__ = getattr(theirs, "__name__", "seam")
    seam.__doc__ = getattr(ours, "__doc__", None)
    seam.__wrapped__ = theirs
    return seam


def install():
    """Rebind cnvlib's hot-path seams to the B200 implementations.  Idempotent."""
    if _saved:
        return
    from . import _lib, fix, reference, segmentation, segmetrics, smoothing

    _lib.load()  # fail now, loudly, if the CUDA library is not built
    r_fix = _resolve("cnvlib.fix")
    r_ref = _resolve("cnvlib.reference")
    r_seg = _resolve("cnvlib.segmentation")
    r_smooth = _resolve("cnvlib.smoothing")
    r_metrics = _resolve("cnvlib.segmetrics")
    plan = [
        (r_fix, "do_fix", _route(fix.do_fix, r_fix.do_fix, _fix_needs_reference)),
        (r_ref, "do_reference", _route(reference.do_reference, r_ref.do_reference, _reference_needs_reference)),
        (r_ref, "do_reference_flat", _route(reference.do_reference_flat, r_ref.do_reference_flat, lambda a, k: False)),
        (r_seg, "do_segmentation", _route(segmentation.do_segmentation, r_seg.do_segmentation,
                                          _segment_needs_reference)),
        (r_smooth, "rolling_median", smoothing.rolling_median),
        (r_metrics, "do_segmetrics", _segmetrics_seam(segmetrics.do_segmetrics, r_metrics.do_segmetrics)),
    ]
    cmds = sys.modules.get("cnvlib.commands")
    for mod, name, new in plan:
        _saved[(mod.__name__, name)] = (mod, getattr(mod, name))
        setattr(mod, name, new)
        if cmds is not None and name.startswith("do_") and hasattr(cmds, name):
            _saved[(cmds.__name__, name)] = (cmds, getattr(cmds, name))
            setattr(cmds, name, new)


def uninstall():
    """Restore the reference's own callables."""
    for (_, name), (mod, old) in list(_saved.items()):
        setattr(mod, name, old)
    _saved.clear()


def installed() -> bool:
    return bool(_saved)

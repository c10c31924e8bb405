"""CPU-only: cnvkit_b200.install rebinds the reference's seams and
forwards out-of-scope options to the reference's own callables.  Needs the reference tree, so it
runs in the authoring container only."""
import sys

import pytest

from cnvkit_b200 import install as I


@pytest.mark.needs_reference
def test_install_rebinds_and_restores():
    from oracle import refshim

    R = refshim.load()
    orig = (R.fix.do_fix, R.reference.do_reference, R.segmentation.do_segmentation, R.smoothing.rolling_median)
    try:
        I.install()
        assert I.installed()
        import cnvkit_b200.smoothing as S

        assert R.smoothing.rolling_median is S.rolling_median
        for fn, o in zip((R.fix.do_fix, R.reference.do_reference, R.segmentation.do_segmentation), orig):
            assert fn is not o and fn.__wrapped__ is o
        I.install()  # idempotent
    finally:
        I.uninstall()
    assert (R.fix.do_fix, R.reference.do_reference, R.segmentation.do_segmentation,
            R.smoothing.rolling_median) == orig
    assert not I.installed()


def test_routing_predicates():
    assert I._segment_needs_reference((None, "hmm"), {})
    assert not I._segment_needs_reference((None, "cbs"), {"smooth_cbs": True})    # smooth.CNA runs on the GPU
    assert I._segment_needs_reference((None, "cbs"), {"variants": [1]})
    assert not I._segment_needs_reference((None, "haar"), {"variants": [1]})          # depth + BAF union runs on the GPU
    assert not I._segment_needs_reference((None, "cbs"), {"threshold": 1e-4})
    assert not I._segment_needs_reference((None,), {"method": "haar"})
    assert I._fix_needs_reference((1, 2, 3), {"bias_smoother": "loess"})
    assert I._fix_needs_reference((1, 2, 3, None, True, True, True, True), {})
    assert not I._fix_needs_reference((1, 2, 3), {})
    assert I._reference_needs_reference((["a"],), {"do_cluster": True})
    assert not I._reference_needs_reference((["a"],), {})


@pytest.mark.needs_reference
def test_out_of_scope_options_reach_the_reference():
    from oracle import refshim

    R = refshim.load()
    calls = []
    real = R.segmentation.do_segmentation
    R.segmentation.do_segmentation = lambda *a, **k: calls.append((a, k)) or "theirs"
    try:
        I.install()
        assert R.segmentation.do_segmentation("cnarr", "hmm") == "theirs"
        assert calls and calls[0][0][1] == "hmm"
    finally:
        I.uninstall()
        R.segmentation.do_segmentation = real


@pytest.mark.needs_reference
def test_segmetrics_seam_merges_reference_columns_with_ours(golden):
    """Statistics outside the GPU path come from the reference, 'ci' from ours, columns in the
    reference's order (segmetrics.py:93-119).  The GPU call is replaced by the numpy oracle here."""
    import numpy as np

    from oracle import refshim
    from oracle import segmetrics as OS

    R = refshim.load()
    from cnvlib import segmetrics as ref_sm

    import os

    cnr = R.read_cna(os.path.join(R.root, "test", "formats", "amplicon.cnr"))
    cns = R.read_cna(os.path.join(R.root, "test", "formats", "amplicon.cns"))

    def ours(cnarr, segarr, interval_stats=(), alpha=0.05, bootstraps=100, smoothed=10, skip_low=False):
        d, s = cnarr.data, segarr.data
        lo, hi = OS.bin_ranges(d["chromosome"], d["start"], d["end"], s["chromosome"], s["start"], s["end"])
        out = segarr.copy()
        out["ci_lo"], out["ci_hi"] = OS.ci_table(d["log2"].to_numpy(), d["weight"].to_numpy(), lo, hi, alpha,
                                                 bootstraps, smoothed)
        return out

    seam = I._segmetrics_seam(ours, ref_sm.do_segmetrics)
    kw = dict(location_stats=["mean", "median"], spread_stats=["stdev"], interval_stats=["pi", "ci"], smoothed=False)
    got = seam(cnr, cns, **kw)
    exp = ref_sm.do_segmetrics(cnr, cns, **kw)
    assert list(got.data.columns) == list(exp.data.columns)
    for col in exp.data.columns[len(cns.data.columns):]:
        np.testing.assert_allclose(got[col].to_numpy(), exp[col].to_numpy(), rtol=0, atol=1e-9, equal_nan=True)
    only_ci = seam(cnr, cns, interval_stats=["ci"], smoothed=False)
    assert list(only_ci.data.columns) == list(cns.data.columns) + ["ci_lo", "ci_hi"]
    no_ci = seam(cnr, cns, location_stats=["mean"])
    assert "ci_lo" not in no_ci.data.columns and "mean" in no_ci.data.columns


def test_in_scope_inputs_never_fall_back_to_the_reference():
    """north_star: "no CPU fallback".  Only the explicit option predicates forward a call to cnvlib; an in-scope
    call that the GPU path cannot serve raises instead of silently rerunning on the CPU."""
    def ours(x):
        if x == "odd":
            raise NotImplementedError("an input the GPU path declines")
        return "ours"

    seam = I._route(ours, lambda x: "theirs", lambda a, k: a[0] == "out-of-scope option")
    assert seam("plain") == "ours" and seam("out-of-scope option") == "theirs"
    with pytest.raises(NotImplementedError):
        seam("odd")

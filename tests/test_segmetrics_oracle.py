"""CPU-only: the bootstrap-CI oracle (oracle/segmetrics.py) against vectors the unmodified reference
produced (tests/golden/segmetrics.npz, written by oracle/make_golden.py gen_segmetrics), the restated
PCG64 + Lemire stream against numpy itself, and the host logic of cnvkit_b200.segmetrics."""
import os
import re

import numpy as np
import pandas as pd
import pytest

from oracle import segmetrics as OS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT_CASES = ((0.05, 100), (0.5, 100), (0.05, 30), (0.01, 100))
TOL = 1e-9    # closed-form jackknife vs the reference's concatenations; everything else is the same arithmetic


def test_unit_segments_match_reference(golden):
    g = golden("segmetrics")
    for i, k in enumerate(g["unit_sizes"]):
        v, w = g[f"unit{i}_values"], g[f"unit{i}_weights"]
        assert len(v) == k
        for alpha, boots in UNIT_CASES:
            ci = OS.confidence_interval_bootstrap(v, w, alpha, boots, False)
            np.testing.assert_allclose(ci, g[f"unit{i}_ci_a{alpha}_b{boots}"], rtol=0, atol=TOL)


def test_amplicon_table_matches_reference(golden):
    g = golden("segmetrics")
    lo, hi = OS.bin_ranges(g["cnr__chromosome"], g["cnr__start"], g["cnr__end"], g["cns__chromosome"],
                           g["cns__start"], g["cns__end"])
    k = hi - lo
    np.testing.assert_array_equal(k, g["sel_hi"] - g["sel_lo"])
    assert ((k == 0) | ((lo == g["sel_lo"]) & (hi == g["sel_hi"]))).all()
    assert (k == 0).any() and (k == 1).any()          # amplicon.cns has empty and single-bin segments
    for tag, kw, mask in (("bca", dict(smoothed=False), k >= 0),
                          ("bca_a0.5_b50", dict(smoothed=False, alpha=0.5, bootstraps=50), k >= 0),
                          ("default", dict(), (k > 10) | (k < 2))):
        ci_lo, ci_hi = OS.ci_table(g["cnr__log2"], g["cnr__weight"], lo, hi, **kw)
        np.testing.assert_allclose(ci_lo[mask], g[f"amplicon_{tag}_ci_lo"][mask], rtol=0, atol=TOL, equal_nan=True)
        np.testing.assert_allclose(ci_hi[mask], g[f"amplicon_{tag}_ci_hi"][mask], rtol=0, atol=TOL, equal_nan=True)


def test_smoothed_interval_is_statistically_the_references(golden):
    # the reference's smoothing noise is unseeded: compare widths and centres with its 8 realisations
    g = golden("segmetrics")
    ratios = []
    for i, k in enumerate(g["unit_sizes"]):
        v, w = g[f"unit{i}_values"], g[f"unit{i}_weights"]
        reps = g[f"unit{i}_ci_smoothed_a0.5_reps"]
        ci = OS.confidence_interval_bootstrap(v, w, 0.5, 100, True, seg_index=i)
        ref_w = (reps[:, 1] - reps[:, 0]).mean()
        ratios.append((ci[1] - ci[0]) / ref_w)
        centre = reps.mean()
        assert abs(ci.mean() - centre) < 1.5 * ref_w
    ratios = np.asarray(ratios)
    assert 0.6 < ratios.min() and ratios.max() < 1.6 and abs(np.median(ratios) - 1) < 0.15


@pytest.mark.parametrize("k", [2, 3, 7, 100, 5000, 123457, 3_000_000_000, 2_500_000_001, 1 << 31])
def test_restated_pcg64_lemire_stream_is_numpys(k):
    mine, used = OS.pcg64_lemire_indices(k, 4000)
    ref = np.random.default_rng(0xA5EED).integers(0, k, size=4000)
    np.testing.assert_array_equal(mine, ref)
    if k > 1 << 31:
        assert used > 4000        # the rejection branch ran


def test_generator_constants():
    st = np.random.PCG64(0xA5EED).state["state"]
    assert st["state"] == OS.PCG_STATE_A5EED and st["inc"] == OS.PCG_INC_A5EED
    src = open(os.path.join(ROOT, "cnvkit_b200", "csrc", "capi.cu")).read()
    words = [int(x, 16) for x in re.findall(r"pcg\[\d\] = (0x[0-9a-f]+)ull", src)]
    assert (words[0] << 64 | words[1]) == OS.PCG_STATE_A5EED and (words[2] << 64 | words[3]) == OS.PCG_INC_A5EED


def test_host_bin_ranges_and_argument_handling(golden):
    from cnvkit_b200 import segmetrics as SM

    g = golden("segmetrics")
    bins = pd.DataFrame({"chromosome": g["cnr__chromosome"].astype(object), "start": g["cnr__start"], "end": g["cnr__end"]})
    segs = pd.DataFrame({"chromosome": g["cns__chromosome"].astype(object), "start": g["cns__start"], "end": g["cns__end"]})
    lo, hi = SM.segment_bin_ranges(bins, segs)
    olo, ohi = OS.bin_ranges(bins["chromosome"], bins["start"], bins["end"], segs["chromosome"], segs["start"], segs["end"])
    np.testing.assert_array_equal(hi - lo, ohi - olo)
    assert ((hi == lo) | ((lo == olo) & (hi == ohi))).all()
    # unknown chromosome -> nothing selected; unsorted starts -> the reference's ValueError
    other = pd.DataFrame({"chromosome": ["chrNope"], "start": [0], "end": [10 ** 9]})
    assert SM.segment_bin_ranges(bins, other)[1][0] == SM.segment_bin_ranges(bins, other)[0][0]
    bad = bins.iloc[::-1].reset_index(drop=True)
    with pytest.raises((ValueError, NotImplementedError)):
        SM.segment_bin_ranges(bad, segs)
    assert SM._smooth_upto(True) == 2 ** 31 - 1 and SM._smooth_upto(False) == -1 and SM._smooth_upto(10) == 10
    with pytest.raises(ValueError, match="alpha must be between 0 and 1"):
        SM.make_params(1.5)
    assert SM.effective_bootstraps(0.05, 100, warn=False) == 100
    assert SM.effective_bootstraps(0.01, 100, warn=False) == 200


def test_out_of_scope_statistics_are_refused(golden):
    from conftest import golden_table

    from cnvkit_b200 import segmetrics as SM

    g = golden("segmetrics")
    cnr, cns = golden_table(g, "cnr"), golden_table(g, "cns")
    with pytest.raises(NotImplementedError):
        SM.do_segmetrics(cnr, cns, location_stats=["mean"], interval_stats=["ci"])
    with pytest.raises(NotImplementedError):
        SM.do_segmetrics(cnr, cns, interval_stats=["pi"])
    out = SM.do_segmetrics(cnr, cns)        # nothing requested: a copy, no device call
    assert list(out.data.columns) == list(cns.data.columns) and out is not cns


@pytest.mark.needs_reference
def test_host_logic_with_nan_weights_and_skip_low_matches_reference(monkeypatch):
    """do_segmetrics' host side (bin selection, NaN-weight compaction, skip_low) against the unmodified
    reference, with the device call replaced by the numpy oracle."""
    from cnvkit_b200 import cohort
    from cnvkit_b200 import segmetrics as SM
    from oracle import refshim

    R = refshim.load()
    from cnvlib import segmetrics as ref_sm

    cnr = R.read_cna(os.path.join(R.root, "test", "formats", "amplicon.cnr"))
    cns = R.read_cna(os.path.join(R.root, "test", "formats", "amplicon.cns"))
    d = cnr.data.copy()
    rng = np.random.default_rng(2)
    d.loc[rng.choice(len(d), 40, replace=False), "weight"] = np.nan
    d.loc[rng.choice(len(d), 25, replace=False), "log2"] = -30.0          # dropped by skip_low
    cnr = cnr.as_dataframe(d)

    def fake(values, weights, lo, hi, alpha=0.05, bootstraps=100, smoothed=False, **kw):
        return OS.ci_table(np.asarray(values), np.asarray(weights), lo, hi, alpha, bootstraps, smoothed)

    monkeypatch.setattr(SM, "bootstrap_ci", fake)
    for kw in (dict(smoothed=False), dict(smoothed=False, skip_low=True), dict(smoothed=False, alpha=0.5, bootstraps=3)):
        got = SM.do_segmetrics(cnr, cns, interval_stats=["ci"], **kw)
        want = ref_sm.do_segmetrics(cnr, cns, interval_stats=["ci"], **kw)
        assert list(got.data.columns) == list(want.data.columns)
        np.testing.assert_allclose(got["ci_lo"].to_numpy(), want["ci_lo"].to_numpy(), rtol=0, atol=TOL, equal_nan=True)
        np.testing.assert_allclose(got["ci_hi"].to_numpy(), want["ci_hi"].to_numpy(), rtol=0, atol=TOL, equal_nan=True)
    assert cohort._ci_options(None) is None and cohort._ci_options(False) is None
    assert cohort._ci_options(True) == dict(alpha=0.5, bootstraps=100, smoothed=True)
    assert cohort._ci_options(dict(alpha=0.1)) == dict(alpha=0.1, bootstraps=100, smoothed=10)
    with pytest.raises(ValueError):
        cohort._ci_options(dict(alpah=0.1))

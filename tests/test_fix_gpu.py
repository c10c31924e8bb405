"""GPU parity for the fix path (csrc/fix.cu, select.cu, biweight.cu) against vectors produced by
the unmodified reference (tests/golden/regression.npz, fix_units.npz) and the numpy oracle.

Bars (BASELINE.json north_star): log2 ratios bit-exact; weights within 1e-12 relative (numpy's
pairwise summation order is not replayed on the device)."""
import numpy as np
import pandas as pd
import pytest

from conftest import golden_table
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    from cnvkit_b200 import fix

    return fix


def test_center_by_window_golden(F, golden):
    g = golden("regression")
    for i in range(int(g["cbw_ncalls"])):
        got = F.center_by_window(g[f"cbw{i}_before"], float(g[f"cbw{i}_fraction"]), g[f"cbw{i}_key"])
        np.testing.assert_array_equal(got, g[f"cbw{i}_after"])


def test_center_by_window_single_bin_and_ties(F):
    # test_cnvlib.py:713-740: a single bin must not crash; bias == value -> 0
    np.testing.assert_array_equal(F.center_by_window(np.array([1.5]), 0.5, np.array([0.4])), [0.0])
    rng = np.random.default_rng(4)
    x = rng.normal(0, 1, 3001)
    key = np.round(rng.uniform(0.3, 0.7, 3001), 2)  # heavy ties -> stable order matters
    key[::97] = np.nan  # NaN keys sort last (rmask can be NaN)
    np.testing.assert_array_equal(F.center_by_window(x, 0.05, key), O.center_by_window(x, key, 0.05))


def test_edge_bias_golden(F, golden):
    g = golden("regression")
    for i in range(int(g["edge_ncalls"])):
        got = F.get_edge_bias(g[f"edge{i}_chromosome"].astype(object), g[f"edge{i}_start"], g[f"edge{i}_end"], 250)
        np.testing.assert_array_equal(got, g[f"edge{i}_bias"])


def test_edge_identities(F):
    # test_cnvlib.py:689-711: a lone wide tile loses insert/(2*size); abutting tiles gain
    chrom = np.array(["chr1"] * 3, dtype=object)
    start = np.array([1000, 1500, 5000])
    end = np.array([1500, 2000, 5100])
    got = F.get_edge_bias(chrom, start, end, 250)
    np.testing.assert_array_equal(got, O.edge_bias(chrom, start, end, 250))
    assert got[2] == -(250 / (2 * 100) - (250 - 100) ** 2 / (2 * 250 * 100))


def test_center_all_golden(F, golden):
    g = golden("fix_units")
    chrom = g["cnr2__chromosome"].astype(object)
    for skip in (0, 1):
        got = F.center_all(g["cnr2__log2"] + 0.37, chrom, depth=g["cnr2__depth"], skip_low=bool(skip))
        np.testing.assert_array_equal(got, g[f"cnr2_centered_skip{skip}"])


def test_biweight_golden(golden):
    from cnvkit_b200 import descriptives as D

    g = golden("fix_units")
    for i in range(int(g["bw_n"])):
        x = g[f"bw{i}_x"]
        np.testing.assert_allclose(D.biweight_location(x), g[f"bw{i}_loc"], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(D.biweight_midvariance(x), g[f"bw{i}_var"], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(D.biweight_midvariance(x, initial=0.1), g[f"bw{i}_var_init0"], rtol=1e-12,
                                   atol=1e-300)
    assert np.isnan(D.biweight_location(np.array([np.nan])))
    assert D.biweight_midvariance(np.array([2.0])) == 0


def test_do_fix_regression_fixture(F, golden):
    """test/test_regression.py:91-126: the reference's own frozen pipeline, stage `fix`."""
    g = golden("regression")
    tgt, anti, ref = (golden_table(g, k) for k in ("sample_target", "sample_antitarget", "reference"))
    want = golden_table(g, "cnr")
    before = tgt.data.copy()
    got = F.do_fix(tgt, anti, ref)
    pd.testing.assert_frame_equal(tgt.data, before)  # inputs are not mutated
    assert list(got.data.columns) == list(want.data.columns)
    for col in ("chromosome", "start", "end", "gene"):
        np.testing.assert_array_equal(got.data[col].to_numpy(), want.data[col].to_numpy())
    np.testing.assert_array_equal(got.data["log2"].to_numpy(), want.data["log2"].to_numpy())  # bit-exact
    np.testing.assert_array_equal(got.data["depth"].to_numpy(), want.data["depth"].to_numpy())
    np.testing.assert_allclose(got.data["weight"].to_numpy(), want.data["weight"].to_numpy(), rtol=1e-12, atol=0)


def test_do_fix_errors_and_empty_antitarget(F, golden):
    g = golden("regression")
    tgt, anti, ref = (golden_table(g, k) for k in ("sample_target", "sample_antitarget", "reference"))
    empty = anti[:0]
    got = F.do_fix(tgt, empty, ref)  # amplicon-style: no antitargets (fix.py:235-236)
    assert len(got) > 0 and "weight" in got
    dup = tgt.as_dataframe(pd.concat([tgt.data, tgt.data.iloc[:3]], ignore_index=True))
    with pytest.raises(ValueError, match="Duplicated genomic coordinates in sample set"):
        F.do_fix(dup, anti, ref)
    with pytest.raises(ValueError, match="Reference is missing"):
        F.do_fix(tgt, anti, ref[100:])
    shifted = ref.as_dataframe(ref.data.assign(start=ref.data["start"] + 1))
    with pytest.raises(ValueError, match="Reference shares no bins"):
        F.do_fix(tgt, anti, shifted)
    with pytest.raises(ValueError, match="Unknown bias_smoother"):
        F.do_fix(tgt, anti, ref, bias_smoother="nope")


def test_center_by_precomputed_order_equals_center_by_window(F):
    """Cohort entry points: cnvk_trait_order_dev + cnvk_center_by_order_dev reproduce
    cnvk_center_by_window_dev bit for bit (the visiting order depends on the trait and the bin count only)."""
    import torch

    from cnvkit_b200 import _dev, _lib, smoothing

    rng = np.random.default_rng(12)
    for n, frac in ((1, 0.5), (7, 0.5), (2000, 0.05), (60000, 0.01)):
        log2 = rng.normal(0, 0.3, n)
        key = np.round(rng.beta(8, 8, n), 3)          # many ties: the stable order matters
        want = F.center_by_window(log2, frac, key)
        d_log2, d_key = _dev.up(log2), _dev.up(key)
        d_perm = _dev.shuffle_permutation(n)
        d_ord = _dev.empty(n, torch.int32)
        _lib.call("cnvk_trait_order_dev", _dev.ptr(d_key), _dev.ptr(d_perm), n, _dev.ptr(d_ord), _dev.stream())
        wing = smoothing._width2wing(frac, n) if n >= 2 else 1
        _lib.call("cnvk_center_by_order_dev", _dev.ptr(d_log2), _dev.ptr(d_ord), n, wing, _dev.stream())
        np.testing.assert_array_equal(d_log2.cpu().numpy(), want)
        order = d_ord.cpu().numpy().view(np.uint32)
        assert sorted(order.tolist()) == list(range(n))


def test_row_shuffled_sample_gives_the_sorted_result(F, golden):
    """API callers may pass tables whose chromosomes are not contiguous runs: load_adjust_coverages re-orders
    them genomically, and the device columns uploaded before that must follow the same order."""
    g = golden("regression")
    tgt, anti, ref = (golden_table(g, k) for k in ("sample_target", "sample_antitarget", "reference"))
    want = F.do_fix(tgt, anti, ref)
    rng = np.random.default_rng(3)
    shuf = tgt.as_dataframe(tgt.data.iloc[rng.permutation(len(tgt))].reset_index(drop=True))
    got = F.do_fix(shuf, anti, ref)
    for c in ("chromosome", "start", "end"):
        assert got.data[c].tolist() == want.data[c].tolist()
    np.testing.assert_array_equal(got.data["log2"].to_numpy(), want.data["log2"].to_numpy())
    np.testing.assert_allclose(got.data["weight"].to_numpy(), want.data["weight"].to_numpy(), rtol=1e-12)

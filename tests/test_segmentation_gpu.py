"""GPU parity for the segmentation driver (cnvkit_b200.segmentation.do_segmentation): pre-filters,
%.6g quantisation, CBS / Haar batched over arms, transfer_fields -- against tables produced by the
unmodified reference (haar) and against the CPU oracle pipeline (cbs)."""
import numpy as np
import pandas as pd
import pytest

from conftest import golden_table
from oracle import cbs as OC
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from cnvkit_b200 import segmentation

    return segmentation


def assert_tables_match(got, want, float_rtol=1e-9):
    assert list(got.data.columns) == list(want.data.columns)
    assert len(got) == len(want)
    for col in want.data.columns:
        g, w = got.data[col].to_numpy(), want.data[col].to_numpy()
        if col in ("log2", "depth", "weight"):
            np.testing.assert_allclose(g.astype(float), w.astype(float), rtol=float_rtol, atol=1e-12, err_msg=col)
        else:
            np.testing.assert_array_equal(g, w, err_msg=col)


def test_outlier_mask_golden(S, golden):
    h = golden("haar")
    for factor, key in ((10, "outlier_mask_f10"), (3, "outlier_mask_f3")):
        got = S.drop_outliers_mask(h["cnr__log2"], h["arm_offsets"], 50, factor)
        np.testing.assert_array_equal(got, h[key])
    g = golden("smoothing")
    for i in range(4):
        x = g[f"sg{i}_x"]
        np.testing.assert_array_equal(S.drop_outliers_mask(x, np.array([0, len(x)]), 50, 3), g[f"sg{i}_mask"])
    # arms of <= 50 bins are never examined (smoothing.py:458-459)
    assert not S.drop_outliers_mask(np.r_[np.zeros(49), 100.0], np.array([0, 50]), 50, 3).any()


def test_round_sig6_matches_printf():
    import torch
    from cnvkit_b200 import _dev, _lib

    rng = np.random.default_rng(6)
    x = np.concatenate([
        rng.normal(0, 1, 20000), rng.normal(0, 1, 2000) * 10.0 ** rng.integers(-12, 12, 2000),
        np.array([0.0, -0.0, 1.0, 0.1234565, 0.1234575, 1234565.0, 1234575.0, 9.999995, 9999995.0, 0.99999949999,
                  123456.5, 123457.5, 2.5e-5, 1e-300, 1e300, np.nan, np.inf, -np.inf, 5e-324, 1.0000005,
                  0.000999999500000001]),
        np.round(rng.uniform(-30, 30, 5000), 6), (rng.integers(0, 2_000_000, 5000) * 5 + 5) / 1e7,
    ])
    d_x = _dev.up(x)
    d_o = _dev.empty(len(x), torch.float64)
    unhandled = np.zeros(1, dtype=np.uint64)
    _lib.call("cnvk_round_sig_dev", _dev.ptr(d_x), len(x), 6, _dev.ptr(d_o), _lib.ptr(unhandled), _dev.stream())
    got = d_o.cpu().numpy()
    want = np.array([float("%.6g" % v) for v in x])
    handled = ~((np.abs(x) < 1e-16) & (x != 0)) & ~(np.abs(x) > 1e21) | ~np.isfinite(x)
    np.testing.assert_array_equal(got[handled], want[handled])
    assert int(unhandled[0]) == int((~handled).sum())
    # idempotent: a .cnr read from disk is already quantised
    d_o2 = _dev.empty(len(x), torch.float64)
    _lib.call("cnvk_round_sig_dev", _dev.ptr(d_o), len(x), 6, _dev.ptr(d_o2), 0, _dev.stream())
    np.testing.assert_array_equal(d_o2.cpu().numpy()[handled], got[handled])


def test_do_segmentation_haar_reference_tables(S, golden):
    h = golden("haar")
    cnr = golden_table(h, "cnr")
    assert_tables_match(S.do_segmentation(cnr, "haar"), golden_table(h, "cns_haar"))
    assert_tables_match(S.do_segmentation(cnr, "haar", threshold=0.01, skip_low=True), golden_table(h, "cns_haar_t01"))
    g = golden("regression")
    cnr2 = golden_table(g, "cnr")
    assert_tables_match(S.do_segmentation(cnr2, "haar"), golden_table(g, "cns_haar"))  # test_regression.py stage 2
    assert_tables_match(S.do_segmentation(cnr2, "haar", skip_outliers=0), golden_table(g, "cns_haar_nooutlier"))


def oracle_cbs_pipeline(cnr, threshold=1e-4, skip_outliers=10):
    """Independent CPU pipeline: by_arm -> filters -> %.6g -> oracle CBS -> (arm, bin lo, bin hi, mean)."""
    from cnvkit_b200.cnary import arm_ranges

    d = cnr.data.reset_index(drop=True)
    x, w = d["log2"].to_numpy(), d["weight"].to_numpy()
    rows = []
    for chrom, lo, hi in arm_ranges(d):
        idx = np.arange(lo, hi)
        idx = idx[np.isfinite(x[idx])]
        if skip_outliers:
            idx = idx[~O.outlier_mask(x[idx], factor=skip_outliers)]
        idx = idx[~(w[idx] == 0)]
        xq = np.array([float("%.6g" % v) for v in x[idx]])
        wq = np.array([float("%.6g" % v) for v in w[idx]])
        ok = ~np.isnan(wq) & (wq > 0) & ~np.isnan(xq)
        idx, xq, wq = idx[ok], xq[ok], wq[ok]
        if not len(idx):
            continue
        arm, slo, shi, mean, _ = OC.segment(xq, wq, [0, len(idx)], OC.make_params(alpha=threshold))
        for a, b, m in zip(slo, shi, mean):
            rows.append((chrom, d["start"][idx[a]], d["end"][idx[b - 1]], int(b - a), m, lo, hi))
    return rows


def test_do_segmentation_cbs_vs_oracle_pipeline(S, golden):
    """C1: cnvkit.py segment -m cbs on the reference's exome .cnr (test/formats/p2-20_1.cnr)."""
    h = golden("haar")
    cnr = golden_table(h, "cnr")
    got = S.do_segmentation(cnr, "cbs")
    want = oracle_cbs_pipeline(cnr)
    assert len(got) == len(want) > 47
    assert list(got.data.columns) == ["chromosome", "start", "end", "gene", "log2", "depth", "probes", "weight"]
    d = cnr.data
    g = got.data
    # interior boundaries identical; arm-first/last segments stretched to the arm's bins (transfer_fields)
    for k, (chrom, s, e, probes, mean, alo, ahi) in enumerate(want):
        row = g.iloc[k]
        assert row["chromosome"] == chrom and row["probes"] == probes
        first = k == 0 or want[k - 1][5] != alo
        last = k == len(want) - 1 or want[k + 1][5] != alo
        assert row["start"] == (d["start"][alo] if first else s)
        assert row["end"] == (d["end"][ahi - 1] if last else e)
        assert row["log2"] == pytest.approx(mean, rel=1e-9)
    # transfer_fields aggregates: weights of all bins inside each span
    for _, row in g.iterrows():
        sel = (d["chromosome"] == row["chromosome"]) & (d["end"] > row["start"]) & (d["start"] < row["end"])
        assert row["weight"] == pytest.approx(d["weight"][sel].sum(), rel=1e-9)
        assert row["depth"] == pytest.approx(np.average(d["depth"][sel], weights=d["weight"][sel]), rel=1e-9)
    assert g["gene"].map(lambda s: "Antitarget" not in s.split(",")).all()


def test_api_behaviour(S, golden):
    h = golden("haar")
    cnr = golden_table(h, "cnr")
    with pytest.raises(ValueError, match="'method' must be one of"):
        S.do_segmentation(cnr, "bogus")
    empty = cnr[:0]
    assert S.do_segmentation(empty, "cbs") is empty           # test_segmentation.py:188-212
    out, text = S.do_segmentation(cnr[:300], "cbs", save_dataframe=True, processes=2)
    assert len(out) > 0 and text.startswith('"ID"\t"chrom"') and text.count("\n") == len(out) + 1
    # the SEG text is R's per-arm stdout, i.e. the rows BEFORE transfer_fields stretches arm ends
    # (cnvlib/segmentation/__init__.py:180-188): 1-based first-bin start, last-bin end, filtered bin counts
    rows = [ln.split("\t") for ln in text.strip().split("\n")[1:]]
    assert [int(r[4]) for r in rows] == out.data["probes"].tolist()
    sub = cnr[:300].data
    starts1 = set((sub["start"] + 1).tolist())
    assert all(int(r[2]) in starts1 for r in rows) and all(int(r[3]) in set(sub["end"].tolist()) for r in rows)
    assert all(r[0] == f'"{cnr.sample_id}"' for r in rows)
    assert S.do_segmentation(cnr[:300], "haar", save_dataframe=True)[1] == ""
    # an arm whose weights are all NaN is segmented with weight 1.0 (the R script runs per arm, cbs.py:22-28);
    # the other arms keep their weights and the arm does not vanish
    part = cnr[:3000].copy()
    first_chrom = part.data["chromosome"].iat[0]
    part.data.loc[part.data["chromosome"] == first_chrom, "weight"] = np.nan
    got = S.do_segmentation(part, "cbs")
    assert first_chrom in set(got.data["chromosome"])
    # NaN / inf log2 bins are dropped before segmentation (test_segmentation.py:741-806)
    bad = cnr[:2000].copy()
    bad.data.loc[bad.data.index[5], "log2"] = np.nan
    bad.data.loc[bad.data.index[9], "log2"] = np.inf
    res = S.do_segmentation(bad, "haar")
    assert np.isfinite(res.data["log2"]).all()
    # single bin -> one segment carrying that bin's log2 (test_r.py:215-231)
    one = S.do_segmentation(cnr[:1], "cbs")
    assert len(one) == 1 and one.data["log2"].iat[0] == pytest.approx(float("%.6g" % cnr.data["log2"].iat[0]))

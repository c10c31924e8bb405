"""GPU parity for the bootstrap confidence interval (csrc/bootstrap.cu through the C ABI) against the
reference's own vectors (tests/golden/segmetrics.npz) and the numpy oracle (oracle/segmetrics.py).

Tolerance: 1e-9 absolute on log2 values of order 1.  The resampled indices and every replicate mean are
reproduced exactly (the BCa count of replicates below the observed mean depends on it); what is left is the
closed-form jackknife and CUDA's normcdf / normcdfinv against scipy's in the BCa correction."""
import numpy as np
import pytest

from oracle import segmetrics as OS

pytestmark = pytest.mark.gpu
TOL = 1e-9
UNIT_CASES = ((0.05, 100), (0.5, 100), (0.05, 30), (0.01, 100))


@pytest.fixture(scope="module")
def SM():
    from cnvkit_b200 import segmetrics

    return segmetrics


def concat_units(g):
    vals, wts, lo, hi = [], [], [], []
    pos = 0
    for i, k in enumerate(g["unit_sizes"]):
        vals.append(g[f"unit{i}_values"])
        wts.append(g[f"unit{i}_weights"])
        lo.append(pos)
        pos += int(k)
        hi.append(pos)
    return np.concatenate(vals), np.concatenate(wts), np.asarray(lo), np.asarray(hi)


def test_unit_segments_match_reference_vectors(SM, golden):
    g = golden("segmetrics")
    v, w, lo, hi = concat_units(g)
    for alpha, boots in UNIT_CASES:
        ci_lo, ci_hi = SM.bootstrap_ci(v, w, lo, hi, alpha, boots, smoothed=False)
        ref = np.array([g[f"unit{i}_ci_a{alpha}_b{boots}"] for i in range(len(lo))])
        np.testing.assert_allclose(ci_lo, ref[:, 0], rtol=0, atol=TOL)
        np.testing.assert_allclose(ci_hi, ref[:, 1], rtol=0, atol=TOL)


def test_extreme_replicates_equal_numpys(SM):
    # weights of exactly 1 make the smoothing noise vanish (sqrt(1 - w) = 0) and smoothed=True skips BCa, so
    # the interval is a plain percentile pair of the bootstrap distribution; with alpha/2 ~ 1/(B-1) these
    # are (next to) its second-smallest and second-largest replicate means -- numpy's draws, numpy's sums.
    rng = np.random.default_rng(7)
    alpha, boots = 0.02, 101
    for k in (2, 5, 8, 77, 128, 129, 1000, 4099):
        v = rng.normal(0, 1, k)
        w1 = np.ones(k)
        idx = np.random.default_rng(0xA5EED).integers(0, k, size=(boots, k))
        dist = np.array([np.average(v[i], weights=w1[i]) for i in idx])
        exp = np.percentile(dist, [100 * (alpha / 2), 100 * (1 - alpha / 2)])
        lo1, hi1 = SM.bootstrap_ci(v, w1, [0], [k], alpha=alpha, bootstraps=boots, smoothed=True)
        np.testing.assert_allclose([lo1[0], hi1[0]], exp, rtol=0, atol=2e-15)


def test_smoothed_path_equals_oracle_with_the_same_noise(SM, golden):
    g = golden("segmetrics")
    v, w, lo, hi = concat_units(g)
    for alpha, smoothed in ((0.5, True), (0.05, True), (0.05, 10), (0.5, 130)):
        ci_lo, ci_hi = SM.bootstrap_ci(v, w, lo, hi, alpha, 100, smoothed=smoothed)
        olo, ohi = OS.ci_table(v, w, lo, hi, alpha, 100, smoothed)
        np.testing.assert_allclose(ci_lo, olo, rtol=0, atol=TOL)
        np.testing.assert_allclose(ci_hi, ohi, rtol=0, atol=TOL)


def test_amplicon_do_segmetrics(SM, golden):
    from conftest import golden_table

    g = golden("segmetrics")
    cnr, cns = golden_table(g, "cnr"), golden_table(g, "cns")
    k = g["sel_hi"] - g["sel_lo"]
    for tag, kw, mask in (("bca", dict(smoothed=False), k >= 0),
                          ("bca_a0.5_b50", dict(smoothed=False, alpha=0.5, bootstraps=50), k >= 0),
                          ("default", dict(), (k > 10) | (k < 2))):
        sm = SM.do_segmetrics(cnr, cns, interval_stats=["ci"], **kw)
        assert list(sm.data.columns) == list(cns.data.columns) + ["ci_lo", "ci_hi"]
        np.testing.assert_allclose(sm["ci_lo"].to_numpy()[mask], g[f"amplicon_{tag}_ci_lo"][mask], rtol=0, atol=TOL,
                                   equal_nan=True)
        np.testing.assert_allclose(sm["ci_hi"].to_numpy()[mask], g[f"amplicon_{tag}_ci_hi"][mask], rtol=0, atol=TOL,
                                   equal_nan=True)
    # the batch call (batch.py:559-566) and the reference's property tests (test_analysis.py:270-311)
    sm = SM.do_segmetrics(cnr, cns, interval_stats=["ci"], alpha=0.5, smoothed=True)
    lo, hi = OS.bin_ranges(g["cnr__chromosome"], g["cnr__start"], g["cnr__end"], g["cns__chromosome"],
                           g["cns__start"], g["cns__end"])
    olo, ohi = OS.ci_table(g["cnr__log2"], g["cnr__weight"], lo, hi, 0.5, 100, True)
    np.testing.assert_allclose(sm["ci_lo"].to_numpy(), olo, rtol=0, atol=TOL, equal_nan=True)
    np.testing.assert_allclose(sm["ci_hi"].to_numpy(), ohi, rtol=0, atol=TOL, equal_nan=True)
    sm50 = SM.do_segmetrics(cnr, cns, interval_stats=["ci"], bootstraps=50, smoothed=True)
    d = sm50.data
    means = np.array([np.mean(g["cnr__log2"][a:b]) if b > a else np.nan for a, b in zip(lo, hi)])
    one = (d["probes"] == 1).to_numpy() & (k == 1)
    assert one.any()
    assert ((d["ci_lo"].to_numpy()[one] <= means[one]) & (means[one] <= d["ci_hi"].to_numpy()[one])).all()
    many = (k > 3)
    assert (d["ci_lo"].to_numpy()[many] < means[many]).all() and (d["ci_hi"].to_numpy()[many] > means[many]).all()
    # the reference's batch-mode intervals (one unseeded realisation): same scale
    ref_w = g["amplicon_batch_ci_hi"] - g["amplicon_batch_ci_lo"]
    our_w = sm["ci_hi"].to_numpy() - sm["ci_lo"].to_numpy()
    ok = k >= 2
    ratio = our_w[ok] / ref_w[ok]
    assert 0.7 < np.median(ratio) < 1.4


def test_long_segment_with_rejected_draws(SM):
    # 2^32 mod k is large here, so thousands of raw draws are rejected and most chunks of the stream shift
    k, alpha = 700_001, 0.2
    rng = np.random.default_rng(11)
    v = rng.normal(0.3, 0.25, k)
    w = rng.uniform(0.5, 1, k)
    _, used = OS.pcg64_lemire_indices(k, 200_000)
    assert used > 200_000
    ci_lo, ci_hi = SM.bootstrap_ci(v, w, [0], [k], alpha, 20, smoothed=False)
    exp = OS.confidence_interval_bootstrap(v, w, alpha, 20, False)
    np.testing.assert_allclose([ci_lo[0], ci_hi[0]], exp, rtol=0, atol=TOL)
    # several segments inside one table, in a different order than the rows
    lo = np.array([500_000, 0, 200_000, 699_999, 5])
    hi = np.array([700_001, 150_000, 200_900, 700_001, 5])
    got = SM.bootstrap_ci(v, w, lo, hi, 0.05, 100, smoothed=False)
    exp_lo, exp_hi = OS.ci_table(v, w, lo, hi, 0.05, 100, False)
    np.testing.assert_allclose(got[0], exp_lo, rtol=0, atol=TOL, equal_nan=True)
    np.testing.assert_allclose(got[1], exp_hi, rtol=0, atol=TOL, equal_nan=True)
    assert np.isnan(got[0][4]) and np.isnan(got[1][4])


def test_device_pointer_entry_and_edges(SM):
    import torch

    from cnvkit_b200 import _dev

    rng = np.random.default_rng(3)
    v = rng.normal(0, 0.3, 5000)
    w = rng.uniform(0.1, 1, 5000)
    lo = np.array([0, 10, 11, 11, 2000])
    hi = np.array([10, 11, 11, 2000, 5000])
    host = SM.bootstrap_ci(v, w, lo, hi, 0.05, 100, smoothed=10)
    dv, dw = _dev.up(v), _dev.up(w)
    dev = SM.bootstrap_ci(None, None, lo, hi, 0.05, 100, smoothed=10, device_ptrs=(dv.data_ptr(), dw.data_ptr()),
                          stream=_dev.stream())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(dev[0].cpu().numpy(), host[0])
    np.testing.assert_array_equal(dev[1].cpu().numpy(), host[1])
    assert host[0][1] == v[10] and host[1][1] == v[10]            # one bin: both bounds are its value
    assert np.isnan(host[0][2]) and np.isnan(host[1][2])          # no bins
    empty = SM.bootstrap_ci(v, w, [], [], 0.05)
    assert len(empty[0]) == 0
    with pytest.raises(ValueError):
        SM.bootstrap_ci(v, w, [0], [10], alpha=0.0)
    from cnvkit_b200._lib import CnvkError

    with pytest.raises(CnvkError):
        SM.bootstrap_ci(v, w, [0], [6000], 0.05)                  # range outside the columns

"""GPU: direct checks of the kernels rewritten in round 2 for the exome path, at the sizes where their code paths
switch -- the register-resident segmented median (<= 16 Ki elements per segment) against the memory-walking one,
the cluster biweight at apply_weights' size, the upper-tail quantile mask with other quantiles and factors."""
import numpy as np
import pytest
import torch

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def segmented_median(x, use, offsets):
    from cnvkit_b200 import _dev, _lib

    d_x = _dev.up(np.ascontiguousarray(x, dtype=np.float64))
    d_use = None if use is None else _dev.up(np.ascontiguousarray(use, dtype=np.uint8))
    d_off = _dev.up(np.ascontiguousarray(offsets, dtype=np.int64))
    nseg = len(offsets) - 1
    d_med, d_cnt = _dev.empty(nseg, torch.float64), _dev.empty(nseg, torch.int32)
    _lib.call("cnvk_segmented_median_dev", _dev.ptr(d_x), _dev.ptr(d_use), _dev.ptr(d_off), nseg, _dev.ptr(d_med),
              _dev.ptr(d_cnt), _dev.stream())
    return d_med.cpu().numpy(), d_cnt.cpu().numpy()


def test_segmented_median_both_paths_and_their_boundary():
    """pandas Series.median semantics (cnary.py:253-313, 688-692): NaN skipped, even counts average the two middle
    ELEMENTS -- so the result is bit-exact, whichever path a segment takes."""
    rng = np.random.default_rng(5)
    sizes = [0, 1, 2, 3, 31, 32, 33, 1023, 1024, 1025, 16383, 16384, 16385, 16386, 40001, 7, 0, 5000]
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    n = int(offsets[-1])
    x = rng.normal(0, 1, n)
    x[rng.random(n) < 0.02] = np.nan
    x[rng.random(n) < 0.05] = 0.25                          # ties around typical medians
    x[rng.random(n) < 0.01] = -0.0
    use = (rng.random(n) < 0.8).astype(np.uint8)
    for mask in (None, use):
        med, cnt = segmented_median(x, mask, offsets)
        for s, (a, b) in enumerate(zip(offsets[:-1], offsets[1:])):
            v = x[a:b] if mask is None else x[a:b][mask[a:b] != 0]
            v = v[~np.isnan(v)]
            assert cnt[s] == len(v)
            if len(v) == 0:
                assert np.isnan(med[s])
            else:
                assert med[s] == np.median(v), (s, sizes[s], mask is None)


def test_segmented_median_all_equal_and_all_nan_segments():
    x = np.r_[np.full(300, 1.5), np.full(20000, -2.0), np.full(10, np.nan), [3.0, np.nan, 1.0, 2.0]]
    offsets = np.array([0, 300, 20300, 20310, 20314])
    med, cnt = segmented_median(x, None, offsets)
    assert med[0] == 1.5 and med[1] == -2.0 and np.isnan(med[2]) and med[3] == 2.0
    assert list(cnt) == [300, 20000, 0, 3]


@pytest.mark.parametrize("n", [2, 3, 100, 4097, 44281, 132893])
def test_biweight_of_a_long_array_vs_oracle(n):
    """apply_weights' residual spread (fix.py:466-503): location and midvariance of 10^4-10^5 values; reductions are
    in a fixed order on the device, numpy's pairwise order is not replayed -> 1e-12 relative."""
    from cnvkit_b200 import descriptives as D

    rng = np.random.default_rng(n)
    x = rng.standard_t(3, n) * 0.2 + 0.05
    loc = O.biweight_location(x)
    np.testing.assert_allclose(D.biweight_location(x), loc, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(D.biweight_midvariance(x), O.biweight_midvariance(x), rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(D.biweight_midvariance(x, initial=0.0), O.biweight_midvariance(x, initial=0.0), rtol=1e-12,
                               atol=1e-15)
    # repeated calls give identical bits (rank-ordered cluster reduction)
    assert D.biweight_midvariance(x) == D.biweight_midvariance(x)


@pytest.mark.parametrize("factor", [1, 3, 10])
def test_outlier_mask_quantile_paths(factor):
    """rolling 0.95-quantile of 51 distances (smoothing.py:439-463): the device keeps the four largest in registers;
    arms around the 50-bin threshold, spikes, ties."""
    from cnvkit_b200 import segmentation as S

    rng = np.random.default_rng(factor)
    sizes = [10, 50, 51, 52, 100, 511, 512, 513, 3000]
    offsets = np.concatenate([[0], np.cumsum(sizes)])
    x = rng.normal(0, 0.2, int(offsets[-1]))
    x[rng.random(len(x)) < 0.03] += rng.choice([-3.0, 3.0])
    x = np.round(x, 2)                                       # many equal distances
    got = S.drop_outliers_mask(x, offsets, 50, factor)
    want = np.concatenate([O.outlier_mask(x[a:b], factor=factor) for a, b in zip(offsets[:-1], offsets[1:])])
    np.testing.assert_array_equal(got, want)

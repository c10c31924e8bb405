"""GPU parity: csrc/rolling.cu (through the C ABI) vs the reference's golden vectors and the
numpy oracle.  Bit-exact (order statistics select input elements)."""
import numpy as np
import pytest

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from cnvkit_b200 import smoothing

    return smoothing


def test_rolling_median_golden(S, golden):
    g = golden("smoothing")
    for i, (n, width) in enumerate(g["rm_cases"]):
        width = int(width) if width >= 1 else float(width)
        np.testing.assert_array_equal(S.rolling_median(g[f"rm{i}_x"], width), g[f"rm{i}_y"])


def test_rolling_median_nan_golden(S, golden):
    g = golden("smoothing")
    np.testing.assert_array_equal(S.rolling_median(g["rm_nan_x"], 0.05), g["rm_nan_y"])


def test_rolling_quantile_golden(S, golden):
    g = golden("smoothing")
    for i in range(4):
        np.testing.assert_array_equal(S.rolling_quantile(g[f"rq{i}_x"], 50, 0.95), g[f"rq{i}_y"])


@pytest.mark.parametrize("n,width", [(2, 0.9), (3, 3), (64, 0.5), (1025, 0.01), (5000, 0.3), (20000, 0.01)])
def test_rolling_median_vs_oracle(S, n, width):
    rng = np.random.default_rng(n)
    x = rng.normal(0, 1, n)
    x[:: max(2, n // 7)] = 0.0  # ties incl. signed zero handling
    np.testing.assert_array_equal(S.rolling_median(x, width), O.rolling_median(x, width))


def test_rolling_median_short_input_unchanged(S):
    # cnvlib/smoothing.py:79-82 (test_properties.py:391-438)
    assert S.rolling_median(np.array([]), 0.5).shape == (0,)
    np.testing.assert_array_equal(S.rolling_median(np.array([1.5]), 0.5), [1.5])


def test_rolling_median_full_size_properties(S):
    """C3-sized input (3M bins, window 30 001): size-independent properties.

    median of a constant is the constant; median commutes with monotone maps; output lies
    within [min, max]; a window-wide plateau is reproduced exactly.
    """
    n = 3_000_000
    rng = np.random.default_rng(3)
    x = rng.normal(0, 0.25, n)
    width = max(0.01, n**-0.5)
    y = S.rolling_median(x, width)
    assert y.shape == (n,) and np.isfinite(y).all()
    assert y.min() >= x.min() and y.max() <= x.max()
    np.testing.assert_array_equal(S.rolling_median(2.0 * x + 1.0, width), 2.0 * y + 1.0)
    # spot-check 64 positions against direct window medians
    wing = O.width2wing(width, n)
    padded = O.pad_mirror(x, wing)
    for i in rng.integers(0, n, 64):
        assert y[i] == np.median(padded[i : i + 2 * wing + 1])


def test_bad_wing_raises(S):
    from cnvkit_b200 import _lib

    x = np.zeros(10)
    out = np.zeros(10)
    with pytest.raises(_lib.CnvkError):
        _lib.call("cnvk_rolling_median", _lib.ptr(x), 10, 10, _lib.ptr(out))

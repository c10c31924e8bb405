"""GPU: coverage finalisation (cnvlib/coverage.py:637-643, :403-409) against the numpy restatement of those lines.
depth is one IEEE division (bit-exact); log2 goes through the device's log2, which is within 1 ulp of glibc's, and
is identical once written with the %.6g the .cnn files carry."""
import numpy as np
import pandas as pd
import pytest

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def test_depth_and_log2():
    from cnvkit_b200 import coverage

    rng = np.random.default_rng(2)
    n = 200_000
    start = rng.integers(0, 10**8, n)
    end = start + rng.integers(-5, 2000, n)          # some zero-width / reversed bins (user-supplied BEDs)
    base = rng.integers(0, 10**6, n).astype(float)
    base[rng.random(n) < 0.1] = 0.0
    depth, log2 = coverage.depth_log2(base, start, end)
    d_want, l_want = O.coverage_depth_log2(base, start, end)
    np.testing.assert_array_equal(depth, d_want)
    assert (log2[(end - start <= 0) | (base == 0)] == -20.0).all()
    np.testing.assert_allclose(log2, l_want, rtol=4e-16, atol=0)
    fmt = lambda a: np.char.mod("%.6g", a)
    np.testing.assert_array_equal(fmt(log2), fmt(l_want))
    tab = pd.DataFrame({"chromosome": "chr1", "start": start[:10], "end": end[:10], "basecount": base[:10]})
    out = coverage.finalize_basecounts(tab)
    assert (out["gene"] == "-").all() and list(out.columns[-2:]) == ["depth", "log2"]

"""GPU: DNAcopy smooth.CNA (--smooth-cbs, cnvlib/segmentation/cbs.py:51-58) on the device against the CPU
restatement (oracle/cbs_oracle.c cbso_smooth_cna; PARITY UNPINNED like CBS itself -- R is absent), plus the
behaviour the option documents: isolated outliers are pulled to the local median +- 2 trimmed SD, runs are kept."""
import numpy as np
import pytest

from oracle import cbs as OC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    from cnvkit_b200.segmentation import cbs

    return cbs


def test_smooth_cna_equals_oracle(G):
    rng = np.random.default_rng(8)
    sizes = [1, 2, 3, 5, 21, 22, 500, 4097, 70000]
    xs = []
    for n in sizes:
        x = rng.normal(0, 0.2, n)
        if n > 20:
            x[rng.choice(n, size=max(1, n // 50), replace=False)] += rng.choice([-3.0, 3.0, 1.5], size=max(1, n // 50))
            x[n // 2: n // 2 + 12] += 2.0             # a run: not an outlier
        xs.append(np.round(x, 4))                     # heavy ties in |diff| exercise the radix select
    x = np.concatenate(xs)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    want, sd = OC.smooth_cna(x, off)
    got = G.smooth_cna(x, off)
    changed = want != x
    assert changed.sum() > 100
    np.testing.assert_array_equal(got[~changed], x[~changed])       # untouched bins are copied bit for bit
    np.testing.assert_array_equal(got != x, changed)                # the same bins are moved
    np.testing.assert_allclose(got[changed], want[changed], rtol=1e-14, atol=0)


def test_isolated_outliers_are_pulled_in_runs_are_kept(G):
    rng = np.random.default_rng(0)
    x = rng.normal(0, 0.2, 500)
    x[100] += 3
    x[101] -= 3
    x[0] += 4
    x[499] -= 5
    x[250:260] += 3
    out = G.smooth_cna(x, [0, 500])
    assert np.flatnonzero(out != x).tolist() == [0, 100, 101, 499]
    d = np.sort(np.abs(np.diff(x)))
    nk = int(np.rint(0.95 * 499))
    sd = np.sqrt(1.3177980408013314 * np.sum(d[:nk] ** 2) / (2 * nk))      # inflfact(0.025) = 1.3178
    assert out[100] == pytest.approx(np.median(x[90:111]) + 2 * sd, rel=1e-12)
    assert out[499] == pytest.approx(np.median(x[489:500]) - 2 * sd, rel=1e-12)


def test_do_segmentation_smooth_cbs(golden):
    """smooth_cbs=True no longer falls back to cnvlib: breakpoints come from the smoothed values, segment means
    from the original ones (cbs.py:64-73), and single-bin arms are left alone (cbs.py:56)."""
    from conftest import golden_table
    from cnvkit_b200 import segmentation as S

    cnr = golden_table(golden("haar"), "cnr")[:4000]
    plain = S.do_segmentation(cnr, "cbs")
    smooth = S.do_segmentation(cnr, "cbs", smooth_cbs=True)
    assert len(smooth) > 0 and len(smooth) <= len(plain) + 2
    d = cnr.data
    for _, row in smooth.data.iloc[:5].iterrows():
        assert np.isfinite(row["log2"])
    one = S.do_segmentation(cnr[:1], "cbs", smooth_cbs=True)
    assert len(one) == 1

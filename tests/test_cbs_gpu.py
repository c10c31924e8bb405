"""GPU parity: csrc/cbs.cu through the C ABI vs the CPU oracle (oracle/cbs_oracle.c).

The oracle is a *restatement* of DNAcopy (parity unpinned -- see its header); what is asserted
here is that the CUDA path reproduces it exactly: identical breakpoints, segment means within
1e-9 relative, on the reference's own .cnr fixture and on synthetic data with planted
change-points, through every decision path (shortcut split, tail approximation, hybrid and
exact permutation tests, ternary-split flank tests)."""
import numpy as np
import pytest

from oracle import cbs as OC

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    from cnvkit_b200.segmentation import cbs

    return cbs


def q6(a):
    """%.6g quantisation applied at the DNAcopy boundary (segmentation/__init__.py:299-301)."""
    return np.array([float("%.6g" % v) for v in a])


def planted(rng, n, nseg, sigma=0.2):
    x = rng.normal(0, sigma, n)
    cuts = np.sort(rng.choice(np.arange(5, n - 5), size=min(nseg, max(0, (n - 10) // 12)), replace=False))
    for a, b in zip(cuts[::2], cuts[1::2]):
        x[a:b] += rng.choice([-1.0, -0.4, 0.3, 0.58, 1.0])
    return x


def compare(G, x, w, off, alpha, **kw):
    # large inputs use the oracle's block-pruned search (equal to its exhaustive scan by
    # tests/test_oracle_pins.py::test_cbs_oracle_pruned_search_equals_exhaustive); small ones the exhaustive scan
    P = OC.make_params(alpha=alpha, prune=len(x) > 20000, **kw)
    o_arm, o_lo, o_hi, o_mean, o_st = OC.segment(x, w, off, P, nthreads=16)
    g_arm, g_lo, g_hi, g_mean, g_st = G.segment_arms(x, w, off, alpha=alpha, **kw)
    np.testing.assert_array_equal(g_arm, o_arm)
    np.testing.assert_array_equal(g_lo, o_lo)
    np.testing.assert_array_equal(g_hi, o_hi)
    np.testing.assert_allclose(g_mean, o_mean, rtol=1e-9, atol=0)
    return o_st, g_st


def test_getbdry_matches_oracle(G):
    for max_ones in (1, 2, 3, 11):
        np.testing.assert_array_equal(G.getbdry(0.05, 10000, max_ones), OC.getbdry(0.05, 10000, max_ones))


def test_planted_changepoints_mixed_arm_sizes(G):
    rng = np.random.default_rng(11)
    sizes = [1, 2, 3, 4, 5, 7, 40, 150, 199, 200, 201, 257, 600, 1023, 1024, 1025, 2500, 5000]
    xs, ws = [], []
    for n in sizes:
        xs.append(planted(rng, n, 6) if n > 30 else rng.normal(0, 0.2, n))
        ws.append(rng.uniform(0.5, 1.0, n))
    x, w = q6(np.concatenate(xs)), q6(np.concatenate(ws))
    off = np.concatenate([[0], np.cumsum(sizes)])
    o_st, g_st = compare(G, x, w, off, 1e-4)
    assert g_st["nodes"] == o_st["nodes"]
    # prune on/off give the same table
    a = G.segment_arms(x, w, off, alpha=1e-4, prune=True)
    b = G.segment_arms(x, w, off, alpha=1e-4, prune=False)
    for u, v in zip(a[:4], b[:4]):
        np.testing.assert_array_equal(u, v)
    assert a[4]["obs_subtiles_scanned"] <= b[4]["obs_subtiles_scanned"]


@pytest.mark.parametrize("alpha", [0.05, 0.01])
def test_null_and_weak_signals_exercise_permutation_paths(G, alpha):
    """Loose alpha + weak steps: many nodes land in the hybrid / exact permutation tests and
    in the ternary-split flank tests; decisions must still agree one for one."""
    rng = np.random.default_rng(int(alpha * 1000))
    sizes = [60, 120, 180, 199, 230, 400, 800, 1500] * 3
    xs, ws = [], []
    for n in sizes:
        x = rng.normal(0, 1.0, n)
        if n % 3 == 0:
            a = n // 3
            x[a : a + max(6, n // 10)] += rng.choice([0.8, 1.2, -1.0])
        xs.append(x)
        ws.append(rng.uniform(0.2, 1.0, n))
    x, w = q6(np.concatenate(xs)), q6(np.concatenate(ws))
    off = np.concatenate([[0], np.cumsum(sizes)])
    o_st, g_st = compare(G, x, w, off, alpha)
    assert o_st["perm_tests"] > 0 and g_st["perm_nodes"] == o_st["perm_tests"]


def test_reference_fixture_p2_20_1(G, golden):
    """C1: the reference's own exome .cnr (test/formats/p2-20_1.cnr), arms from by_arm()."""
    g = golden("haar")
    x, w = g["cnr__log2"], g["cnr__weight"]
    keep = w > 0
    assert keep.all()
    compare(G, x, w, g["arm_offsets"], 1e-4)


def test_unit_weights_and_flat_arms(G):
    rng = np.random.default_rng(5)
    x = np.concatenate([np.full(300, 0.25), planted(rng, 700, 4), np.zeros(50)])
    w = np.ones_like(x)
    off = np.array([0, 300, 1000, 1050])
    compare(G, q6(x), w, off, 1e-4)
    arm, lo, hi, mean, _ = G.segment_arms(q6(x), w, off, alpha=1e-4)
    assert (arm == 0).sum() == 1 and (arm == 2).sum() == 1


def test_empty_and_single_bin(G):
    arm, lo, hi, mean, st = G.segment_arms(np.zeros(0), np.zeros(0), np.array([0]), alpha=1e-4)
    assert len(arm) == 0
    arm, lo, hi, mean, st = G.segment_arms(np.array([0.7]), np.array([0.5]), np.array([0, 1]), alpha=1e-4)
    assert list(lo) == [0] and list(hi) == [1] and mean[0] == pytest.approx(0.7)  # test_r.py:215-231


def test_planted_step_two_segments(G):
    """test/test_r.py:234-291: step at bin 60 of 120, sigma 0.02 -> exactly 2 segments."""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(0.0, 0.02, 60), rng.normal(-1.0, 0.02, 60)])
    arm, lo, hi, mean, _ = G.segment_arms(q6(x), np.ones(120), np.array([0, 120]), alpha=1e-4)
    assert list(lo) == [0, 60] and list(hi) == [60, 120]


def test_c3_scale_properties(G):
    """One long WGS-like arm (125k bins): planted breakpoints are recovered, every segment mean
    equals the weighted mean of its bins, the table tiles the arm, prune on/off agree."""
    rng = np.random.default_rng(33)
    n = 125_000
    x = rng.normal(0, 0.25, n)
    truth = [20_000, 20_400, 61_000, 90_000, 90_050]
    lvl = [0.0, 0.58, 0.0, -1.0, 0.0, 1.0]
    edges = [0] + truth + [n]
    for k in range(len(edges) - 1):
        x[edges[k] : edges[k + 1]] += lvl[k]
    w = rng.uniform(0.5, 1.0, n)
    x, w = np.round(x, 5), np.round(w, 5)
    off = np.array([0, n])
    arm, lo, hi, mean, st = G.segment_arms(x, w, off, alpha=1e-4)
    assert lo[0] == 0 and hi[-1] == n and (lo[1:] == hi[:-1]).all()
    for t in truth:
        assert np.min(np.abs(hi - t)) <= 3, (t, hi)
    for a, b, m in zip(lo, hi, mean):
        assert m == pytest.approx(np.average(x[a:b], weights=w[a:b]), rel=1e-9)
    arm2, lo2, hi2, mean2, st2 = G.segment_arms(x, w, off, alpha=1e-4, prune=False)
    np.testing.assert_array_equal(lo, lo2)
    np.testing.assert_array_equal(hi, hi2)
    assert st["obs_subtiles_scanned"] < st2["obs_subtiles_scanned"]


def test_tailp_bracket_gives_the_series_decisions(G):
    """The closed-form bracket of tailp (cbs_tailp_bounds_kernel) must never change a decision:
    same tables with the bracket on (default) and off (always sum the nu series), on weak signals
    chosen so that many nodes reach the tail approximation, for several alpha (nrejc steps)."""
    rng = np.random.default_rng(2024)
    sizes = [260, 400, 900, 1500, 3000, 8000, 20000]
    for alpha in (1e-4, 1e-3, 1e-2, 0.05):
        xs, ws = [], []
        for n in sizes:
            x = rng.normal(0, 0.2, n)
            for _ in range(4):
                a = int(rng.integers(10, n - 40))
                x[a:a + int(rng.integers(8, 30))] += rng.choice([-1, 1]) * rng.uniform(0.1, 0.35)
            xs.append(x)
            ws.append(rng.uniform(0.5, 1.0, n))
        x, w = q6(np.concatenate(xs)), q6(np.concatenate(ws))
        off = np.concatenate([[0], np.cumsum(sizes)])
        a = G.segment_arms(x, w, off, alpha=alpha)
        b = G.segment_arms(x, w, off, alpha=alpha, tailp_bracket=False)
        for u, v in zip(a[:4], b[:4]):
            np.testing.assert_array_equal(u, v)
        compare(G, x, w, off, alpha)


def test_full_c3_sample_equals_oracle(G):
    """BASELINE's full size: the 3 M-bin synthetic WGS sample (48 arms) segmented on the GPU gives the
    table of the CPU oracle (block-pruned search; equal to its exhaustive scan by
    tests/test_oracle_pins.py) -- identical breakpoints, means to 1e-9 -- and recovers the planted
    change-points."""
    from cnvkit_b200 import synth

    d = synth.wgs_sample()
    g = G.segment_arms(d["log2"], d["weight"], d["arm_offsets"], alpha=1e-4)
    o = OC.segment(d["log2"], d["weight"], d["arm_offsets"], OC.make_params(alpha=1e-4, prune=True), nthreads=16)
    for a, b in zip(g[:3], o[:3]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(g[3], o[3], rtol=1e-9, atol=0)
    cuts = set(g[1].tolist()) | set(g[2].tolist())
    found = sum(1 for lo, hi, step in d["truth"] if (hi - lo) >= 50 and abs(step) >= 0.5
                and min(abs(c - lo) for c in cuts) <= 5 and min(abs(c - hi) for c in cuts) <= 5)
    big = sum(1 for lo, hi, step in d["truth"] if (hi - lo) >= 50 and abs(step) >= 0.5)
    assert found >= 0.95 * big, (found, big)


def test_many_small_arms_in_one_batch(G):
    """A cohort-style batch: thousands of arms of 1..600 bins in one call (every kernel's node / tile
    indexing, empty arms included) against the oracle."""
    rng = np.random.default_rng(77)
    sizes = rng.integers(0, 600, size=4000)
    sizes[rng.integers(0, len(sizes), 200)] = 0
    xs, ws = [], []
    for n in sizes:
        x = rng.normal(0, 0.2, n)
        if n > 60 and rng.random() < 0.5:
            a = int(rng.integers(5, n - 30))
            x[a:a + int(rng.integers(10, 25))] += rng.choice([-0.8, 0.6, 1.0])
        xs.append(x)
        ws.append(rng.uniform(0.5, 1.0, n))
    x, w = q6(np.concatenate(xs)), q6(np.concatenate(ws))
    off = np.concatenate([[0], np.cumsum(sizes)])
    P = OC.make_params(alpha=1e-4, prune=True)
    o = OC.segment(x, w, off, P, nthreads=16)
    g = G.segment_arms(x, w, off, alpha=1e-4)
    for a, b in zip(g[:3], o[:3]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(g[3], o[3], rtol=1e-9, atol=0)
    assert len(g[0]) > int((sizes > 0).sum())

"""Statistical pins of the CBS restatement (CPU oracle; the CUDA path runs the same checks in
tests/test_cbs_stats_gpu.py).

The oracle restates DNAcopy from SURVEY.md Appendix A and nothing in /root/reference pins its numbers
(R is absent -- "parity unpinned").  What CAN be pinned is checked here:
  * the counter-based permutations are bijections with uniform one- and two-point marginals;
  * hybrid / exact / flank permutation p-values agree with a numpy true-shuffle evaluation of the same
    statistic within Monte-Carlo error (north_star's p-value clause) on 52 seeded nodes;
  * null calibration of the whole decision chain (exact path n <= 200, hybrid tail approximation +
    short-arc permutations n > 200): the false-split rate on pure-noise arms is at most alpha (binomial
    band) -- the guard available against a mis-recalled constant;
  * the structural assertions of the reference's test/test_r.py:215-291.
"""
import numpy as np
import pytest

import cbs_stat_helpers as H
from oracle import cbs as OC

NPERM = 10000


# ------------------------------------------------------------------ permutations
@pytest.mark.parametrize("n", [2, 3, 5, 37, 199, 200, 256, 257, 1000, 1024, 1025, 5000, 70000])
def test_prp_is_a_bijection(n):
    m = 40 if n <= 5000 else 3
    P = OC.prp_fill(0xA5EED, 11, 11 + n, 0, 0, m, n)
    np.testing.assert_array_equal(np.sort(P, axis=1), np.tile(np.arange(n), (m, 1)))
    # different (node, test, permutation) keys give different permutations
    if n >= 37:
        Q = OC.prp_fill(0xA5EED, 11, 11 + n, 1, 0, 2, n)
        assert (P[0] != P[1]).any() and (P[0] != Q[0]).any()


def _chi_z(counts, expected):
    counts = np.asarray(counts, dtype=float).ravel()
    expected = np.broadcast_to(np.asarray(expected, dtype=float), counts.shape).ravel()
    dof = counts.size - 1
    return (((counts - expected) ** 2 / expected).sum() - dof) / np.sqrt(2.0 * dof)


@pytest.mark.parametrize("n", [50, 200, 256, 853, 4096])
def test_prp_one_and_two_point_marginals_are_uniform(n):
    """Positions that share a Feistel half are the pairs a too-short network correlates (4 rounds on an
    8-bit domain fail this test with z > 100): chi-square of pi(t2) - pi(t1) mod n over 2e5 permutations."""
    M = 200000
    bits = 8
    while (1 << bits) < n:
        bits += 1
    b = bits - bits // 2
    pairs = [(0, 1), (n // 2, n // 2 + 1), (0, n - 1), (3, 7)]
    if (1 << b) < n:
        pairs += [(0, 1 << b), (1, 1 + (1 << b))]
    if (3 << b) < n:
        pairs += [(1 << b, 3 << b)]
    pos = sorted({t for pr in pairs for t in pr} | {n // 3})
    col = {t: k for k, t in enumerate(pos)}
    P = OC.prp_at(0xA5EED, 0, n, 0, 0, M, n, pos)
    for t in (0, n // 3, n - 1):
        assert abs(_chi_z(np.bincount(P[:, col[t]], minlength=n), M / n)) < 5.0
    for t1, t2 in pairs:
        d = (P[:, col[t2]] - P[:, col[t1]]) % n
        z = _chi_z(np.bincount(d, minlength=n)[1:], M / (n - 1))
        assert abs(z) < 5.0, (n, t1, t2, z)
    # successive permutation indices are unrelated: value at a fixed position, perm p vs p+1
    if n <= 256:
        joint = np.bincount(P[:-1, 0] * n + P[1:, 0], minlength=n * n)
        assert abs(_chi_z(joint, (M - 1) / (n * n))) < 5.0


# ------------------------------------------------------------------ p-values vs true shuffles
def _check_z(zs):
    zs = np.asarray(zs)
    assert np.abs(zs).max() < 4.5, zs
    assert (np.abs(zs) <= 3.0).mean() >= 0.97, zs
    rms = np.sqrt(np.mean(zs**2))
    assert 0.6 < rms < 1.4, rms


@pytest.mark.parametrize("kind,seed", [("hybrid", 101), ("exact", 202)])
def test_node_and_flank_pvalues_agree_with_true_shuffles(kind, seed):
    """26 nodes per kind: all 10 000 counter-based permutations (no early stopping) against 2 000 numpy
    shuffles of the same statistic; flank p-values likewise.  |z| within Monte-Carlo error."""
    P = OC.make_params(alpha=1e-4)
    zs, zf = [], []
    for k, (x, w) in enumerate(H.stat_nodes(kind, 26, seed)):
        n = len(x)
        lo = 17 * k                                   # the node sits inside a longer arm: keys use (lo, hi)
        xa = np.concatenate([np.zeros(lo), x])
        wa = np.concatenate([np.ones(lo), w])
        d = OC.node_diag(xa, wa, lo, lo + n, P, NPERM)
        assert d["hybrid"] == (kind == "hybrid")
        M = H.NodeModel(x, w)
        t2, _, _ = M.observed()
        assert d["ostat1"] == pytest.approx(np.sqrt(t2), rel=1e-9)
        rng = np.random.default_rng(seed * 1000 + k)
        m = 2000
        p2 = M.node_pvalue_true_shuffle(d["ostat"], d["hybrid"], m, rng)
        zs.append(H.z_score(d["nrej"], NPERM, p2, m))
        i, j = d["arc"]
        if i != 0 and j != n and min(i, j - i, n - j) > 1:
            for (a0, n1, n2), pf in zip(((0, i, j - i), (i, j - i, n - j)), d["flank_p"]):
                p2f = H.flank_pvalue_true_shuffle(M.xc[a0 : a0 + n1 + n2], w[a0 : a0 + n1 + n2], n1, n2, 20000, rng)
                zf.append(H.z_score(round(pf * NPERM), NPERM, p2f, 20000))
    _check_z(zs)
    assert len(zf) >= 20
    _check_z(zf)


def test_tailp_matches_direct_simulation():
    """The hybrid tail approximation (arcs longer than kmax) is an analytic formula; check it against a
    direct Monte-Carlo estimate of P(max over arcs wider than kmax of T >= b) on white noise, n = 300.
    Siegmund's approximation is accurate to a few tens of percent in this range -- enough to catch a
    mis-recalled constant (the 0.0997356 = 1/(4 sqrt(2 pi)) prefactor, delta, the grid)."""
    n, kmax, reps = 300, 25, 4000
    rng = np.random.default_rng(5)
    X = rng.normal(size=(reps, n))
    X -= X.mean(axis=1, keepdims=True)
    S = np.concatenate([np.zeros((reps, 1)), np.cumsum(X, axis=1)], axis=1)
    tss = (X**2).sum(axis=1)
    best = np.zeros(reps)
    for k in range(kmax + 1, n // 2 + 1):               # arc and complement give the same statistic
        d = S[:, k:] - S[:, :-k]
        q = (d * d).max(axis=1) / (k * (n - k) / n)
        best = np.maximum(best, q)
    t = np.sqrt(best / ((tss - best) / (n - 2)))
    for b in (3.0, 3.5, 4.0):
        sim = (t >= b).mean()
        ana = OC.tailp(b, (kmax + 1) / n, n)
        assert 0.5 * sim - 0.003 <= ana <= 2.0 * sim + 0.003, (b, sim, ana)


# ------------------------------------------------------------------ null calibration
@pytest.mark.parametrize("n,count", [(50, 1000), (150, 1000), (400, 1000), (5000, 150)])
@pytest.mark.parametrize("weighted", [True, False])
def test_false_split_rate_on_noise_is_at_most_alpha(n, count, weighted):
    alpha = 1e-2
    x, w, off = H.noise_arms(n, count, seed=n * 7 + int(weighted), weighted=weighted)
    P = OC.make_params(alpha=alpha, prune=n > 1000)
    arm = OC.segment(x, w, off, P, nthreads=0)[0]
    k = H.false_split_arms(arm, count)
    lo, hi = H.binom_band(count, alpha)
    assert k <= hi, (k, count, alpha)
    # the chain is conservative (sequential boundary, 0.99999 factor, tail bound) but not dead: over
    # the four sizes together a calibrated test must split some arms
    test_false_split_rate_on_noise_is_at_most_alpha.total = getattr(
        test_false_split_rate_on_noise_is_at_most_alpha, "total", 0) + k


def test_false_split_rate_is_not_degenerate():
    total = getattr(test_false_split_rate_on_noise_is_at_most_alpha, "total", None)
    if total is None:
        pytest.skip("calibration cases not run in this session")
    # 6300 arms at alpha = 0.01: expectation <= 63; a dead test (never splitting) would give 0
    assert 8 <= total <= 100, total


# ------------------------------------------------------------------ reference test_r.py structure
def test_planted_step_gives_two_segments_like_test_r():
    """test/test_r.py:234-291: 120 bins, step at bin 60, sigma 0.02, default_rng(0) -> exactly 2 segments."""
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(0.0, 0.02, 60), rng.normal(1.0, 0.02, 60)])
    x = H.q6(x)
    w = np.ones(120)
    arm, lo, hi, mean, _ = OC.segment(x, w, np.array([0, 120]), OC.make_params(alpha=1e-4), nthreads=1)
    assert lo.tolist() == [0, 60] and hi.tolist() == [60, 120]
    assert mean[0] == pytest.approx(0.0, abs=0.02) and mean[1] == pytest.approx(1.0, abs=0.02)


def test_single_bin_and_empty_arms_like_test_r():
    """test/test_r.py:177-231: empty arms vanish, a single bin is one segment with that bin's log2."""
    x = np.array([0.37])
    arm, lo, hi, mean, _ = OC.segment(x, np.ones(1), np.array([0, 0, 1, 1]), OC.make_params(), nthreads=1)
    assert arm.tolist() == [1] and lo.tolist() == [0] and hi.tolist() == [1] and mean[0] == 0.37

"""Statistical pins of the CUDA CBS path (twin of tests/test_cbs_stats.py, which pins the CPU oracle).

north_star: "permutation p-values agree within Monte Carlo standard error".  R / DNAcopy cannot run here,
so the yardstick is a numpy true-shuffle evaluation of the same statistics (tests/cbs_stat_helpers.py).
Everything goes through the C ABI (cnvk_cbs_node_diag, cnvk_cbs_segment)."""
import numpy as np
import pytest

import cbs_stat_helpers as H
from oracle import cbs as OC

pytestmark = pytest.mark.gpu
NPERM = 10000


@pytest.fixture(scope="module")
def G():
    from cnvkit_b200.segmentation import cbs

    return cbs


def _batch(nodes):
    x = np.concatenate([a for a, _ in nodes])
    w = np.concatenate([b for _, b in nodes])
    off = np.concatenate([[0], np.cumsum([len(a) for a, _ in nodes])]).astype(np.int64)
    return x, w, off


def _check_z(zs):
    zs = np.asarray(zs)
    assert np.abs(zs).max() < 4.5, zs
    assert (np.abs(zs) <= 3.0).mean() >= 0.97, zs
    rms = np.sqrt(np.mean(zs**2))
    assert 0.6 < rms < 1.4, rms


@pytest.mark.parametrize("kind,seed", [("hybrid", 101), ("exact", 202)])
def test_node_diag_equals_oracle_and_true_shuffles(G, kind, seed):
    """26 nodes per kind in ONE device call: observed statistic, arc, tail probability and the exceedance
    count over all 10 000 permutations equal the oracle's (same counter-based permutations), and both agree
    with 2 000 numpy shuffles within Monte-Carlo error; flank p-values likewise."""
    nodes = H.stat_nodes(kind, 26, seed)
    x, w, off = _batch(nodes)
    got = G.node_diag(x, w, off, alpha=1e-4)
    P = OC.make_params(alpha=1e-4)
    zs, zf = [], []
    for k, ((xn, wn), g) in enumerate(zip(nodes, got)):
        n = len(xn)
        o = OC.node_diag(xn, wn, 0, n, P, NPERM)        # every node is its own arm in the device call: keys (0, n)
        assert g["hybrid"] == o["hybrid"] == (kind == "hybrid")
        assert g["arc"] == o["arc"]
        assert g["ostat1"] == o["ostat1"] and g["ostat"] == o["ostat"]
        assert g["nrej"] == o["nrej"], (k, g["nrej"], o["nrej"])
        if kind == "hybrid":
            assert g["delta"] == o["delta"]
            assert g["tailp"] == pytest.approx(o["tailp"], rel=1e-9)
        # flank sums are reduced in a different (fixed) order on the device: counts may differ on exact ties only
        for a, b in zip(g["flank_p"], o["flank_p"]):
            assert abs(a - b) <= 3.0 / NPERM
        M = H.NodeModel(xn, wn)
        rng = np.random.default_rng(seed * 1000 + k)
        p2 = M.node_pvalue_true_shuffle(g["ostat"], g["hybrid"], 2000, rng)
        zs.append(H.z_score(g["nrej"], NPERM, p2, 2000))
        i, j = g["arc"]
        if i != 0 and j != n and min(i, j - i, n - j) > 1:
            for (a0, n1, n2), pf in zip(((0, i, j - i), (i, j - i, n - j)), g["flank_p"]):
                p2f = H.flank_pvalue_true_shuffle(M.xc[a0 : a0 + n1 + n2], wn[a0 : a0 + n1 + n2], n1, n2, 20000, rng)
                zf.append(H.z_score(round(pf * NPERM), NPERM, p2f, 20000))
    _check_z(zs)
    assert len(zf) >= 20
    _check_z(zf)


CAL = {}


@pytest.mark.parametrize("n", [50, 150, 400, 5000])
@pytest.mark.parametrize("alpha", [1e-2, 1e-3])
@pytest.mark.parametrize("weighted", [True, False])
def test_false_split_rate_on_2000_noise_arms(G, n, alpha, weighted):
    """Null calibration of the whole chain (exact path n <= 200; tail approximation + short-arc permutations
    above): the number of pure-noise arms that get split stays inside the binomial band of alpha."""
    count = 2000
    x, w, off = H.noise_arms(n, count, seed=n * 7 + int(weighted), weighted=weighted)
    arm = G.segment_arms(x, w, off, alpha=alpha)[0]
    k = H.false_split_arms(arm, count)
    lo, hi = H.binom_band(count, alpha)
    assert k <= hi, (k, count, alpha)
    CAL[(n, alpha, weighted)] = k


def test_false_split_rate_is_not_degenerate():
    ks = [v for (n, a, wt), v in CAL.items() if a == 1e-2]
    if len(ks) < 8:
        pytest.skip("calibration cases not run in this session")
    # 16 000 arms at alpha = 0.01: a calibrated (slightly conservative) test splits between 0.2 % and 1 %
    assert 32 <= sum(ks) <= 160, CAL


def cliff_sample():
    """One 120 000-bin arm whose maximal statistic (|t| = 6.81) sits between the tail approximation's reach
    (tailp = 1.9e-5 <= alpha = 1e-4) and the 7.0 shortcut: only the permutation test can accept the split, and
    it needs DNAcopy's full 9 500 clean permutations of 24 n arcs each ('known cliff')."""
    rng = np.random.default_rng(0xC11FF)
    n = 120000
    w = rng.uniform(0.5, 1.0, n)
    x = rng.normal(0, 0.25, n)
    x[45000:] += 0.009
    return H.q6(x), H.q6(w), np.array([0, n], dtype=np.int64)


def test_permutation_cliff_sample_equals_oracle(G):
    x, w, off = cliff_sample()
    g = G.segment_arms(x, w, off, alpha=1e-4)
    assert g[4]["perms_run"] >= 9500
    P = OC.make_params(alpha=1e-4, prune=True)
    d = OC.node_diag(x, w, 0, len(x), P, 0)
    assert 6.5 <= d["ostat1"] < 7.0 and d["tailp"] <= 1e-4
    o = OC.segment(x, w, off, P, nthreads=0)
    assert o[4]["perms_run"] >= 9500
    for a, b in zip(g[:3], o[:3]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(g[3], o[3], rtol=1e-9)
    assert len(g[1]) >= 2          # the planted step was accepted

"""Shared pieces of the statistical CBS tests (tests/test_cbs_stats.py on the CPU oracle,
tests/test_cbs_stats_gpu.py on the CUDA path).

north_star: "permutation p-values agree within Monte Carlo standard error".  R / DNAcopy cannot run
here, so the yardstick is an independent evaluation of the SAME statistics under
TRUE uniform shuffles drawn by numpy: if the counter-based permutations of csrc/cbs_rng.cuh /
oracle/cbs_oracle.c were biased, or a kernel evaluated a different statistic, the two Monte-Carlo
estimates would separate by more than their standard errors.

Statistics restated here in numpy (DNAcopy wfindcpt / hwtmaxp / wtmaxp / wtpermp):
  node test   T^2 = BSS (n-2) / (TSS - BSS), BSS = max over arcs of (sum_arc t)^2 / (c (W - c)) with
              t_i = y_pi(i) sqrt(w_i), y = xc sqrt(w), c = arc weight / sqrt(sum w), W = sqrt(sum w);
              hybrid (n > nmin): circular arcs of min.width..kmax probes; exact: all linear arcs
  flank test  | sum_{short side} y_pi(k) sqrt(w_k) / (weight of the short side) - xbar |
"""
from __future__ import annotations

import numpy as np


def q6(a):
    """%.6g quantisation applied at the DNAcopy boundary (segmentation/__init__.py:299-301)."""
    return np.array([float("%.6g" % v) for v in np.asarray(a, dtype=float)])


def t2_of(bss, tss, n):
    tss = np.where(tss <= bss + 1e-4, bss + 1.0, tss)
    return bss / ((tss - bss) / (n - 2.0))


class NodeModel:
    """Numpy restatement of one node's set-up (changepoints() for weighted data)."""

    def __init__(self, x, w):
        x = np.asarray(x, dtype=np.float64)
        w = np.asarray(w, dtype=np.float64)
        self.n = len(x)
        sw = w.sum()
        self.avg = (x * w).sum() / sw
        self.xc = x - self.avg
        self.w = w
        self.rw = np.sqrt(w)
        self.y = self.xc * self.rw
        self.tss = (w * self.xc**2).sum()
        self.cw = np.concatenate([[0.0], np.cumsum(w)]) / np.sqrt(sw)   # prefix points 0..n
        self.total = self.cw[-1]
        self.wn = w / np.sqrt(sw)

    # ---- observed statistic (all linear arcs, al0 <= j-i <= n-al0)
    def observed(self, al0=2):
        S = np.concatenate([[0.0], np.cumsum(self.xc * self.w)])
        t2, i, j = self._max_all_arcs(S[None, :], al0, want_arg=True)
        return float(t2[0]), int(i), int(j)

    def _max_all_arcs(self, S, al0, want_arg=False):
        n = self.n
        D = S[:, None, :] - S[:, :, None]                 # [m, i, j] = S_j - S_i
        C = self.cw[None, :] - self.cw[:, None]           # [i, j]
        kw = np.arange(n + 1)[None, :] - np.arange(n + 1)[:, None]
        ok = (kw >= al0) & (kw <= n - al0)
        den = np.where(ok, C * (self.total - C), 1.0)
        Q = np.where(ok[None], D * D / den[None], -1.0)
        flat = Q.reshape(Q.shape[0], -1)
        bss = flat.max(axis=1)
        t2 = t2_of(bss, self.tss, n)
        if want_arg:
            k = int(flat[0].argmax())
            return t2, k // (n + 1), k % (n + 1)
        return t2

    # ---- permutation statistics under TRUE shuffles
    def shuffled_terms(self, rng, m):
        """m rows of t_i = y_pi(i) sqrt(w_i) for uniform random permutations pi (numpy Generator.permuted)."""
        Y = rng.permuted(np.tile(self.y, (m, 1)), axis=1)
        return Y * self.rw[None, :]

    def hybrid_stat(self, T, al0=2, kmax=25):
        """max over circular arcs of al0..kmax probes, per row of T -> T^2."""
        n = self.n
        ext = np.concatenate([T, T[:, : kmax]], axis=1)
        P = np.concatenate([np.zeros((T.shape[0], 1)), np.cumsum(ext, axis=1)], axis=1)
        wext = np.concatenate([self.wn, self.wn[:kmax]])
        Pw = np.concatenate([[0.0], np.cumsum(wext)])
        best = np.zeros(T.shape[0])
        for k in range(al0, kmax + 1):
            acc = P[:, k : k + n] - P[:, :n]
            c = Pw[k : k + n] - Pw[:n]
            q = acc * acc / (c * (self.total - c))[None, :]
            best = np.maximum(best, q.max(axis=1))
        return t2_of(best, self.tss, n)

    def exact_stat(self, T, al0=2):
        S = np.concatenate([np.zeros((T.shape[0], 1)), np.cumsum(T, axis=1)], axis=1)
        out = np.empty(T.shape[0])
        step = max(1, int(4e6 // ((self.n + 1) ** 2)))
        for a in range(0, T.shape[0], step):
            out[a : a + step] = self._max_all_arcs(S[a : a + step], al0)
        return out

    def node_pvalue_true_shuffle(self, ostat, hybrid, m, rng, al0=2, kmax=25, chunk=500):
        """fraction of m true shuffles whose statistic reaches `ostat` (= 0.99999 T^2 observed)."""
        hits = 0
        for a in range(0, m, chunk):
            T = self.shuffled_terms(rng, min(chunk, m - a))
            st = self.hybrid_stat(T, al0, kmax) if hybrid else self.exact_stat(T, al0)
            hits += int((ostat <= st).sum())
        return hits / m


def flank_pvalue_true_shuffle(xc, w, n1, n2, m, rng):
    """wtpermp under true shuffles: (p, ostat, needs_permutations).  xc, w = the n1+n2 bins of the test."""
    xc = np.asarray(xc, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    n = n1 + n2
    rw = np.sqrt(w)
    xs1, xs2 = (w[:n1] * xc[:n1]).sum(), (w[n1:] * xc[n1:]).sum()
    rn1, rn2 = w[:n1].sum(), w[n1:].sum()
    rn = rn1 + rn2
    xbar = (xs1 + xs2) / rn
    if n1 <= n2:
        sl, rm1, odiff = slice(0, n1), rn1, abs(xs1 / rn1 - xbar)
    else:
        sl, rm1, odiff = slice(n1, n), rn2, abs(xs2 / rn2 - xbar)
    ostat = 0.99999 * odiff
    y = xc * rw
    hits = 0
    for a in range(0, m, 1000):
        k = min(1000, m - a)
        Y = rng.permuted(np.tile(y, (k, 1)), axis=1)
        xs = (Y[:, sl] * rw[None, sl]).sum(axis=1)
        hits += int((ostat <= np.abs(xs / rm1 - xbar)).sum())
    return hits / m


def z_score(k1, m1, p2, m2):
    """two-proportion z of k1/m1 against p2 (from m2 draws), pooled variance; 0 when both are degenerate."""
    p1 = k1 / m1
    pool = (k1 + p2 * m2) / (m1 + m2)
    var = pool * (1.0 - pool) * (1.0 / m1 + 1.0 / m2)
    if var <= 0.0:
        return 0.0
    return (p1 - p2) / np.sqrt(var)


def stat_nodes(kind, count, seed):
    """Seeded nodes for the p-value tests: (x, w) pairs, %.6g-quantised like the real boundary.

    kind 'hybrid': 201..900 bins; 'exact': 12..200 bins.  Half pure noise, half with a weak step or a short
    blip so that the observed statistic sits where the permutation p-value is neither 0 nor 1."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(count):
        n = int(rng.integers(201, 900)) if kind == "hybrid" else int(rng.integers(12, 201))
        x = rng.normal(0.0, 0.2, n)
        style = k % 4
        if style == 1:        # weak step
            c = int(rng.integers(n // 4, 3 * n // 4))
            x[c:] += rng.choice([-1, 1]) * 0.2 * rng.uniform(1.0, 3.0) / np.sqrt(min(c, n - c)) * 2.0
        elif style == 2:      # short blip (the arcs the hybrid permutation statistic looks at)
            a = int(rng.integers(2, n - 12))
            x[a : a + int(rng.integers(2, 9))] += rng.choice([-1, 1]) * rng.uniform(0.15, 0.4)
        w = rng.uniform(0.5, 1.0, n) if k % 2 == 0 else np.ones(n)
        out.append((q6(x), q6(w)))
    return out


def noise_arms(n, count, seed, weighted):
    """`count` pure-noise arms of n bins.  Weighted arms follow the model the weights stand for
    (var(log2_i) = sigma^2 / w_i), so that y_i = xc_i sqrt(w_i) is exchangeable under the null."""
    rng = np.random.default_rng(seed)
    if weighted:
        w = q6(rng.uniform(0.5, 1.0, n * count))
        x = q6(rng.normal(0.0, 0.25, n * count) / np.sqrt(w))
    else:
        w = np.ones(n * count)
        x = q6(rng.normal(0.0, 0.25, n * count))
    off = np.arange(count + 1, dtype=np.int64) * n
    return x, w, off


def false_split_arms(seg_arm, count):
    """number of arms that came back as more than one segment"""
    return int((np.bincount(np.asarray(seg_arm), minlength=count) > 1).sum())


def binom_band(n, p, tail=1e-4):
    """[lo, hi] such that a Binomial(n, p) count falls outside with probability <= 2*tail."""
    from scipy.stats import binom

    return int(binom.ppf(tail, n, p)), int(binom.ppf(1.0 - tail, n, p))

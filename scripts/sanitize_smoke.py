#!/usr/bin/env python
"""Small pass over the kernels that synchronise through global memory (CBS: look-back flags, last-CTA counters, the
node lock, the tile queue, cooperative permutation kernels; rolling median; bootstrap) for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/sanitize_smoke.py
    compute-sanitizer --tool racecheck python scripts/sanitize_smoke.py
Checks the results against the CPU oracle as it goes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

from cnvkit_b200 import segmetrics, smoothing  # noqa: E402
from cnvkit_b200.segmentation import cbs  # noqa: E402
from oracle import cbs as OC  # noqa: E402
from oracle import np_oracle as O  # noqa: E402

rng = np.random.default_rng(0)
sizes = [1, 3, 40, 150, 199, 257, 600, 2500, 9000]
xs = []
for n in sizes:
    x = rng.normal(0, 0.2, n)
    if n > 100:
        x[n // 3: n // 2] += 0.45
        x[n // 2 + 5: n // 2 + 9] -= 0.6
    xs.append(x)
x = np.array([float("%.6g" % v) for v in np.concatenate(xs)])
w = np.array([float("%.6g" % v) for v in rng.uniform(0.5, 1, len(x))])
off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
for alpha in (1e-4, 1e-2):
    g = cbs.segment_arms(x, w, off, alpha=alpha)
    o = OC.segment(x, w, off, OC.make_params(alpha=alpha), nthreads=1)
    for a, b in zip(g[:3], o[:3]):
        np.testing.assert_array_equal(a, b)
    print(f"cbs alpha={alpha}: {len(g[0])} segments == oracle; stats {g[4]}")
d = cbs.node_diag(x[off[5]:off[8]], w[off[5]:off[8]], [0, 257, 857, 3357], alpha=1e-4, nperm=500)
print("node_diag ok", [r["nrej"] for r in d])
y = rng.normal(0, 1, 20000)
np.testing.assert_array_equal(smoothing.rolling_median(y, 0.05), O.rolling_median(y, 0.05))
print("rolling median ok")
ci = segmetrics.bootstrap_ci(x, w, g[1], g[2], alpha=0.05, bootstraps=50, smoothed=10)
print("bootstrap ok", float(ci[0][0]))
sm = cbs.smooth_cna(x, off)
print("smooth_cna ok", int((sm != x).sum()))

"""A node that only the full permutation test can accept: |t| = 6.8 on a 100 k-bin arm (development probe)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cnvkit_b200 import _lib
from cnvkit_b200.segmentation import cbs
n = 100_000
base = np.random.default_rng(5)
noise = base.normal(0, 0.25, n)
w = np.array([float("%.6g" % v) for v in base.uniform(0.5, 1, n)])
off = np.array([0, n])
for delta in (0.0136,):
    x = noise.copy(); x[n // 2:] += delta
    x = np.array([float("%.6g" % v) for v in x])
    for rep in range(3):
        _lib.prof_enable(rep == 2); _lib.prof_reset()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        g = cbs.segment_arms(x, w, off, alpha=1e-4)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    pr = {k: round(v[0], 3) for k, v in _lib.prof_get().items() if v[1] and v[0] > 0.1}
    print(f"delta={delta:.4f}: {1e3*dt:.2f} ms segments={len(g[0])} perms_run={g[4]['perms_run']} perm_arcs={g[4]['perm_arcs']:.3g} {pr}")

"""Race probe for the programmatic-dependent-launch chain of the CBS level kernels: the pruned and the exhaustive
search must give the same table (tests/test_cbs_gpu.py::test_c3_scale_properties), repeatedly, under the mask in
CNVK_PDL.  Usage: CNVK_PDL=0xffef python scripts/pdl_bisect.py [repeats]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnvkit_b200.segmentation import cbs as G  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    rng = np.random.default_rng(33)
    n = 125_000
    x = rng.normal(0, 0.25, n)
    truth = [20_000, 20_400, 61_000, 90_000, 90_050]
    lvl = [0.0, 0.58, 0.0, -1.0, 0.0, 1.0]
    edges = [0] + truth + [n]
    for k in range(len(edges) - 1):
        x[edges[k]: edges[k + 1]] += lvl[k]
    w = rng.uniform(0.5, 1.0, n)
    x, w = np.round(x, 5), np.round(w, 5)
    off = np.array([0, n])
    bad = 0
    ref = None
    for r in range(reps):
        for prune in (True, False):
            _a, lo, hi, mean, _st = G.segment_arms(x, w, off, alpha=1e-4, prune=prune)
            key = (tuple(lo), tuple(hi), tuple(mean))
            if ref is None:
                ref = key
            elif key != ref:
                bad += 1
                print("mismatch rep", r, "prune", prune, list(lo))
    print("CNVK_PDL=%s: %d mismatches in %d runs" % (os.environ.get("CNVK_PDL", "default"), bad, 2 * reps))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

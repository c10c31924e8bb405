#!/usr/bin/env python
"""Host-side profile of the exome sample path (cohort.Panel.run, one sample per call): cProfile of 32 samples
after warm-up, top functions by cumulative and by own time.  Run on the GPU box:
    python scripts/profile_exome_host.py > gpurun_out/exome_host_profile.txt
"""
import cProfile
import io
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from cnvkit_b200 import cohort  # noqa: E402


def main():
    bins, ref, samples, cols, (T, A, R) = bench.exome_world(4)
    panel = cohort.Panel(T, A, R)
    work = cols * 8
    panel.run(work[:4], method="cbs", batch=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    panel.run(work, method="cbs", batch=1)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / len(work)
    print(f"wall per sample: {1e3 * dt:.3f} ms ({1 / dt:.1f} samples/s), batch 1")
    pr = cProfile.Profile()
    pr.enable()
    panel.run(work, method="cbs", batch=1)
    torch.cuda.synchronize()
    pr.disable()
    for key in ("cumulative", "tottime"):
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
        print(s.getvalue())


if __name__ == "__main__":
    main()

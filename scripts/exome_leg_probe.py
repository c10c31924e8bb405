"""Why does bench.py's 4-sample exome leg differ from the 32-sample host profile?  Times repeated
panel.run(cols, batch=1) calls, native batch path on and off (development probe)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import logging

import numpy as np
import torch

import bench
from cnvkit_b200 import cohort

logging.basicConfig(level=logging.ERROR)
bins, ref, samples, cols, (T, A, R) = bench.exome_world(4)
panel = cohort.Panel(T, A, R)
for m in ("cbs", "haar"):
    panel.run(cols[:2], method=m)
for native in (True, False, True):
    if not native:
        panel._native_ok = lambda method: False
    else:
        panel.__dict__.pop("_native_ok", None)
    for m in ("cbs", "haar"):
        ts = []
        for rep in range(6):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            panel.run(cols, method=m, batch=1)
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0) / len(cols))
        print("native" if native else "step-by-step", m, " ".join("%.2f" % t for t in ts), flush=True)

# ---- the same after a threaded cohort run in this process (bench.py runs its cohort leg first)
import gc

resident = [tuple({k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in part.items()} for part in c) for c in cols]
work = [resident[i % len(resident)] for i in range(256)]
t0 = time.perf_counter()
panel.run(work, method="cbs", batch=8, workers=4)
torch.cuda.synchronize()
print("cohort 256 samples, 4 workers: %.1f samples/s" % (256 / (time.perf_counter() - t0)), flush=True)
print("reserved MB", torch.cuda.memory_reserved() >> 20, "gc counts", gc.get_count(), "threads", __import__("threading").active_count(), flush=True)
for label in ("after cohort", "after gc.collect + empty_cache"):
    for m in ("cbs", "haar"):
        ts = []
        for rep in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            panel.run(cols, method=m, batch=1)
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0) / len(cols))
        print(label, m, " ".join("%.2f" % t for t in ts), flush=True)
    gc.collect()
    torch.cuda.empty_cache()

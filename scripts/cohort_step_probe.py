"""Where does a 128-sample cohort step spend its time?  Phase and per-chunk timestamps of bench.py's cohort step
(development probe)."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import logging

import numpy as np
import torch

import bench
from cnvkit_b200 import cohort, shard

logging.basicConfig(level=logging.ERROR)
n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 128
bins, ref, smp, cols, (T, A, R) = bench.exome_world(16)
panel = cohort.Panel(T, A, R)
resident = [tuple({k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in part.items()} for part in c) for c in cols]
gatherer = shard.TableGatherer(ncols=7, capacity=n_samples * 512 + 16, device=None)
events = []
orig = panel._chunk_native


def traced(work, sids, *a, **k):
    t0 = time.perf_counter()
    r = orig(work, sids, *a, **k)
    events.append((threading.get_ident(), t0, time.perf_counter(), len(work)))
    return r


panel._chunk_native = traced


def step():
    t0 = time.perf_counter()
    work = [resident[i % 16] for i in range(n_samples)]
    res = panel.run(work, method="cbs", batch=8, workers=4, sample_ids=[f"S{i}" for i in range(n_samples)])
    t1 = time.perf_counter()
    for i, (_c, cns) in enumerate(res):
        d = cns.data
        gatherer.append([np.full(len(d), i)] + [d[c].to_numpy() for c in ("start", "end", "log2", "probes", "weight", "depth")])
    t2 = time.perf_counter()
    gatherer.flush(dst=0)
    t3 = time.perf_counter()
    return t0, t1, t2, t3


for rep in range(4):
    events.clear()
    torch.cuda.synchronize()
    t0, t1, t2, t3 = step()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    print("rep %d: run %.1f ms, append %.1f ms, flush %.1f ms, sync %.1f ms; total %.1f ms = %.0f samples/s" % (
        rep, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t4 - t0), n_samples / (t4 - t0)))
    tids = sorted({e[0] for e in events})
    for tid in tids:
        ch = [(1e3 * (a - t0), 1e3 * (b - t0)) for t, a, b, _n in events if t == tid]
        print("   thread %d chunks: %s" % (tids.index(tid), " ".join("%.0f-%.0f" % c for c in ch)))

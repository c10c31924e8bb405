"""Throughput probe: C2-size exome samples through cohort.Panel (development)."""
import sys, os, time, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cnvkit_b200 import synth, cohort, _lib
logging.basicConfig(level=logging.ERROR)
b = synth.exome_bins(); r = synth.exome_reference(b)
S = [synth.exome_sample(b, r, seed=k) for k in range(8)]
T, A, R = synth.exome_tables(b, r, S[0])
t0 = time.perf_counter(); panel = cohort.Panel(T, A, R); torch.cuda.synchronize(); print("panel build %.1f ms" % (1e3*(time.perf_counter()-t0)))
t = b["is_target"]
cols = [({"log2": s["log2"][t], "depth": s["depth"][t]}, {"log2": s["log2"][~t], "depth": s["depth"][~t]}) for s in S]
for method in ("cbs", "haar"):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        trips = [panel.fix(*c) for c in cols]
        torch.cuda.synchronize(); t1 = time.perf_counter()
        cns = panel.segment(trips, method)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"{method} rep{rep}: fix {1e3*(t1-t0)/len(S):.1f} ms/sample  segment {1e3*(t2-t1)/len(S):.1f} ms/sample  segs {[len(c) for c in cns][:4]}")
import cProfile, pstats, io
pr = cProfile.Profile(); pr.enable()
trips = [panel.fix(*c) for c in cols]; cns = panel.segment(trips, "cbs")
pr.disable(); st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(45); print(st.getvalue()[:7000])

"""Throughput probe: file-to-file cohort run (cohort.run_files): .cnn files in, .cnr/.cns files (with the
batch command's bootstrap ci columns) out.  Development."""
import logging
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cnvkit_b200 import cohort, synth, tabio  # noqa: E402

logging.basicConfig(level=logging.ERROR)
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 16
b = synth.exome_bins()
r = synth.exome_reference(b)
d = tempfile.mkdtemp()
t0 = time.perf_counter()
tf, af = [], []
R = None
for k in range(NS):
    T, A, R = synth.exome_tables(b, r, synth.exome_sample(b, r, seed=k), sample_id=f"S{k}")
    tf.append(os.path.join(d, f"S{k}.targetcoverage.cnn"))
    af.append(os.path.join(d, f"S{k}.antitargetcoverage.cnn"))
    tabio.write(T, tf[-1])
    tabio.write(A, af[-1])
print(f"wrote {2 * NS} input files in {time.perf_counter() - t0:.2f} s", flush=True)
cohort.run_files(tf[:4], af[:4], R, os.path.join(d, "warm"), ci=True, workers=4, io_threads=8, batch=1)
for ci, write_cnr in ((True, True), (True, False), (None, True)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = cohort.run_files(tf, af, R, os.path.join(d, f"out{ci}{write_cnr}"), ci=ci, workers=4, io_threads=8, batch=2,
                           write_cnr=write_cnr)
    dt = time.perf_counter() - t0
    print(f"run_files ci={ci} write_cnr={write_cnr}: {NS} samples in {dt:.2f} s = {NS / dt:.1f} samples/s "
          f"({len(os.sched_getaffinity(0))} host cores)", flush=True)

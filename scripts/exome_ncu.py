"""One C2 exome sample through cohort.Panel (fix + CBS), for an ncu launch list (development probe)."""
import sys, os, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cnvkit_b200 import synth, cohort
logging.basicConfig(level=logging.ERROR)
b = synth.exome_bins(); r = synth.exome_reference(b)
S = [synth.exome_sample(b, r, seed=k) for k in range(3)]
T, A, R = synth.exome_tables(b, r, S[0])
panel = cohort.Panel(T, A, R)
t = b["is_target"]
cols = [({"log2": s["log2"][t], "depth": s["depth"][t]}, {"log2": s["log2"][~t], "depth": s["depth"][~t]}) for s in S]
panel.run(cols[:2], method="cbs", batch=1)          # warm-up
torch.cuda.synchronize()
marker = torch.zeros(1, device="cuda"); marker += 1  # at::... kernel marks the start of the measured sample
torch.cuda.synchronize()
panel.run(cols[2:3], method="cbs", batch=1)
torch.cuda.synchronize()
print("n_target_kept", panel.t.n, "n_anti_kept", panel.a.n, "n_cnr", panel.n)

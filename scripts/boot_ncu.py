"""Dev probe: the bootstrap-CI kernels on an exome-sized table (64 segments over 177 k bins, 100 smoothed
replicates = 17.7 M resampled values), timed with CUDA events around the device-pointer entry point.
Run plain first, then under `ncu --set full -k regex:boot_ -c 4` for the captures in profiles/."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from cnvkit_b200 import _dev, segmetrics  # noqa: E402

rng = np.random.default_rng(1)
n = 177_174
cuts = np.sort(rng.choice(np.arange(50, n - 50), size=63, replace=False))
lo = np.concatenate([[0], cuts])
hi = np.concatenate([cuts, [n]])
v = rng.normal(0, 0.25, n)
w = rng.uniform(0.3, 1.0, n)
dv, dw = _dev.up(v), _dev.up(w)
args = dict(alpha=0.5, bootstraps=100, smoothed=True, device_ptrs=(dv.data_ptr(), dw.data_ptr()), stream=_dev.stream())
for _ in range(3):
    out = segmetrics.bootstrap_ci(None, None, lo, hi, **args)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
a.record()
t0 = time.perf_counter()
for _ in range(reps):
    out = segmetrics.bootstrap_ci(None, None, lo, hi, **args)
b.record()
torch.cuda.synchronize()
print(f"bootstrap_ci_dev: {a.elapsed_time(b) / reps:.3f} ms device span, {1e3 * (time.perf_counter() - t0) / reps:.3f} ms wall per call; "
      f"{100 * n / (a.elapsed_time(b) / reps * 1e-3) / 1e9:.2f} G resampled values/s; ci[0] = {float(out[0][0]):.6f}, {float(out[1][0]):.6f}")

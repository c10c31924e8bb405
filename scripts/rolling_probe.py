"""Device time of rolling_median for exome-sized columns (development probe)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cnvkit_b200 import _lib, _dev
rng = np.random.default_rng(0)
for n, wing in ((150000, 750), (50000, 250), (150000, 3000), (150000, 7500), (3000000, 15000)):
    x = _dev.up(rng.normal(0, 1, n)); out = _dev.empty(n, torch.float64)
    for _ in range(3):
        _lib.call("cnvk_rolling_median_dev", _dev.ptr(x), n, wing, _dev.ptr(out), _dev.stream())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10):
        _lib.call("cnvk_rolling_median_dev", _dev.ptr(x), n, wing, _dev.ptr(out), _dev.stream())
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"n={n} wing={wing}: {ms*1e3:.0f} us  ({16*n/ms/1e6:.1f} GB/s algorithmic)")

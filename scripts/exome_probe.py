"""Where does the time go for one C2 exome sample through the public API? (development probe)"""
import cProfile, pstats, time, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cnvkit_b200 import synth, fix, segmentation, _lib
import logging
logging.basicConfig(level=logging.ERROR)
b = synth.exome_bins(); r = synth.exome_reference(b); s = synth.exome_sample(b, r)
T, A, R = synth.exome_tables(b, r, s)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    cnr = fix.do_fix(T, A, R)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    cns = segmentation.do_segmentation(cnr, "cbs")
    torch.cuda.synchronize(); t2 = time.perf_counter()
    cnh = segmentation.do_segmentation(cnr, "haar")
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"rep{rep}: fix {1e3*(t1-t0):.1f} ms  cbs {1e3*(t2-t1):.1f} ms ({len(cns)} segs)  haar {1e3*(t3-t2):.1f} ms ({len(cnh)} segs)")
_lib.prof_enable(True); _lib.prof_reset()
pr = cProfile.Profile(); pr.enable()
cnr = fix.do_fix(T, A, R); cns = segmentation.do_segmentation(cnr, "cbs"); cnh = segmentation.do_segmentation(cnr, "haar")
pr.disable()
print({k: round(v[0], 3) for k, v in _lib.prof_get().items() if v[1]})
st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(45); print(st.getvalue()[:9000])

#!/usr/bin/env python
"""Headline benchmark: bins segmented / s -- weighted CBS of one synthetic WGS sample (C3 of
BASELINE.json: 3 M 1-kb bins, 24 chromosomes, 48 arms, nperm = 10 000, alpha = 1e-4).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--bins B]

A "step" is one complete segmentation of the sample (every arm, every recursion level,
every permutation test) producing the segment table.  N > 1 (launched under torchrun, one rank
per GPU): arms are sharded over ranks by longest-processing-time on n^2, each rank segments its
arms with no data-path collective, and the per-rank tables are gathered with NCCL (strong
scaling: total work fixed).  `value` = bins / max-over-ranks device time with the log2/weight
columns already resident in HBM; `e2e` = the same call through the host-buffer C ABI
(cnvk_cbs_segment) from pinned host memory, H2D and the table read-back inside the timed region.

`--impl reference` times the CPU oracle (oracle/cbs_oracle.c, OpenMP over arms; a restatement of
DNAcopy -- R is not installed here) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL
# prints its version banner there), so fd 1 is pointed at stderr for the whole run and the JSON line is
# written to a duplicate of the original stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


METRIC = "bins segmented/s (CBS, WGS 1 kb)"
UNIT = "bins/s"
ALPHA = 1e-4
FLOP_PER_ARC = 7  # sub, sub, sub, mul, mul, mul, compare -- weighted BSS, division-free test
PREP_BYTES_PER_BIN = 48  # sums pass reads log2, weight (16 B); scan pass reads them again and writes S, cw (32 B)
PREP_TRAFFIC = 109.9e6   # ncu dram bytes of the level-1 launches (sums 48.0 MB + scan 53.9 + 7.9 MB; profiles/r02f_summary.md)
PREP_NOTE = ("two tile kernels per level: sums (min/max, double-double totals) and scan (centring, blocked FP64 prefix "
             "scan, carries by chained look-back, block extrema); they move exactly the algorithmic 48 B/bin (xc, y, wn of "
             "round 1 are no longer written: the exchangeable terms are materialised only for nodes that enter a "
             "permutation test); traffic = ncu dram bytes of the level-1 launches (profiles/r02f_summary.md), most of "
             "the S / cw writes are still in L2 when the capture ends; both kernels are latency-bound (3 CTAs/SM in the "
             "scan kernel, long-scoreboard stalls), 1.1 TB/s at level 1")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bins", type=int, default=3_000_000)
    ap.add_argument("--no-prune", action="store_true", help="exhaustive arc scan (roofline study)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time per reference step")
    ap.add_argument("--workload", default="wgs", choices=["wgs", "cohort", "reference"],
                    help="wgs: C3 (headline, default); cohort: C5 exome cohort fix + CBS sharded by sample")
    ap.add_argument("--samples", type=int, default=1024, help="cohort size for --workload cohort")
    ap.add_argument("--workers", type=int, default=4, help="host threads (one CUDA stream each) per rank, cohort workload")
    ap.add_argument("--cohort-batch", type=int, default=8, help="samples per segmentation call, cohort workload")
    ap.add_argument("--no-exome-leg", action="store_true", help="skip the C2 exome leg of the default run")
    ap.add_argument("--no-cohort-leg", action="store_true", help="skip the C5 cohort leg of the default run")
    ap.add_argument("--no-extra-legs", action="store_true", help="skip the cliff / public-API / C1 legs (N = 1)")
    ap.add_argument("--cohort-samples", type=int, default=1024, help="cohort size of the C5 leg of the default run")
    ap.add_argument("--sharding", default="samples", choices=["samples", "arms"],
                    help="N > 1: 'samples' = one WGS sample per rank (weak scaling, default); "
                         "'arms' = the arms of ONE sample over the ranks (strong scaling, BASELINE configs[2])")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------ helpers
class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(f[0]))
                smmax.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def reference_sample(data, cpu_seconds, cores):
    """The CPU arm's sample.  The oracle's block-pruned search (what DNAcopy's wtmaxo does; same result
    as its exhaustive scan) segments the whole 3 M-bin sample in a few seconds, so the sample is the full
    workload unless the time budget is tiny, in which case the shortest arms are used."""
    off = data["arm_offsets"]
    sizes = np.diff(off)
    if cpu_seconds >= 2.0 or len(sizes) <= 2:
        picked = list(range(len(sizes)))
    else:
        picked = sorted(np.argsort(sizes, kind="stable")[: max(2, len(sizes) // 2)].tolist())
    x = np.concatenate([data["log2"][off[a]:off[a + 1]] for a in picked])
    w = np.concatenate([data["weight"][off[a]:off[a + 1]] for a in picked])
    o = np.concatenate([[0], np.cumsum(sizes[picked])]).astype(np.int64)
    return x, w, o, picked


def wgs_config(args, world, by_arms):
    """The `config` object of the headline line -- the same keys and values for both arms (--impl b200 and
    --impl reference), so that the driver's same_config test compares like with like."""
    return {
        "workload": f"C3 synthetic WGS, {args.bins} 1-kb bins, 48 arms, CBS alpha={ALPHA} nperm=10000",
        "bins_per_sample": int(args.bins), "arms_per_sample": 48, "alpha": ALPHA, "nperm": 10000,
        "samples_in_flight": 1 if (by_arms or world == 1) else world,
        "sharding": "single GPU" if world == 1 else
        (f"arms of one sample over {world} ranks by LPT on n^2; one NCCL gather of the tables" if by_arms else
         f"one sample per rank ({world} replicas of the same synthetic sample), no data-path collective; "
         f"one NCCL all-gather of the tables per batch of steps"),
        "l2": "256 MB flush write between timed steps",
    }


def bind_to_gpu_numa(torch, local_rank):
    """Pin this rank's host threads to the CPUs of the GPU's NUMA node BEFORE allocating pinned staging
    buffers (first touch places them on that node), so that N ranks' host->device copies do not all cross
    one socket.  Best effort; returns a short description for the JSON line."""
    try:
        props = torch.cuda.get_device_properties(local_rank)
        dom = getattr(props, "pci_domain_id", 0)
        bus, devid = props.pci_bus_id, props.pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0"
        node = int(open(path + "/numa_node").read().strip())
        cpulist = open(path + "/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return f"node {node}: not bound"
        os.sched_setaffinity(0, cpus)
        return f"bound to NUMA node {node} ({len(cpus)} cpus)"
    except Exception as exc:  # sysfs layout / torch attribute missing: run unbound
        return f"not bound ({type(exc).__name__})"


# ------------------------------------------------------------------------------------------ exome legs
EXOME_METRIC = "exome samples/s (fix + CBS segment, 200k bins)"


def exome_world(n_distinct, seed0=1):
    """C2 bin set + pooled reference + `n_distinct` tumours as plain columns (host)."""
    from cnvkit_b200 import synth

    bins = synth.exome_bins()
    ref = synth.exome_reference(bins)
    samples = [synth.exome_sample(bins, ref, seed=seed0 + k) for k in range(n_distinct)]
    t = bins["is_target"]
    cols = [({"log2": s["log2"][t], "depth": s["depth"][t]}, {"log2": s["log2"][~t], "depth": s["depth"][~t]})
            for s in samples]
    tables = synth.exome_tables(bins, ref, samples[0], sample_id="S0")
    return bins, ref, samples, cols, tables


def exome_cpu_baseline(bins, ref, sample, cores, methods=("cbs", "haar")):
    """One C2 sample through the CPU oracle pipeline (oracle/pipeline.py: pandas rolling medians as in the
    reference, C restatements of DNAcopy CBS / HaarSeg with OpenMP over arms)."""
    from oracle import pipeline as OP

    refc = dict(log2=ref["log2"], depth=ref["depth"], spread=ref["spread"], gc=bins["gc"], rmask=bins["rmask"])
    t0 = time.perf_counter()
    fx = OP.do_fix(bins, sample, refc)
    t1 = time.perf_counter()
    k = np.flatnonzero(fx["keep"])
    chrom, start, end = np.asarray(bins["chromosome"], dtype=object)[k], bins["start"][k], bins["end"][k]
    out = {"fix_s": t1 - t0}
    for m in methods:
        c0 = time.perf_counter()
        res, _ = OP.segment_columns(chrom, start, end, fx["log2"], fx["weight"], m, nthreads=cores, prune=True)
        out[m + "_s"] = time.perf_counter() - c0
        out[m + "_segments"] = int(len(res[0]))
    return out


def exome_reference_python(bins, ref, sample):
    """The reference's OWN Python (cnvlib.fix.do_fix + cnvlib.segmentation.do_segmentation(method='haar'), one
    process, unmodified code from baseline/_ref or /root/reference) on one C2 sample -- the CPU pipeline a user
    runs today, minus R (absent, so CBS itself cannot be timed with the reference's code).  None if the reference
    tree did not travel."""
    try:
        from oracle import refshim

        if not refshim.available():
            return None
        R = refshim.load()
    except Exception as exc:   # missing dependency of the reference on this box: report, do not fail the bench
        return {"error": repr(exc)}
    import logging

    from cnvkit_b200 import synth

    logging.disable(logging.WARNING)
    T, A, Rf = synth.exome_tables(bins, ref, sample, sample_id="S0")
    mk = lambda t: R.cnary.CopyNumArray(t.data.copy(), {"sample_id": "S0"})
    t0 = time.perf_counter()
    cnr = R.fix.do_fix(mk(T), mk(A), mk(Rf))
    t1 = time.perf_counter()
    cns = R.segmentation.do_segmentation(cnr, "haar")
    t2 = time.perf_counter()
    logging.disable(logging.NOTSET)
    return {"kind": "reference", "cores": 1, "fix_s": t1 - t0, "haar_s": t2 - t1, "bins": int(len(cnr)),
            "segments_haar": int(len(cns)), "samples_per_s_fix_haar": 1.0 / (t2 - t0),
            "sample": "one C2 sample through cnvlib.fix.do_fix + cnvlib.segmentation.do_segmentation('haar'), the "
                      f"unmodified reference from {R.root}, one process (R/DNAcopy absent: its CBS cannot run)"}


def exome_leg(args, torch, dev):
    """C2: fix + CBS + Haar of single exome samples through cohort.Panel (host columns in, .cns tables
    out -- copies inside the timed region), plus per-kernel device time for the rolling-median path."""
    from cnvkit_b200 import _lib, cohort

    bins, ref, samples, cols, (T, A, R) = exome_world(4)
    t0 = time.perf_counter()
    panel = cohort.Panel(T, A, R)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    for m in ("cbs", "haar"):
        panel.run(cols[:2], method=m)          # warm-up
    out = {"metric": EXOME_METRIC, "unit": "samples/s", "workload": "C2 synthetic exome: 150k targets + 50k antitargets, "
           "48 arms; one sample per call (batch 1)", "panel_build_s": build_s}
    for m in ("cbs", "haar"):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = panel.run(cols, method=m, batch=1)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / len(cols)
        out[f"ms_per_sample_fix_{m}"] = 1e3 * dt
        out[f"segments_{m}"] = [len(r[1]) for r in res]
    out["value"] = 1.0 / (out["ms_per_sample_fix_cbs"] * 1e-3)
    # the same with the batch command's bootstrap confidence intervals per segment (batch.py:559-566)
    panel.run(cols[:2], method="cbs", ci=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res_ci = panel.run(cols, method="cbs", batch=1, ci=True, keep_cnr=True)
    torch.cuda.synchronize()
    out["ms_per_sample_fix_cbs_ci_incl_cnr_copy"] = 1e3 * (time.perf_counter() - t0) / len(cols)
    cnr0, cns0 = res_ci[0]
    from cnvkit_b200 import segmetrics as _sm
    b_lo, b_hi = _sm.segment_bin_ranges(cnr0.data, cns0.data)
    v0, w0 = cnr0["log2"].to_numpy(), cnr0["weight"].to_numpy()
    _sm.bootstrap_ci(v0, w0, b_lo, b_hi, 0.5, 100, True)
    t0 = time.perf_counter()
    for _ in range(5):
        _sm.bootstrap_ci(v0, w0, b_lo, b_hi, 0.5, 100, True)
    ci_ms = 1e3 * (time.perf_counter() - t0) / 5
    from oracle import segmetrics as _OS
    t0 = time.perf_counter()
    _OS.ci_table(v0, w0, b_lo, b_hi, 0.5, 100, True)
    ci_cpu_s = time.perf_counter() - t0
    out["bootstrap_ci"] = {"segments": int(len(b_lo)), "bins": int(len(v0)), "resampled_values": int(100 * (b_hi - b_lo).sum()),
                           "ms_host_call": ci_ms, "cpu_port_s": ci_cpu_s,
                           "note": "cnvk_bootstrap_ci with host columns (H2D + 4 kernels + D2H) vs oracle/segmetrics.py "
                                   "(numpy, 1 core) on the first sample's segments; alpha 0.5, 100 smoothed replicates"}
    # device time by kernel class for one sample (fix + cbs + haar)
    _lib.prof_enable(True)
    _lib.prof_reset()
    panel.run(cols[:1], method="cbs")
    panel.run(cols[:1], method="haar")
    prof = _lib.prof_get()
    _lib.prof_enable(False)
    out["kernel_ms_per_sample"] = {k: v[0] for k, v in prof.items() if v[1] > 0}
    # rolling median: 16 B/element algorithmic over the 4 centre-by-window passes x 2 runs
    n_t, n_a = int(panel.t.n), int(panel.a.n)
    rm_elems = 2 * (2 * n_t + 2 * n_a)
    rm_ms = sum(out["kernel_ms_per_sample"].get(k, 0.0) for k in ("radix_sort", "wavelet_build", "window_query"))
    out["rolling_median"] = {"elements": rm_elems, "ms": rm_ms, "achieved_GBps": 16.0 * rm_elems / (rm_ms * 1e-3) / 1e9
                             if rm_ms else None, "note": "16 B/element algorithmic; on-chip selection bound, not HBM"}
    # fix path: 80 B/bin fused elementwise + 24 B/bin per centre-by-window pass (4 passes per sample),
    # over the two runs above (cbs + haar => the fix kernels ran twice)
    fix_ms = out["kernel_ms_per_sample"].get("fix", 0.0) / 2.0
    n_bins = n_t + n_a
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6540.8))
    fix_bytes = 80.0 * n_bins + 24.0 * (2 * n_t + 2 * n_a)
    out["fix_roofline"] = {"kernel": "fix.cu / select.cu / biweight.cu kernels of one sample (class `fix`)", "bound": "hbm",
                           "achieved": fix_bytes / (fix_ms * 1e-3) / 1e9 if fix_ms else None, "peak": hbm, "unit": "GB/s",
                           "frac": (fix_bytes / (fix_ms * 1e-3) / 1e9 / hbm) if fix_ms else None, "traffic": None,
                           "algorithmic": f"80 B/bin x {n_bins} bins + 24 B/bin x 4 centre-by-window passes",
                           "kernel_ms_per_sample": fix_ms,
                           "note": "dozens of small launches per sample (medians by radix select, gathers, elementwise "
                                   "passes over 150 k / 50 k bins): launch-latency-bound, not bandwidth-bound"}
    cores = len(os.sched_getaffinity(0))
    cpu = exome_cpu_baseline(bins, ref, samples[0], cores)
    out["cpu_baseline_reference_python"] = exome_reference_python(bins, ref, samples[0])
    out["cpu_baseline"] = {"value": 1.0 / (cpu["fix_s"] + cpu["cbs_s"]), "unit": "samples/s", "cores": cores,
                           "kind": "port", "detail": cpu,
                           "sample": "one C2 sample, oracle/pipeline.py (fix: pandas rolling median as the reference; "
                                     "CBS: oracle/cbs_oracle.c OpenMP over arms)"}
    return out


# ------------------------------------------------------------------------------------------ arms
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from cnvkit_b200 import synth
    from oracle import cbs as OC

    cores = len(os.sched_getaffinity(0))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    data = synth.wgs_sample(total_bins=args.bins)
    x, w, off, picked = reference_sample(data, args.cpu_seconds, cores)
    P = OC.make_params(alpha=ALPHA, prune=True)
    OC.lib()
    times = []
    nseg = 0
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        arm, lo, hi, mean, st = OC.segment(x, w, off, P, nthreads=cores)
        dt = time.perf_counter() - t0
        nseg = len(arm)
        if step >= args.warmup:
            times.append(dt)
        if step == 0 and dt > 60:  # keep the whole run within minutes
            args.warmup = 0
            if args.steps > 2:
                args.steps = 2
    ms = 1e3 * float(np.mean(times))
    value = len(x) / (ms * 1e-3)
    sample = (f"{len(picked)} of {len(data['arm_offsets']) - 1} arms of the {args.bins}-bin sample ({len(x)} bins), "
              f"full CBS incl. permutations")
    by_arms = args.sharding == "arms" and world > 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong" if by_arms else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": wgs_config(args, world, by_arms),
        "segments": nseg,
        "note": "CPU throughput of ONE sample on all host cores; it does not grow with --gpus (the GPU arm at N > 1 "
                "segments N samples at once, so the N-GPU ratio is a throughput ratio)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample + "; oracle/cbs_oracle.c (restated DNAcopy with block-pruned wtmaxo; R "
                                            "unavailable), OpenMP over arms"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def cliff_sample():
    """One 120 000-bin arm whose maximal statistic (|t| = 6.81) lies between the reach of the hybrid tail
    approximation (tailp = 1.9e-5 <= alpha) and the 7.0 shortcut: only DNAcopy's full 9 500 clean permutations of
    24 n short arcs each can accept the split (tests/test_cbs_stats_gpu.py uses the same sample)."""
    from cnvkit_b200 import synth

    rng = np.random.default_rng(0xC11FF)
    n = 120000
    w = rng.uniform(0.5, 1.0, n)
    x = rng.normal(0, 0.25, n)
    x[45000:] += 0.009
    return synth.round_sig6(x), synth.round_sig6(w), np.array([0, n], dtype=np.int64)


def extra_legs(args, torch, dev, data):
    """N = 1 only: (a) the permutation-cliff sample, (b) C3 through the public do_segmentation(cnarr, 'cbs') with the
    table in host memory (pandas in, .cns out), (c) C1: the reference's own exome-sized .cnr shape through
    do_segmentation.  Reported beside the headline; none of them enters `value`."""
    import pandas as pd

    from cnvkit_b200 import _lib, synth
    from cnvkit_b200.cnary import CopyNumArray
    from cnvkit_b200.segmentation import cbs, do_segmentation

    out = {}
    stream = torch.cuda.current_stream()
    # (a) cliff
    cx, cw, coff = cliff_sample()
    dcx, dcw = torch.from_numpy(cx).to(dev), torch.from_numpy(cw).to(dev)
    ms = []
    st = None
    for k in range(3):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        r = cbs.segment_arms(None, None, coff, alpha=ALPHA, device_ptrs=(dcx.data_ptr(), dcw.data_ptr()),
                             stream=_lib.stream_ptr(stream))
        b.record(stream)
        torch.cuda.synchronize()
        st = r[4]
        if k:
            ms.append(a.elapsed_time(b))
    out["cliff"] = {"workload": "one 120 000-bin arm, |t| = 6.81: accepted only by 9 500 clean hybrid permutations "
                                "(24 n arcs each)", "ms_per_step": float(np.mean(ms)), "bins_per_s": 120000 / (np.mean(ms) * 1e-3),
                    "perms_run": st["perms_run"], "perm_arcs": st["perm_arcs"], "segments": int(len(r[0])),
                    "perm_arcs_per_s": st["perm_arcs"] / (np.mean(ms) * 1e-3)}
    # (b) public API on C3
    chrom = np.asarray(synth.CHROM_NAMES, dtype=object)[data["chromosome_id"]]
    df = pd.DataFrame({"chromosome": chrom, "start": data["start"], "end": data["end"], "gene": "-",
                       "log2": data["log2"], "depth": np.exp2(data["log2"]) * 100.0, "weight": data["weight"]})
    cnarr = CopyNumArray(df, {"sample_id": "C3"})
    t_api = []
    for k in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cns = do_segmentation(cnarr, "cbs", threshold=ALPHA)
        torch.cuda.synchronize()
        if k:
            t_api.append(time.perf_counter() - t0)
    out["public_api_c3"] = {"call": "cnvkit_b200.segmentation.do_segmentation(cnarr, 'cbs') -- pandas table in host "
                                    "memory in, .cns CopyNumArray out (by_arm, outlier mask, %.6g, CBS, transfer_fields)",
                            "ms_per_sample": 1e3 * float(np.mean(t_api)), "bins_per_s": len(df) / float(np.mean(t_api)),
                            "segments": int(len(cns))}
    # (c) C1-shaped sample: the first 19 089 bins' worth of an exome panel (47 arms) -- the reference's own
    # p2-20_1.cnr cannot travel to the GPU box; tests/golden holds its Haar tables, this leg times the same shape
    bins = synth.exome_bins(n_targets=14000, n_antitargets=5089, seed=0xC1)
    ref = synth.exome_reference(bins)
    smp = synth.exome_sample(bins, ref, seed=7)
    T, A, R = synth.exome_tables(bins, ref, smp, sample_id="C1")
    from cnvkit_b200 import fix as _fix

    t_c1 = []
    for k in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cnr = _fix.do_fix(T, A, R)
        cns1 = do_segmentation(cnr, "cbs", threshold=ALPHA)
        torch.cuda.synchronize()
        if k:
            t_c1.append(time.perf_counter() - t0)
    out["c1_shape"] = {"workload": f"C1 shape: {len(T) + len(A)} bins, do_fix + do_segmentation('cbs') through the "
                                   "pandas API, one sample", "ms_per_sample": 1e3 * float(np.mean(t_c1)),
                       "segments": int(len(cns1))}
    return out


def run_b200(args):
    import torch
    import torch.distributed as dist

    from cnvkit_b200 import _lib, shard, synth
    from cnvkit_b200.segmentation import cbs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(torch, local_rank) if world > 1 else "single rank: not bound"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    by_arms = args.sharding == "arms" and world > 1
    if by_arms:
        data = synth.wgs_sample(total_bins=args.bins)
        owner = shard.arm_owner(data["arm_offsets"], world)
        nbins_total = int(data["arm_offsets"][-1])
    else:  # cohort mode: every rank owns one whole sample; the same synthetic sample on every rank, so that
        # per-GPU work is exactly fixed (CBS cost is data-dependent: one borderline node that needs all
        # 9 500 permutations costs several times the rest of the sample -- see the `cliff` object)
        data = synth.wgs_sample(total_bins=args.bins)
        owner = np.full(len(data["arm_offsets"]) - 1, rank)
        nbins_total = int(data["arm_offsets"][-1]) * world
    x, w, off, mine = shard.take_arms(data["log2"], data["weight"], data["arm_offsets"], owner, rank)
    prune = not args.no_prune

    hx = torch.from_numpy(x).pin_memory()
    hw = torch.from_numpy(w).pin_memory()
    dx = hx.to(dev)
    dw = hw.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()
    steps_max = max(args.steps, 3) + 2
    gatherer = shard.TableGatherer(ncols=5, capacity=4096 * steps_max, device=dev if world > 1 else None)

    def step_resident():
        return cbs.segment_arms(None, None, off, alpha=ALPHA, prune=prune,
                                device_ptrs=(dx.data_ptr(), dw.data_ptr()), stream=_lib.stream_ptr(stream))

    def step_host():
        return cbs.segment_arms(hx.numpy(), hw.numpy(), off, alpha=ALPHA, prune=prune)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def stage_table(res):
        """this step's table -> the pinned gather payload (host memory only; no device work, no sync)"""
        arm, lo, hi, mean = res[:4]
        g_arm = mine[arm] if len(arm) else np.zeros(0, np.int64)
        shift = (data["arm_offsets"][g_arm] - off[arm]) if len(arm) else np.zeros(0, np.int64)
        gatherer.append([np.full(len(arm), 0 if by_arms else rank), g_arm, lo + shift, hi + shift, mean])

    def timed(fn, steps):
        """K steps between two barriers; one CUDA-event pair per step on the launching stream (L2 flushed
        before each pair); the segment tables of the K steps are gathered with ONE collective after the last
        step, timed by its own event pair and charged to the K steps."""
        pairs, res = [], None
        barrier()
        for _ in range(steps):
            flush.zero_()
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            a.record(stream)
            res = fn()
            b.record(stream)
            stage_table(res)
            pairs.append((a, b))
        g0 = torch.cuda.Event(enable_timing=True)
        g1 = torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        full = gatherer.flush(dst=0)
        g1.record(stream)
        barrier()
        step_ms = [a.elapsed_time(b) for a, b in pairs]
        gather_ms = g0.elapsed_time(g1)
        nrows = int(len(full)) // steps if full is not None else 0
        timed.last_steps = {"min": float(np.min(step_ms)), "median": float(np.median(step_ms)),
                            "max": float(np.max(step_ms)), "argmax": int(np.argmax(step_ms))}
        return (float(np.sum(step_ms)) + gather_ms) / steps, gather_ms, res, nrows

    # warm-up (>= 3): allocator pools, module load, clocks, NCCL's lazy communicator set-up
    # (the host-buffer step first: it needs the larger scratch slab, and a slab that grows inside the timed loop costs
    # one 10 ms step -- `step_ms_spread` would show it)
    # The clock sampler (an nvidia-smi poller on rank 0) starts BEFORE the warm-up, so that the warm-up steps run
    # straight into the timed ones: with its 0.3 s start-up pause between them the GPU idled and the first timed step
    # was 8-10 % slower than the median (`step_ms_spread.argmax` was always 0).
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    timed(step_host, 1)
    timed(step_resident, max(args.warmup, 3))

    launches0 = _lib.launch_count()
    t0 = time.time()
    ms_res, gather_ms, res, nseg_total = timed(step_resident, args.steps)
    step_spread = dict(timed.last_steps)
    t1 = time.time()
    launches = _lib.launch_count() - launches0
    # per-kernel attribution: the same K steps again with one CUDA-event pair around every kernel class
    # (kept out of the headline timing: creating and recording ~150 events per step costs ~0.5 ms)
    _lib.prof_enable(True)
    _lib.prof_reset()
    ms_prof, _, _, _ = timed(step_resident, args.steps)
    t1 = time.time()
    prof = _lib.prof_get()
    _lib.prof_enable(False)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    stats = res[4]

    e2e_steps = max(2, min(args.steps, 5))
    ms_host, gather_ms_h, res_h, _ = timed(step_host, e2e_steps)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_step = max_over_ranks(ms_res)
    ms_e2e = max_over_ranks(ms_host)

    cohort = None
    if not args.no_cohort_leg and not args.no_prune:
        try:
            cohort = cohort_core(args, torch, dist, dev, world, rank, local_rank, samples=args.cohort_samples, steps=1,
                                 quiet=True)
        except Exception as exc:  # the headline line must still print
            cohort = {"error": repr(exc)}

    if rank == 0:
        # ---- rooflines (device time per kernel class inside the timed region, CUDA events)
        steps = args.steps
        per_step = {k: v[0] / steps for k, v in prof.items() if v[1] > 0}
        dominant = max(per_step, key=per_step.get) if per_step else "cbs_scan"
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6540.8))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6540.8 GB/s (MEASURED_PEAKS.json absent)"
        tf = np.zeros(1)
        _lib.call("cnvk_fp64_peak", 4096, 5, tf.ctypes.data)
        fp64_peak = float(tf[0])
        arc_ms = per_step.get("cbs_scan", 0.0) + per_step.get("cbs_band", 0.0)
        arcs = stats["obs_arcs"]
        arc_tflops = (arcs * FLOP_PER_ARC) / (arc_ms * 1e-3) / 1e12 if arc_ms > 0 else 0.0
        roof_scan = {
            "kernel": "cbs_scan_kernel + cbs_band_kernel", "bound": "fp64", "achieved": arc_tflops, "peak": fp64_peak,
            "unit": "TFLOP/s", "frac": arc_tflops / fp64_peak if fp64_peak else None, "traffic": 22.4e6,
            "peak_source": "measured in this run: register-only DFMA stream (cnvk_fp64_peak); "
                           "MEASURED_PEAKS.json holds no FP64 figure",
            "algorithmic": f"{FLOP_PER_ARC} FP64 op/arc x {arcs} arcs evaluated per step "
                           f"({stats['obs_subtiles_scanned']} of {stats['obs_subtiles_total']} 256x256 sub-tiles "
                           f"survive the bound test; the rest are narrow-band arcs)",
            "kernel_ms_per_step": arc_ms,
            "note": "non-fused FP64 ops: 1 flop per issued DP instruction, so frac <= 0.5 of the DFMA peak; ncu "
                    "(profiles/r02f_summary.md): FP64 pipe 38-44 % active in cbs_scan, 27-39 % in cbs_band, SM throughput "
                    "58 % in the level-2 band launch; traffic = ncu dram bytes of the level-1 band + scan launches "
                    "(20.7 + 1.7 MB: the CTA-level bound skips 57 % of the strips) -- DRAM is idle, the pipe is the bound",
        }
        prep_ms = per_step.get("cbs_prep", 0.0)
        prep_bytes = PREP_BYTES_PER_BIN * stats["prep_bins"]
        prep_gbs = prep_bytes / (prep_ms * 1e-3) / 1e9 if prep_ms > 0 else 0.0
        roof_prep = {
            "kernel": "cbs_prep_* kernels", "bound": "hbm", "achieved": prep_gbs, "peak": hbm_peak,
            "unit": "GB/s", "frac": prep_gbs / hbm_peak, "traffic": PREP_TRAFFIC, "peak_source": hbm_src,
            "algorithmic": f"{PREP_BYTES_PER_BIN} B/bin x "
                           f"{stats['prep_bins']} bins centred and prefix-summed over {stats['levels']} levels",
            "kernel_ms_per_step": prep_ms,
            "note": PREP_NOTE,
        }
        perm_ms = per_step.get("cbs_perm_hybrid", 0.0) + per_step.get("cbs_perm_exact", 0.0)
        perm_tflops = (stats["perm_arcs"] * 3) / (perm_ms * 1e-3) / 1e12 if perm_ms > 0 else 0.0
        roof_perm = {
            "kernel": "cbs_perm_hybrid_kernel + cbs_perm_exact_kernel (cooperative, persistent)", "bound": "fp64",
            "achieved": perm_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": perm_tflops / fp64_peak if fp64_peak else None, "traffic": 192768.0,
            "peak_source": roof_scan["peak_source"],
            "algorithmic": f"3 FP64 op/arc (add, multiply, compare) x {stats['perm_arcs']} arcs of "
                           f"{stats['perms_run']} permutations over {stats['perm_nodes']} nodes",
            "kernel_ms_per_step": perm_ms,
            "note": "the hot loop issues ONE FP64 add per arc (the multiply and compare are replaced by an integer "
                    "compare of the high word against a conservative limit), so the FP64 pipe is not the bound: the "
                    "permutation + gather prologue (~60-130 integer ops per permuted element) and, for this sample, the "
                    "barriers of 5 chunks are; the `cliff` object holds the large-node figure (1.6 T arcs/s); traffic = "
                    "ncu dram bytes of the level-3 hybrid launch (profiles/r02fin_perm_band_full_raw.csv: 0.19 MB read, "
                    "7 KB written -- the node's columns live in L2): issue active 33 %, ALU pipe 40 %, FP64 pipe 4 %",
        }
        roofline = dict({"cbs_prep": roof_prep, "cbs_perm_hybrid": roof_perm, "cbs_perm_exact": roof_perm}.get(dominant, roof_scan))
        roofline["dominant_class_by_time"] = dominant
        # CPU baseline on a bounded sample (rank 0, N = 1 only)
        cpu = None
        if world == 1:
            from oracle import cbs as OC

            cores = len(os.sched_getaffinity(0))
            sx, sw, soff, picked = reference_sample(data, args.cpu_seconds, cores)
            P = OC.make_params(alpha=ALPHA, prune=True)
            OC.lib()
            c0 = time.perf_counter()
            OC.segment(sx, sw, soff, P, nthreads=cores)
            cdt = time.perf_counter() - c0
            cpu = {"value": len(sx) / cdt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{len(picked)} of {len(data['arm_offsets']) - 1} arms ({len(sx)} bins) of the same sample, "
                             f"one pass, oracle/cbs_oracle.c (restated DNAcopy with block-pruned wtmaxo; R "
                             f"unavailable) with OpenMP over arms"}
        exome = None
        if world == 1 and not args.no_exome_leg and not args.no_prune:
            try:
                exome = exome_leg(args, torch, dev)
            except Exception as exc:  # the headline line must still print
                exome = {"error": repr(exc)}
        extras = None
        if world == 1 and not args.no_extra_legs and not args.no_prune:
            try:
                extras = extra_legs(args, torch, dev, data)
            except Exception as exc:
                extras = {"error": repr(exc)}
        h2d = int(hx.numel() * 8 + hw.numel() * 8)
        d2h = int(len(res_h[0]) * (4 + 8 + 8 + 8))
        cfg = wgs_config(args, world, by_arms)
        line = {
            "metric": METRIC, "value": nbins_total / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if by_arms else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "segments": nseg_total, "prune": prune, "numa": numa,
            "gather_ms_per_batch": gather_ms,
            "step_ms_spread": step_spread,   # of this rank's K timed steps (value uses their sum, as the contract says)
            "e2e": {"value": nbins_total / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "gpu_launches_per_step": int(launches) // max(args.steps, 1),
            "roofline": roofline,
            "rooflines": [roof_prep, roof_scan, roof_perm],
            "cpu_baseline": cpu,
            "clocks": clocks,
            "kernel_ms_per_step": per_step,
            "kernel_ms_note": "device time per kernel class from a second pass of the same steps with CUDA events "
                              f"around every class ({ms_prof:.3f} ms/step with the events on)",
            "cbs_stats": stats,
            "cohort": cohort,
            "exome": exome,
        }
        if extras:
            line.update(extras)
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_reference_build(args):
    """C4: pooled reference from S synthetic normals on the C2 bin set -- .cnn files in, reference table
    out (do_reference: TSV parse, per-normal bias correction at window fraction 0.1, per-bin Tukey
    biweight across samples).  Single GPU."""
    import tempfile

    import pandas as pd
    import torch

    from cnvkit_b200 import _lib, reference, synth, tabio

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S = args.samples if args.samples != 1024 else 32
    bins = synth.exome_bins()
    rng = np.random.default_rng(0xC4)
    n = len(bins["start"])
    effect = rng.normal(0, 0.4, n)
    t = bins["is_target"]
    base = pd.DataFrame({"chromosome": bins["chromosome"], "start": bins["start"], "end": bins["end"],
                         "gene": bins["gene"]})
    tmp = tempfile.mkdtemp(prefix="cnvk_c4_")
    tfn, afn, sexes = [], [], {}
    cols = []
    for k in range(S):
        smp = synth.exome_normal(bins, effect, k, male=bool(k % 2))
        cols.append(smp)
        df = base.assign(log2=smp["log2"], depth=smp["depth"], gc=bins["gc"], rmask=bins["rmask"])
        sid = f"N{k:03d}"
        sexes[sid] = not bool(k % 2)
        for sel, kind, dest in ((t, "targetcoverage", tfn), (~t, "antitargetcoverage", afn)):
            path = os.path.join(tmp, f"{sid}.{kind}.cnn")
            tabio.write(df[sel].reset_index(drop=True), path)
            dest.append(path)
    file_bytes = sum(os.path.getsize(f) for f in tfn + afn)
    if args.impl == "reference":
        from oracle import np_oracle as O
        from oracle import pipeline as OP

        cores = len(os.sched_getaffinity(0))
        idx = np.flatnonzero(t)
        c0 = time.perf_counter()
        x = cols[0]["log2"][idx].astype(float)
        for key in (bins["gc"][idx], bins["rmask"][idx]):
            x = OP.center_by_window(x, key, 0.1)
        per_normal = (time.perf_counter() - c0) * (n / len(idx))
        m = 4000
        mat = np.stack([c["log2"][:m] for c in cols])
        c0 = time.perf_counter()
        for c in range(m):
            loc = O.biweight_location(mat[:, c])
            O.biweight_midvariance(mat[:, c], initial=loc)
        per_bin = (time.perf_counter() - c0) / m
        total = S * per_normal + 1.5 * n * per_bin      # log2 location+spread, depth location
        sample = (f"1 normal's window corrections + {m} bins of summarize_info, extrapolated to {S} normals x {n} bins; "
                  "oracle/np_oracle.py + pandas rolling median, 1 core")
        emit(({"impl": "reference", "metric": "normals/s (reference build)", "value": S / total,
                          "unit": "normals/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": 1e3 * total,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": {"workload": f"C4 reference build, {S} normals x {n} bins"},
                          "cpu_baseline": {"value": S / total, "unit": "normals/s", "cores": 1, "kind": "port",
                                           "sample": sample},
                          "e2e": {"value": S / total, "unit": "normals/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": 0}}))
        return
    torch.cuda.set_device(0)
    _lib.load()
    for _ in range(max(1, min(args.warmup, 2))):
        reference.do_reference(tfn[:4], afn[:4], sexes=sexes)
    launches0 = _lib.launch_count()
    times = []
    _lib.prof_enable(True)
    _lib.prof_reset()
    for _ in range(args.steps):
        torch.cuda.synchronize()
        c0 = time.perf_counter()
        ref = reference.do_reference(tfn, afn, sexes=sexes)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - c0)
    prof = _lib.prof_get()
    _lib.prof_enable(False)
    ms = 1e3 * float(np.mean(times))
    per_step = {k: v[0] / args.steps for k, v in prof.items() if v[1] > 0}
    bw_ms = per_step.get("biweight", 0.0)
    bw_bytes = n * (2 * (S + 1) * 8 + S * 8 + 24)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6540.8))
    emit(({
        "metric": "normals/s (reference build)", "value": S / (ms * 1e-3), "unit": "normals/s", "n_gpus": 1,
        "steps": args.steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C4 reference build: {S} normals x {n} bins ({2 * S} .cnn files, {file_bytes / 1e6:.0f} MB "
                               f"of TSV), do_reference with GC/RepeatMasker/edge centring at fraction 0.1",
                   "l2": "every normal's columns are fresh host data; the [S+1, n] matrices exceed L2"},
        "e2e": {"value": S / (ms * 1e-3), "unit": "normals/s", "h2d_bytes_per_step": int(S * n * 16),
                "d2h_bytes_per_step": int(n * 24)},
        "gpu_launches": int((_lib.launch_count() - launches0) / args.steps),
        "roofline": {"kernel": "biweight_columns_kernel (summarize_info)", "bound": "hbm",
                     "achieved": bw_bytes / (bw_ms * 1e-3) / 1e9 if bw_ms else None, "peak": hbm, "unit": "GB/s",
                     "frac": (bw_bytes / (bw_ms * 1e-3) / 1e9 / hbm) if bw_ms else None, "traffic": None,
                     "algorithmic": "read 2(S+1)+S doubles per bin (log2 twice: location, midvariance; depth once), "
                                    "write 24 B per bin"},
        "kernel_ms_per_step": per_step, "rows": int(len(ref)),
        "note": "e2e includes TSV parsing (native reader) of every file; value == e2e (no resident variant)"}))


def cohort_core(args, torch, dist, dev, world, rank, local_rank, samples, steps, quiet=False):
    """C5: `samples` exome tumours fixed against one pooled reference and CBS-segmented, sharded by sample over
    the ranks in blocks (strong scaling: the cohort is fixed), .cns rows gathered on rank 0 with one NCCL
    collective per step.  Returns the result object on rank 0 (None elsewhere).  The process group, if any,
    is the caller's."""
    from cnvkit_b200 import _lib, cohort, shard

    n_distinct = 16
    bins, ref, smp, cols, (T, A, R) = exome_world(n_distinct)
    panel = cohort.Panel(T, A, R)
    owner = shard.block_owner(samples, world)
    mine = np.flatnonzero(owner == rank)
    pinned = [tuple({k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in part.items()} for part in c)
              for c in cols]
    resident = [tuple({k: v.to(dev) for k, v in part.items()} for part in c) for c in pinned]
    stream = torch.cuda.current_stream()
    gatherer = shard.TableGatherer(ncols=7, capacity=max(len(mine), 1) * 512 + 16, device=dev if world > 1 else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(source, which):
        work = [source[i % n_distinct] for i in which]
        if source is pinned:
            work = [tuple({k: v.numpy() for k, v in part.items()} for part in c) for c in work]
        res = panel.run(work, method="cbs", batch=args.cohort_batch, workers=args.workers, sample_ids=[f"S{i}" for i in which])
        names = ("start", "end", "log2", "probes", "weight", "depth")
        ids, parts = [], [[] for _ in names]
        for i, (_c, cns) in zip(which, res):
            d = cns.data
            ids.append(np.full(len(d), i))
            for j, c in enumerate(names):
                parts[j].append(d[c].to_numpy())
        if ids:     # one staged block per step instead of one per sample
            gatherer.append([np.concatenate(ids)] + [np.concatenate(p) for p in parts])
        full = gatherer.flush(dst=0)
        return len(full) if full is not None else 0

    def timed(source, nsteps):
        out = []
        n_rows = 0
        for _ in range(nsteps):
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            n_rows = step(source, mine)
            b.record(stream)
            barrier()
            out.append(a.elapsed_time(b))
        return out, n_rows

    # warm-up on a slice of this rank's share (module load, allocator pools, NCCL communicator)
    for _ in range(2):
        step(resident, mine[: min(len(mine), 32)])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = _lib.launch_count()
    t0 = time.time()
    ev, n_rows = timed(resident, steps)
    t1 = time.time()
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ev_h, _ = timed(pinned, 1)

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = max_over_ranks(float(np.mean(ev)))
    ms_h = max_over_ranks(float(np.mean(ev_h)))
    if rank != 0:
        return None
    per_sample_bytes = sum(v.numel() * 8 for part in pinned[0] for v in part.values())
    return {
        "metric": EXOME_METRIC, "value": samples / (ms * 1e-3), "unit": "samples/s", "n_gpus": world,
        "steps": steps, "warmup": 2, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C5 cohort of {samples} C2 exome tumours ({n_distinct} distinct synthetic samples "
                               f"cycled), fix + CBS, {args.cohort_batch} samples per segmentation call, {args.workers} host threads / streams per rank",
                   "samples": int(samples),
                   "sharding": f"samples over {world} rank(s) in blocks; one NCCL all-gather of the .cns rows per step",
                   "l2": "inputs of one step (>= 400 MB of columns per rank) exceed the 126 MB L2",
                   "segment_rows": n_rows},
        "e2e": {"value": samples / (ms_h * 1e-3), "unit": "samples/s", "ms_per_step": ms_h,
                "h2d_bytes_per_step": int(per_sample_bytes * len(mine)), "d2h_bytes_per_step": int(n_rows * 7 * 8)},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": None, "cpu_baseline": None,
        "note": "a step = the whole cohort; host-side table assembly (pandas) is inside the timed region; strong-scaling "
                "efficiency = value(N) / (N x value(1)) with the same cohort size at every N"}


def run_cohort(args):
    """--workload cohort: the C5 leg alone (it is also part of the default run, key `cohort`)."""
    import torch
    import torch.distributed as dist

    from cnvkit_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank != 0:
            return
        cores = len(os.sched_getaffinity(0))
        bins, ref, samples, cols, _ = exome_world(2)
        times = []
        for step in range(args.warmup + args.steps):
            c = exome_cpu_baseline(bins, ref, samples[step % 2], cores, methods=("cbs",))
            if step >= args.warmup:
                times.append(c["fix_s"] + c["cbs_s"])
        ms = 1e3 * float(np.mean(times))
        sample = "1 sample of the cohort per step, oracle/pipeline.py (pandas rolling median, C CBS with OpenMP over arms)"
        emit(({
            "impl": "reference", "metric": EXOME_METRIC, "value": 1e3 / ms, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C5 cohort of {args.samples} C2 exome tumours, fix + CBS", "sample": sample},
            "cpu_baseline": {"value": 1e3 / ms, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": 1e3 / ms, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    bind_to_gpu_numa(torch, local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    line = cohort_core(args, torch, dist, dev, world, rank, local_rank, samples=args.samples, steps=args.steps)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.workload == "cohort":
        run_cohort(args)
    elif args.workload == "reference":
        run_reference_build(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

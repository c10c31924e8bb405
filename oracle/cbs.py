"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/cbs_oracle.c (see its header:
PARITY UNPINNED restatement of DNAcopy weighted CBS)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_double, c_int, c_int64, c_longlong, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


class Params(Structure):
    _fields_ = [
        ("alpha", c_double),
        ("nperm", c_int),
        ("min_width", c_int),
        ("kmax", c_int),
        ("nmin", c_int),
        ("eta", c_double),
        ("hybrid", c_int),
        ("ngrid", c_int),
        ("tol", c_double),
        ("seed", c_uint64),
        ("prune", c_int),
    ]


class Stats(Structure):
    _fields_ = [
        (k, c_longlong)
        for k in ("nodes", "obs_arcs", "perm_arcs", "perm_tests", "perms_run", "tperm_tests")
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def make_params(alpha=1e-4, nperm=10000, min_width=2, kmax=25, nmin=200, eta=0.05, hybrid=True, seed=0xA5EED,
                prune=False):
    """prune=True: block-pruned search for the observed maximum (what DNAcopy's wtmaxo does); same
    result as the exhaustive default, used for the CPU baseline timings."""
    return Params(alpha, nperm, min_width, kmax, nmin, eta, int(hybrid), 100, 1e-6, seed, int(prune))


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.check_call(["make", "-C", _HERE])
        L = ctypes.CDLL(_SO)
        L.cbso_tailp.restype = c_double
        L.cbso_tailp.argtypes = [c_double, c_double, c_int, c_int, c_double]
        L.cbso_getbdry.restype = c_int
        L.cbso_getbdry.argtypes = [c_double, c_int, c_int, c_void_p]
        L.cbso_prp.restype = c_uint32
        L.cbso_prp.argtypes = [c_uint64, c_uint32, c_uint32, c_uint32, c_uint32, c_uint32, c_uint32]
        L.cbso_find_cpt.restype = c_int
        L.cbso_find_cpt.argtypes = [c_void_p, c_void_p, c_int, c_int, POINTER(Params), c_void_p, c_void_p,
                                    POINTER(c_double), POINTER(Stats)]
        L.cbso_segment_arm.restype = c_int
        L.cbso_segment_arm.argtypes = [c_void_p, c_void_p, c_int, POINTER(Params), c_void_p, c_void_p, c_void_p,
                                       POINTER(Stats)]
        L.cbso_segment.restype = c_int
        L.cbso_segment.argtypes = [c_void_p, c_void_p, c_void_p, c_int, POINTER(Params), c_void_p, c_void_p,
                                   c_void_p, POINTER(Stats), c_int]
        L.cbso_perm_pvalue_flank.restype = c_double
        L.cbso_perm_pvalue_flank.argtypes = [c_void_p, c_void_p, c_int, c_int, POINTER(Params), c_uint32, c_uint32,
                                             c_uint32]
        L.cbso_prp_fill.restype = None
        L.cbso_prp_fill.argtypes = [c_uint64, c_uint32, c_uint32, c_uint32, c_uint32, c_uint32, c_uint32, c_void_p]
        L.cbso_prp_at.restype = None
        L.cbso_prp_at.argtypes = [c_uint64, c_uint32, c_uint32, c_uint32, c_uint32, c_uint32, c_uint32, c_void_p,
                                  c_uint32, c_void_p]
        L.cbso_node_diag.restype = c_int
        L.cbso_node_diag.argtypes = [c_void_p, c_void_p, c_int, c_int, POINTER(Params), c_int, c_void_p, c_void_p]
        L.cbso_smooth_cna.restype = c_int
        L.cbso_smooth_cna.argtypes = [c_void_p, c_void_p, c_int, c_int, c_double, c_double, c_double, c_void_p, c_void_p]
        _lib = L
    return _lib


def smooth_cna(x, arm_offsets, smooth_region=10, outlier_sd_scale=4.0, smooth_sd_scale=2.0, trim=0.025):
    """DNAcopy smooth.CNA of every arm (restated; see cbs_oracle.c).  Returns (smoothed, trimmed_sd_per_arm)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    off = np.ascontiguousarray(arm_offsets, dtype=np.int64)
    out = np.empty_like(x)
    sd = np.zeros(len(off) - 1, dtype=np.float64)
    lib().cbso_smooth_cna(x.ctypes.data, off.ctypes.data, len(off) - 1, smooth_region, outlier_sd_scale, smooth_sd_scale,
                          trim, out.ctypes.data, sd.ctypes.data)
    return out, sd


def tailp(b, delta, m, ngrid=100, tol=1e-6):
    return lib().cbso_tailp(b, delta, m, ngrid, tol)


def getbdry(eta, nperm, max_ones):
    out = np.zeros(max_ones * (max_ones + 1) // 2, dtype=np.int32)
    lib().cbso_getbdry(eta, nperm, max_ones, out.ctypes.data)
    return out


def sbdry_for(params):
    max_ones = int(np.floor(params.nperm * params.alpha)) + 1
    return getbdry(params.eta, params.nperm, max_ones)


def prp(seed, lo, hi, tid, perm, n):
    L = lib()
    return np.array([L.cbso_prp(seed, lo, hi, tid, perm, n, t) for t in range(n)], dtype=np.int64)


def prp_fill(seed, lo, hi, tid, perm0, nperms, n):
    """Permutations perm0..perm0+nperms-1 of the test (lo, hi, tid) as an int64 [nperms, n] matrix."""
    out = np.zeros((nperms, n), dtype=np.uint32)
    lib().cbso_prp_fill(seed, lo, hi, tid, perm0, nperms, n, out.ctypes.data)
    return out.astype(np.int64)


def prp_at(seed, lo, hi, tid, perm0, nperms, n, positions):
    """Values of permutations perm0..perm0+nperms-1 at the given positions: int64 [nperms, len(positions)]."""
    pos = np.ascontiguousarray(positions, dtype=np.uint32)
    out = np.zeros((nperms, len(pos)), dtype=np.uint32)
    lib().cbso_prp_at(seed, lo, hi, tid, perm0, nperms, n, pos.ctypes.data, len(pos), out.ctypes.data)
    return out.astype(np.int64)


def node_diag(x, w, lo, hi, params, nperm_run=None):
    """One node without early stopping (statistical tests): dict with the observed sqrt(T^2) (`ostat1`), the
    value permutations are compared with (`ostat`), the arc, `tailp` / `delta` (hybrid nodes), the number of
    the first `nperm_run` permutations that reach `ostat` (`nrej`) and the two flank p-values with the
    permutations forced (`flank_p`, -1 when the arc touches an end of the node)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    nrun = params.nperm if nperm_run is None else int(nperm_run)
    out = np.zeros(6, dtype=np.float64)
    iout = np.zeros(5, dtype=np.int32)
    lib().cbso_node_diag(x.ctypes.data, w.ctypes.data, lo, hi, params, nrun, out.ctypes.data, iout.ctypes.data)
    return dict(ostat1=out[0], tailp=out[1], delta=out[2], ostat=out[3], flank_p=(out[4], out[5]),
                arc=(int(iout[0]), int(iout[1])), nrej=int(iout[2]), hybrid=bool(iout[3]), flat=bool(iout[4]),
                nperm_run=nrun)


def find_cpt(x, w, lo, hi, params, sbdry=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    if sbdry is None:
        sbdry = sbdry_for(params)
    icpt = np.zeros(2, dtype=np.int32)
    ostat1 = c_double(0)
    st = Stats()
    ncpt = lib().cbso_find_cpt(x.ctypes.data, w.ctypes.data, lo, hi, params, sbdry.ctypes.data, icpt.ctypes.data,
                               ostat1, st)
    return ncpt, icpt[:ncpt].tolist(), ostat1.value, st.as_dict()


def segment(x, w, arm_offsets, params, nthreads=0):
    """Returns (seg_arm, seg_lo, seg_hi, seg_mean, stats) with global bin indices."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    off = np.ascontiguousarray(arm_offsets, dtype=np.int64)
    n_arms = len(off) - 1
    n = len(x)
    seg_end = np.zeros(n, dtype=np.int64)
    seg_mean = np.zeros(n, dtype=np.float64)
    nseg = np.zeros(n_arms, dtype=np.int32)
    st = Stats()
    lib().cbso_segment(x.ctypes.data, w.ctypes.data, off.ctypes.data, n_arms, params, seg_end.ctypes.data,
                       seg_mean.ctypes.data, nseg.ctypes.data, st, nthreads)
    arm, lo, hi, mean = [], [], [], []
    for a in range(n_arms):
        s0 = off[a]
        prev = off[a]
        for s in range(nseg[a]):
            arm.append(a)
            lo.append(prev)
            hi.append(seg_end[s0 + s])
            mean.append(seg_mean[s0 + s])
            prev = seg_end[s0 + s]
    return (np.asarray(arm, dtype=np.int32), np.asarray(lo, dtype=np.int64), np.asarray(hi, dtype=np.int64),
            np.asarray(mean, dtype=np.float64), st.as_dict())

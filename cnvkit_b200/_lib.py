"""ctypes binding of libcnvkit_b200.so (the C ABI declared in include/cnvkit_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
caller gets an exception.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C cnvkit_b200/csrc``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint32, c_uint64, c_ulonglong, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcnvkit_b200.so")

_lib = None

c_double_p = POINTER(c_double)
c_int64_p = POINTER(c_int64)
c_int32_p = POINTER(c_int32)
c_uint32_p = POINTER(c_uint32)

# name -> (restype, argtypes); must list every symbol include/cnvkit_b200.h declares.
SIGNATURES = {
    "cnvk_last_error": (c_char_p, []),
    "cnvk_abi_version": (c_int, []),
    "cnvk_launch_count": (c_ulonglong, []),
    "cnvk_device_info": (c_int, [c_int, POINTER(c_int), POINTER(c_uint64), POINTER(c_uint64)]),
    "cnvk_fp64_peak": (c_int, [c_int, c_int, c_void_p]),
    "cnvk_nvtx_enable": (None, [c_int]),
    "cnvk_prof_enable": (None, [c_int]),
    "cnvk_prof_reset": (None, []),
    "cnvk_prof_count": (c_int, []),
    "cnvk_prof_name": (c_char_p, [c_int]),
    "cnvk_prof_get": (c_int, [c_void_p, c_void_p, c_int]),
    "cnvk_rolling_median_dev": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "cnvk_rolling_median": (c_int, [c_void_p, c_int64, c_int64, c_void_p]),
    "cnvk_rolling_quantile_dev": (c_int, [c_void_p, c_int64, c_int64, c_double, c_void_p, c_void_p]),
    "cnvk_rolling_quantile": (c_int, [c_void_p, c_int64, c_int64, c_double, c_void_p]),
    "cnvk_argsort_f64_dev": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "cnvk_fix_mask_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "cnvk_center_all_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p,
                                    c_void_p]),
    "cnvk_center_by_window_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "cnvk_center_by_window": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64]),
    "cnvk_edge_bias_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "cnvk_fix_group_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64,
                                   c_void_p, c_int64, c_char_p, c_void_p, c_void_p]),
    "cnvk_bias_correct_logr_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p,
                                           c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                                           c_int32, c_void_p, c_void_p]),
    "cnvk_trait_order_dev": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "cnvk_center_by_order_dev": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    "cnvk_fix_group_ordered_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p,
                                           c_void_p, c_int32, c_int32, c_int64, c_void_p, c_void_p]),
    "cnvk_fix_finish_dev": (c_int, [c_void_p] * 9 + [c_int64, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "cnvk_segmented_median_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "cnvk_biweight": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "cnvk_biweight_columns_dev": (c_int, [c_void_p, c_int32, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "cnvk_biweight_columns": (c_int, [c_void_p, c_int32, c_int64, c_int32, c_void_p, c_void_p]),
    "cnvk_outlier_mask_dev": (c_int, [c_void_p, c_void_p, c_int32, c_double, c_double, c_void_p, c_void_p]),
    "cnvk_round_sig_dev": (c_int, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "cnvk_segment_bin_sums_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                          c_void_p]),
    "cnvk_segment_bin_sums_masked_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                                 c_void_p, c_void_p, c_void_p]),
    "cnvk_cbs_segment_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cnvk_cbs_segment": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "cnvk_haar_segment_sigma": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_double, c_void_p, c_int64, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    "cnvk_haar_sigma": (c_int, [c_void_p, c_void_p, c_int32, c_void_p]),
    "cnvk_haar_segment_means": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p]),
    "cnvk_haar_segment_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_double, c_int64, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "cnvk_haar_segment": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_double, c_int64, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "cnvk_cbs_getbdry": (c_int, [c_double, c_int32, c_int32, c_void_p]),
    "cnvk_cbs_input_filter_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cnvk_panel_sizeof": (c_int64, []),
    "cnvk_panel_prepare_dev": (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_double, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "cnvk_panel_transfer_dev": (c_int, [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "cnvk_coverage_finalize_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p, c_void_p, c_void_p]),
    "cnvk_coverage_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p, c_void_p]),
    "cnvk_smooth_cna_dev": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_double, c_double, c_double, c_void_p,
                                    c_void_p, c_void_p]),
    "cnvk_smooth_cna": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_double, c_double, c_double, c_void_p, c_void_p]),
    "cnvk_cbs_node_diag": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "cnvk_fasta_gc_lo_dev": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "cnvk_fasta_gc_lo": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "cnvk_bootstrap_ci_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    "cnvk_bootstrap_ci": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                  c_void_p]),
    "cnvk_tsv_open": (c_int, [c_char_p, c_void_p]),
    "cnvk_tsv_close": (None, [c_void_p]),
    "cnvk_tsv_shape": (c_int, [c_void_p, c_void_p, c_void_p]),
    "cnvk_tsv_colname": (c_char_p, [c_void_p, c_int32]),
    "cnvk_tsv_col_kind": (c_int, [c_void_p, c_int32]),
    "cnvk_tsv_read_f64": (c_int, [c_void_p, c_int32, c_void_p]),
    "cnvk_tsv_read_i64": (c_int, [c_void_p, c_int32, c_void_p]),
    "cnvk_tsv_read_str": (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_int64, c_void_p]),
    "cnvk_tsv_read_str_dict": (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "cnvk_tsv_hash_cols": (c_uint64, [c_void_p, c_void_p, c_int32]),
    "cnvk_ordered_unique_ranges": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_int64]),
    "cnvk_join_names_ranges": (c_int, [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_int64]),
    "cnvk_tsv_write": (c_int, [c_char_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int64]),
}


class CbsParams(ctypes.Structure):
    """cnvk_cbs_params (include/cnvkit_b200.h)."""

    _fields_ = [
        ("alpha", c_double),
        ("nperm", c_int32),
        ("min_width", c_int32),
        ("kmax", c_int32),
        ("nmin", c_int32),
        ("eta", c_double),
        ("hybrid", c_int32),
        ("prune", c_int32),
        ("seed", c_uint64),
    ]


class CbsStats(ctypes.Structure):
    """cnvk_cbs_stats (include/cnvkit_b200.h)."""

    _fields_ = [
        (k, c_int64)
        for k in ("nodes", "levels", "obs_subtiles_total", "obs_subtiles_scanned", "obs_arcs", "perm_arcs",
                  "perm_nodes", "perms_run", "flank_perm_tests", "prep_bins")
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class TsvDict(ctypes.Structure):
    """cnvk_tsv_dict (include/cnvkit_b200.h)."""

    _fields_ = [("offsets", POINTER(c_int64)), ("blob", c_char_p), ("n", c_int64)]


class BootParams(ctypes.Structure):
    """cnvk_boot_params (include/cnvkit_b200.h)."""

    _fields_ = [
        ("alpha", c_double),
        ("bootstraps", c_int32),
        ("smooth_upto", c_int32),
        ("pcg_state", c_uint64 * 2),
        ("pcg_inc", c_uint64 * 2),
        ("noise_seed", c_uint64),
    ]


class PanelGroup(ctypes.Structure):
    """cnvk_panel_group (include/cnvkit_b200.h)."""

    _fields_ = [
        ("n0", c_int64),
        ("n", c_int64),
        ("d_keep_idx", c_void_p),
        ("d_auto", c_void_p),
        ("d_chrom_off", c_void_p),
        ("nchrom", c_int32),
        ("skip_low", c_int32),
        ("d_order", c_void_p * 3),
        ("wing", c_int64),
    ]


class PanelDesc(ctypes.Structure):
    """cnvk_panel (include/cnvkit_b200.h)."""

    _fields_ = [
        ("grp", PanelGroup * 2),
        ("n", c_int64),
        ("d_order", c_void_p),
        ("d_ref_log2", c_void_p),
        ("d_ref_spread", c_void_p),
        ("d_start", c_void_p),
        ("d_end", c_void_p),
        ("d_chrom_id", c_void_p),
        ("d_is_anti", c_void_p),
        ("d_auto", c_void_p),
        ("d_chrom_off", c_void_p),
        ("nchrom", c_int32),
        ("n_arms", c_int32),
        ("d_arm_id", c_void_p),
        ("arm_off", c_void_p),
    ]


PANEL_NAN_INPUT, PANEL_LOW_COVERAGE, PANEL_NONFINITE, PANEL_QUANTISER_RANGE = 1, 2, 4, 8


class CnvkError(RuntimeError):
    """A libcnvkit_b200 call failed (CUDA error or bad arguments)."""


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: the CUDA library is not built. "
                "Run `make -C cnvkit_b200/csrc` (needs nvcc); there is no CPU fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cnvk_last_error().decode(errors="replace")
        raise CnvkError(f"{what} failed (rc={rc}): {msg}")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)


def launch_count() -> int:
    return int(load().cnvk_launch_count())


def prof_enable(on: bool = True) -> None:
    load().cnvk_prof_enable(int(on))


def prof_reset() -> None:
    load().cnvk_prof_reset()


def prof_get() -> dict:
    """{class name: (milliseconds, timed scopes)} accumulated since the last reset."""
    lib = load()
    n = lib.cnvk_prof_count()
    ms = np.zeros(n, dtype=np.float64)
    cnt = np.zeros(n, dtype=np.uint64)
    lib.cnvk_prof_get(ms.ctypes.data, cnt.ctypes.data, n)
    return {lib.cnvk_prof_name(k).decode(): (float(ms[k]), int(cnt[k])) for k in range(n)}


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def ptr(a: np.ndarray) -> int:
    """Host pointer of a contiguous numpy array (0 for None)."""
    return 0 if a is None else a.ctypes.data


def stream_ptr(stream=None) -> int:
    """cudaStream_t of a torch stream (or torch's current stream), as an int."""
    import torch

    if stream is None:
        stream = torch.cuda.current_stream()
    return int(stream.cuda_stream)

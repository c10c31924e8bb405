"""Coverage finalisation on the GPU -- the numeric tail of cnvlib/coverage.py.

`interval_coverages_pileup` (coverage.py:615-643) and `interval_coverages_bedgraph` (:395-409) both end the same
way: a table of per-bin base counts (from `samtools bedcov` / a bedGraph, both outside this path) gets the `gene`
placeholder, `depth = basecount / span` and `log2 = log2(depth)` with NULL_LOG2_COVERAGE for empty bins.  That tail
is `finalize_basecounts`; BAM / bedGraph reading stays with the reference.
"""
from __future__ import annotations

import numpy as np

from . import _lib, params


def depth_log2(basecount, start, end):
    """(depth, log2) columns of coverage.py:637-643 for arrays of base counts and bin coordinates."""
    b = _lib.f64(basecount)
    s = _lib.i64(start)
    e = _lib.i64(end)
    if not (len(b) == len(s) == len(e)):
        raise ValueError("basecount / start / end lengths differ")
    depth = np.empty(len(b), dtype=np.float64)
    log2 = np.empty(len(b), dtype=np.float64)
    _lib.call("cnvk_coverage_finalize", _lib.ptr(b), _lib.ptr(s), _lib.ptr(e), len(b), float(params.NULL_LOG2_COVERAGE),
              _lib.ptr(depth), _lib.ptr(log2))
    return depth, log2


def finalize_basecounts(table):
    """The tail of interval_coverages_pileup / _bedgraph: `table` is a DataFrame with chromosome, start, end,
    basecount (and optionally gene); returns it with `gene` filled, `depth` and `log2` added (the bedGraph variant
    additionally drops `basecount`, :412 -- pass drop_basecount via `.drop(columns=...)` if wanted)."""
    table = table.copy()
    if "gene" in table:
        table["gene"] = table["gene"].fillna("-")
    else:
        table["gene"] = "-"
    depth, log2 = depth_log2(table["basecount"].to_numpy(dtype=np.float64), table["start"].to_numpy(),
                             table["end"].to_numpy())
    return table.assign(depth=depth, log2=log2)

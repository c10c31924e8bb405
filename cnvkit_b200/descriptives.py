"""Robust statistics on the GPU -- host mirror of cnvlib/descriptives.py (biweight only)."""
from __future__ import annotations

import numpy as np

from . import _lib


def _biweight(a, initial):
    a = _lib.f64(np.asarray(a, dtype=float))
    out = np.zeros(2)
    init = None if initial is None else np.array([float(initial)])
    _lib.call("cnvk_biweight", _lib.ptr(a), len(a), _lib.ptr(init), _lib.ptr(out))
    return out


def biweight_location(a, initial=None):
    """cnvlib/descriptives.py:96-142 (c=6, epsilon=1e-3, max_iter=5)."""
    if initial is not None:
        raise NotImplementedError("biweight_location(initial=...) is not used on the hot path")
    return _biweight(a, None)[0]


def biweight_midvariance(a, initial=None):
    """cnvlib/descriptives.py:188-222 (c=9, epsilon=1e-3)."""
    return _biweight(a, initial)[1]

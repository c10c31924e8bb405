"""Signal smoothers on the hot path -- host mirror of cnvlib/smoothing.py.

Only the window arithmetic (``_width2wing``, short-input guards) lives here; the order
statistics themselves run in csrc/rolling.cu.
"""
from __future__ import annotations

import math

import numpy as np

from . import _lib


def _width2wing(width, n: int, min_wing: int = 3) -> int:
    """Half-window from a fractional or absolute width (cnvlib/smoothing.py:53-69)."""
    if 0 < width < 1:
        wing = math.ceil(n * width * 0.5)
    elif width >= 2 and int(width) == width:
        width = min(width, n - 1)
        wing = int(width // 2)
    else:
        raise ValueError(
            "width must be either a fraction between 0 and 1 "
            + f"or an integer greater than 1 (got {width})"
        )
    wing = max(wing, min_wing)
    wing = min(wing, n - 1)
    assert wing >= 1, f"Wing must be at least 1 (got {wing})"
    return wing


def rolling_median(x, width) -> np.ndarray:
    """Rolling median with mirrored edges (cnvlib/smoothing.py:77-87), on the GPU."""
    x = _lib.f64(np.asarray(x, dtype=float))
    if len(x) < 2:
        return x.copy()
    wing = _width2wing(width, len(x))
    out = np.empty_like(x)
    _lib.call("cnvk_rolling_median", _lib.ptr(x), len(x), wing, _lib.ptr(out))
    return out


def rolling_quantile(x, width, quantile) -> np.ndarray:
    """Rolling quantile with mirrored edges (cnvlib/smoothing.py:135-141), on the GPU."""
    x = _lib.f64(np.asarray(x, dtype=float))
    if len(x) < 2:
        return x.copy()
    wing = _width2wing(width, len(x))
    out = np.empty_like(x)
    _lib.call("cnvk_rolling_quantile", _lib.ptr(x), len(x), wing, float(quantile), _lib.ptr(out))
    return out

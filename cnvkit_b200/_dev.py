"""Device-memory plumbing: torch tensors hold HBM buffers, the C ABI gets raw pointers."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("cnvkit_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def up(a, dtype=None):
    """numpy -> CUDA tensor (contiguous)."""
    a = np.ascontiguousarray(a if dtype is None else np.asarray(a, dtype=dtype))
    if not a.flags.writeable:
        a = a.copy()
    return torch.from_numpy(a).to(device(), non_blocking=False)


def empty(n, dtype):
    return torch.empty(int(n), dtype=dtype, device=device())


def ptr(t):
    return 0 if t is None else t.data_ptr()


def stream():
    return _lib.stream_ptr()


_perm_cache = {}


def shuffle_permutation(n: int):
    """np.random.default_rng(0xA5EED).permutation(n) (cnvlib/fix.py:404-405) as a cached CUDA
    uint32 tensor -- it depends only on n."""
    key = (int(n), torch.cuda.current_device())
    t = _perm_cache.get(key)
    if t is None:
        perm = np.random.default_rng(0xA5EED).permutation(int(n)).astype(np.uint32)
        t = torch.from_numpy(perm.view(np.int32)).to(device())
        if len(_perm_cache) > 64:
            _perm_cache.clear()
        _perm_cache[key] = t
    return t

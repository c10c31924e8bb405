"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Loads the *unmodified* reference (etal/cnvkit, mounted read-only at
/root/reference) in the authoring container so that golden vectors can be
generated from the reference's own code.  The reference
needs four third-party modules that are absent here and unused on the hot path
(Bio.File.as_handle, bioframe, statsmodels lowess, pyfaidx); tiny stubs for them
live in ``oracle/refshim_stubs``.  ``cnvlib/__init__.py`` is bypassed with a
namespace shim because it pulls matplotlib through ``commands.py``.

/root/reference does not exist on the GPU box: nothing under ``tests/ -m gpu`` or
``__graft_entry__.smoke()`` may call this module.  Its consumers are ``oracle/make_golden.py``, the
authoring-container-only tests marked ``needs_reference`` and ``bench.py``'s ``cpu_baseline`` legs, which time
the reference's own Python (the ``pip install --target baseline/_ref`` copy, git-ignored) beside the port.
"""
from __future__ import annotations

import os
import sys
import types

# /root/reference in the authoring container; on the GPU box the pip-installed copy under baseline/_ref
# (git-ignored, travels with the gpurun snapshot) -- bench.py's cpu_baseline legs time it there
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = [os.environ.get("CNVKIT_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
REFERENCE_ROOT = next((c for c in _CANDIDATES if c and os.path.isdir(os.path.join(c, "cnvlib"))), "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refshim_stubs")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cnvlib"))


def load():
    """Return a namespace with the reference's hot-path modules."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    if "cnvlib" not in sys.modules:
        pkg = types.ModuleType("cnvlib")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "cnvlib")]
        sys.modules["cnvlib"] = pkg
    from cnvlib import (  # noqa: E402
        cnary,
        descriptives,
        fix,
        importers,
        params,
        reference,
        segmentation,
        smoothing,
    )
    from cnvlib.cmdutil import read_cna
    from cnvlib.segmentation import haar
    from skgenome import tabio

    ns = types.SimpleNamespace(
        cnary=cnary,
        descriptives=descriptives,
        fix=fix,
        importers=importers,
        params=params,
        reference=reference,
        segmentation=segmentation,
        smoothing=smoothing,
        haar=haar,
        read_cna=read_cna,
        tabio=tabio,
        root=REFERENCE_ROOT,
    )
    return ns

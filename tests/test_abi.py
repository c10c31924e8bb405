"""CPU-only: the C-ABI library loads and exports every symbol include/cnvkit_b200.h declares
(no compute calls -- there is no GPU here), and the ctypes table covers the header."""
import os
import re

from cnvkit_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "cnvkit_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cnvk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 10
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_ctypes_table_matches_header():
    assert sorted(_lib.SIGNATURES) == header_symbols()


def test_abi_version_and_host_only_calls():
    lib = _lib.load()
    assert lib.cnvk_abi_version() == 3
    from cnvkit_b200.segmentation import cbs

    b = cbs.getbdry(0.05, 10000, 2)  # host-only entry point
    assert b[0] == 9500 and len(b) == 3


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    import pytest

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.load()


def test_panel_descriptor_layout_matches_the_library():
    """The ctypes mirror of cnvk_panel (device pointers handed to the batch entry points) must have the layout the
    library was compiled with; the size is the canary for a reordered or resized member."""
    import ctypes

    from cnvkit_b200 import _lib

    assert int(_lib.load().cnvk_panel_sizeof()) == ctypes.sizeof(_lib.PanelDesc)
    assert ctypes.sizeof(_lib.PanelGroup) * 2 == _lib.PanelDesc.n.offset
    assert _lib.PanelDesc.arm_off.offset + 8 == ctypes.sizeof(_lib.PanelDesc)


def test_cohort_chunk_plan_covers_every_sample_once():
    from cnvkit_b200.cohort import _chunk_plan

    for n, batch, workers in [(0, 8, 4), (1, 8, 1), (7, 8, 2), (8, 8, 4), (9, 8, 4), (1024, 8, 4), (5, 16, 1)]:
        chunks = _chunk_plan(n, batch, workers)
        flat = [k for c in chunks for k in c]
        assert flat == list(range(n))
        assert all(1 <= len(c) <= batch for c in chunks)

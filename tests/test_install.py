"""CPU-only: cnvkit_b200.install rebinds the reference's seams and
forwards out-of-scope options to the reference's own callables.  Needs the reference tree, so it
runs in the authoring container only."""
import sys

import pytest

from cnvkit_b200 import install as I


@pytest.mark.needs_reference
def test_install_rebinds_and_restores():
    from oracle import refshim

    R = refshim.load()
    orig = (R.fix.do_fix, R.reference.do_reference, R.segmentation.do_segmentation, R.smoothing.rolling_median)
    try:
        I.install()
        assert I.installed()
        import cnvkit_b200.smoothing as S

        assert R.smoothing.rolling_median is S.rolling_median
        for fn, o in zip((R.fix.do_fix, R.reference.do_reference, R.segmentation.do_segmentation), orig):
            assert fn is not o and fn.__wrapped__ is o
        I.install()  # idempotent
    finally:
        I.uninstall()
    assert (R.fix.do_fix, R.reference.do_reference, R.segmentation.do_segmentation,
            R.smoothing.rolling_median) == orig
    assert not I.installed()


def test_routing_predicates():
    assert I._segment_needs_reference((None, "hmm"), {})
    assert I._segment_needs_reference((None, "cbs"), {"smooth_cbs": True})
    assert not I._segment_needs_reference((None, "cbs"), {"threshold": 1e-4})
    assert not I._segment_needs_reference((None,), {"method": "haar"})
    assert I._fix_needs_reference((1, 2, 3), {"bias_smoother": "loess"})
    assert I._fix_needs_reference((1, 2, 3, None, True, True, True, True), {})
    assert not I._fix_needs_reference((1, 2, 3), {})
    assert I._reference_needs_reference((["a"],), {"do_cluster": True})
    assert not I._reference_needs_reference((["a"],), {})


@pytest.mark.needs_reference
def test_out_of_scope_options_reach_the_reference():
    from oracle import refshim

    R = refshim.load()
    calls = []
    real = R.segmentation.do_segmentation
    R.segmentation.do_segmentation = lambda *a, **k: calls.append((a, k)) or "theirs"
    try:
        I.install()
        assert R.segmentation.do_segmentation("cnarr", "hmm") == "theirs"
        assert calls and calls[0][0][1] == "hmm"
    finally:
        I.uninstall()
        R.segmentation.do_segmentation = real

"""GPU: cnvkit_b200.install.install() beneath the REAL cnvlib (the pip-installed copy under baseline/_ref, which
travels with the snapshot; /root/reference in the authoring container): the reference's own callables
`cnvlib.fix.do_fix` / `cnvlib.segmentation.do_segmentation` (the ones commands.py:1177, :1290 and batch.py:524, :547
reach through module attributes) are called with the reference's own CopyNumArray objects, once un-installed (pure
reference, CPU) and once installed (B200), and the tables are compared: fix bit-exact log2 / 1e-10 weights, Haar
identical segments, CBS structural (test/test_r.py:234-310 -- R is absent, so the reference's own CBS cannot run)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    from oracle import refshim

    if not refshim.available():
        pytest.skip("reference tree not present (baseline/_ref is git-ignored; build() installs it where /root/reference exists)")
    return refshim.load()


def _tables(R):
    from cnvkit_b200 import synth

    bins = synth.exome_bins(n_targets=9000, n_antitargets=3000, seed=0x1D)
    ref = synth.exome_reference(bins)
    smp = synth.exome_sample(bins, ref, seed=5)
    T, A, Rf = synth.exome_tables(bins, ref, smp, sample_id="S")
    mk = lambda t: R.cnary.CopyNumArray(t.data.copy(), {"sample_id": "S"})
    return mk(T), mk(A), mk(Rf)


def test_install_under_real_cnvlib(R):
    from cnvkit_b200 import install as I

    T, A, Rf = _tables(R)
    assert not I.installed()
    cnr_ref = R.fix.do_fix(T, A, Rf)                               # the reference's own code
    cns_ref = R.segmentation.do_segmentation(cnr_ref, "haar")
    I.install()
    try:
        assert R.fix.do_fix.__wrapped__ is not None and I.installed()
        cnr = R.fix.do_fix(T, A, Rf)                               # same call, now on the B200
        assert type(cnr) is type(cnr_ref)
        for c in ("chromosome", "start", "end", "gene"):
            assert cnr.data[c].tolist() == cnr_ref.data[c].tolist()
        np.testing.assert_array_equal(cnr.data["log2"].to_numpy(), cnr_ref.data["log2"].to_numpy())
        np.testing.assert_allclose(cnr.data["weight"].to_numpy(), cnr_ref.data["weight"].to_numpy(), rtol=1e-10)
        cns = R.segmentation.do_segmentation(cnr, "haar")
        assert type(cns) is type(cns_ref)
        for c in ("chromosome", "start", "end", "probes", "gene"):
            assert cns.data[c].tolist() == cns_ref.data[c].tolist()
        np.testing.assert_allclose(cns.data["log2"].to_numpy(), cns_ref.data["log2"].to_numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(cns.data["weight"].to_numpy(), cns_ref.data["weight"].to_numpy(), rtol=1e-9)
        # CBS through the reference's entry point: no Rscript needed any more; structural checks of test_r.py
        seg, text = R.segmentation.do_segmentation(cnr, "cbs", save_dataframe=True, processes=2)
        assert len(seg) >= len(set(cnr.data["chromosome"])) and text.startswith('"ID"\t"chrom"')
        assert seg.data["probes"].sum() <= len(cnr) and (seg.data["end"] > seg.data["start"]).all()
        serial = R.segmentation.do_segmentation(cnr, "cbs", processes=1)
        assert serial.data.equals(seg.data)                       # test_r.py:294-310 serial == parallel
        # an out-of-scope option still reaches the reference's own callable
        none_seg = R.segmentation.do_segmentation(cnr, "none")
        assert len(none_seg) > 0
    finally:
        I.uninstall()
    assert not I.installed()

"""The synthetic exome panel (C2's construction at 1/16 scale) pinned to the UNMODIFIED reference:
tests/golden/exome_mini.npz holds the reference's own do_fix and do_segmentation('haar') outputs for two
tumours (oracle/make_golden.py gen_exome_mini).  CPU: the composed oracle (oracle/pipeline.py) reproduces
them; GPU: do_fix, the cohort plan and do_segmentation reproduce them."""
import numpy as np
import pytest

from conftest import golden_table


def world():
    from cnvkit_b200 import synth

    bins = synth.exome_bins(n_targets=9000, n_antitargets=3000, seed=7)
    ref = synth.exome_reference(bins, seed=8)
    samples = [synth.exome_sample(bins, ref, seed=k, n_events=4) for k in range(2)]
    return bins, ref, samples


def check_inputs(g, bins, ref, samples):
    """the generators must still produce the inputs the golden file was made from"""
    np.testing.assert_allclose(g["ref_checksum"], [np.sum(ref["log2"]), np.sum(ref["spread"]), np.sum(bins["gc"]),
                                                   np.sum(bins["start"])], rtol=1e-13)
    for k, s in enumerate(samples):
        np.testing.assert_allclose(g[f"s{k}_in_log2_checksum"], [np.sum(s["log2"]), np.sum(s["depth"])], rtol=1e-13)


def test_oracle_pipeline_matches_reference(golden):
    from oracle import pipeline as OP

    g = golden("exome_mini")
    bins, ref, samples = world()
    check_inputs(g, bins, ref, samples)
    refc = dict(log2=ref["log2"], depth=ref["depth"], spread=ref["spread"], gc=bins["gc"], rmask=bins["rmask"])
    for k, s in enumerate(samples):
        want = golden_table(g, f"s{k}_cnr").data
        got = OP.do_fix(bins, s, refc)
        assert int(got["keep"].sum()) == len(want)
        np.testing.assert_array_equal(bins["start"][got["keep"]], want["start"].to_numpy())
        np.testing.assert_allclose(got["log2"], want["log2"].to_numpy(), rtol=0, atol=1e-12)
        np.testing.assert_allclose(got["weight"], want["weight"].to_numpy(), rtol=1e-10)
        kk = np.flatnonzero(got["keep"])
        chrom = np.asarray(bins["chromosome"], dtype=object)[kk]
        (arm, lo, hi, mean), fidx = OP.segment_columns(chrom, bins["start"][kk], bins["end"][kk],
                                                       want["log2"].to_numpy(), want["weight"].to_numpy(), "haar")
        cns = golden_table(g, f"s{k}_cns_haar").data
        assert len(arm) == len(cns)
        np.testing.assert_array_equal(np.sort(hi - lo), np.sort(cns["probes"].to_numpy()))


@pytest.mark.gpu
def test_gpu_fix_and_haar_match_reference(golden):
    from cnvkit_b200 import cohort, fix, segmentation, synth

    g = golden("exome_mini")
    bins, ref, samples = world()
    check_inputs(g, bins, ref, samples)
    tabs = [synth.exome_tables(bins, ref, s, sample_id=f"S{k}") for k, s in enumerate(samples)]
    panel = cohort.Panel(*tabs[0])
    for k, (T, A, R) in enumerate(tabs):
        want = golden_table(g, f"s{k}_cnr").data
        cnr = fix.do_fix(T, A, R)
        assert list(cnr.data.columns) == list(want.columns)
        for c in ("chromosome", "start", "end", "gene"):
            np.testing.assert_array_equal(cnr.data[c].to_numpy(), want[c].to_numpy())
        np.testing.assert_allclose(cnr.data["log2"].to_numpy(), want["log2"].to_numpy(), rtol=0, atol=1e-12)
        np.testing.assert_allclose(cnr.data["weight"].to_numpy(), want["weight"].to_numpy(), rtol=1e-10)
        d_log2, d_w, d_depth = panel.fix(T, A)
        np.testing.assert_array_equal(d_log2.cpu().numpy(), cnr.data["log2"].to_numpy())
        # segment the reference's own .cnr so that only the segmenter and transfer_fields are compared
        ref_cnr = golden_table(g, f"s{k}_cnr")
        cns = segmentation.do_segmentation(ref_cnr, "haar")
        wc = golden_table(g, f"s{k}_cns_haar").data
        assert len(cns) == len(wc)
        for c in ("chromosome", "start", "end", "probes", "gene"):
            np.testing.assert_array_equal(cns.data[c].to_numpy(), wc[c].to_numpy())
        np.testing.assert_allclose(cns.data["log2"].to_numpy(), wc["log2"].to_numpy(), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(cns.data["weight"].to_numpy(), wc["weight"].to_numpy(), rtol=1e-10)
        np.testing.assert_allclose(cns.data["depth"].to_numpy(), wc["depth"].to_numpy(), rtol=1e-10)

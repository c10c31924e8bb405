"""GPU: the cohort plan (cnvkit_b200.cohort.Panel) gives exactly the tables of the per-sample public
API (do_fix + do_segmentation), for CBS and Haar, including a sample that needs the general path."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    from cnvkit_b200 import synth

    bins = synth.exome_bins(n_targets=9000, n_antitargets=3000, seed=7)
    ref = synth.exome_reference(bins, seed=8)
    samples = [synth.exome_sample(bins, ref, seed=k, n_events=4) for k in range(4)]
    tabs = [synth.exome_tables(bins, ref, s, sample_id=f"S{k}") for k, s in enumerate(samples)]
    return bins, ref, tabs


def assert_same(a, b):
    pd.testing.assert_frame_equal(a.data.reset_index(drop=True), b.data.reset_index(drop=True), check_exact=True)


@pytest.mark.parametrize("method", ["cbs", "haar"])
def test_cohort_equals_per_sample_api(world, method):
    from cnvkit_b200 import cohort, fix, segmentation

    bins, ref, tabs = world
    T0, A0, R = tabs[0]
    panel = cohort.Panel(T0, A0, R)
    got = panel.run([(T, A) for T, A, _ in tabs], method=method, batch=3, keep_cnr=True,
                    sample_ids=[f"S{k}" for k in range(len(tabs))])
    assert len(got) == len(tabs)
    for (T, A, _), (cnr, cns) in zip(tabs, got):
        want_cnr = fix.do_fix(T, A, R)
        assert_same(cnr, want_cnr)
        want_cns = segmentation.do_segmentation(want_cnr, method)
        assert_same(cns, want_cns)
        assert cns.meta.get("sample_id") == T.sample_id
        assert len(cns) >= 2


def test_cohort_columns_input_and_general_path(world):
    from cnvkit_b200 import cohort, fix, segmentation

    bins, ref, tabs = world
    T0, A0, R = tabs[0]
    panel = cohort.Panel(T0, A0, R)
    T1, A1, _ = tabs[1]
    # plain column dicts instead of tables
    cols = ({"log2": T1.data["log2"].to_numpy(), "depth": T1.data["depth"].to_numpy()},
            {"log2": A1.data["log2"].to_numpy(), "depth": A1.data["depth"].to_numpy()})
    # a sample with NaN coverage in one target bin changes the kept set -> general path
    T2, A2, _ = tabs[2]
    t2 = T2.data.copy()
    t2.loc[17, "log2"] = np.nan
    T2n = T2.as_dataframe(t2)
    got = panel.run([cols, (T2n, A2)], method="cbs", keep_cnr=True, sample_ids=["S1", "S2"])
    want1 = fix.do_fix(T1, A1, R)
    assert_same(got[0][0], want1)
    assert_same(got[0][1], segmentation.do_segmentation(want1, "cbs"))
    want2 = fix.do_fix(T2n, A2, R)
    assert_same(got[1][0], want2)
    assert len(got[1][0]) == len(want1) - int(panel.t.keep[17])
    assert_same(got[1][1], segmentation.do_segmentation(want2, "cbs"))
    with pytest.raises(ValueError):
        panel.fix({"log2": np.zeros(3), "depth": np.zeros(3)}, cols[1])


@pytest.mark.parametrize("method", ["cbs", "haar"])
def test_cohort_vs_cpu_oracle_pipeline(world, method):
    """fix + segment of the CUDA path against the composed CPU oracle (oracle/pipeline.py, built from
    the units pinned to the reference): ratios bit-exact, weights 1e-12, identical segment
    boundaries, segment means 1e-9."""
    from cnvkit_b200 import cohort
    from oracle import pipeline as OP

    bins, ref, tabs = world
    T0, A0, R = tabs[0]
    panel = cohort.Panel(T0, A0, R)
    refc = dict(log2=ref["log2"], depth=ref["depth"], spread=ref["spread"], gc=bins["gc"], rmask=bins["rmask"])
    for T, A, _ in tabs[:2]:
        sample = {"log2": np.empty(len(bins["start"])), "depth": np.empty(len(bins["start"]))}
        t = bins["is_target"]
        for c in ("log2", "depth"):
            sample[c][t] = T.data[c].to_numpy()
            sample[c][~t] = A.data[c].to_numpy()
        want = OP.do_fix(bins, sample, refc)
        d_log2, d_w, d_depth = panel.fix(T, A)
        np.testing.assert_array_equal(d_log2.cpu().numpy(), want["log2"])
        np.testing.assert_allclose(d_w.cpu().numpy(), want["weight"], rtol=1e-12, atol=0)
        k = np.flatnonzero(want["keep"])
        chrom = np.asarray(bins["chromosome"], dtype=object)[k]
        start, end = bins["start"][k], bins["end"][k]
        # segment the oracle's own .cnr columns on both sides so that only the segmenter differs
        (o_arm, o_lo, o_hi, o_mean), fidx = OP.segment_columns(chrom, start, end, want["log2"], d_w.cpu().numpy(),
                                                               method, nthreads=4)
        cns = panel.segment([(d_log2, d_w, d_depth)], method)[0]
        # interior boundaries (the first/last segment of an arm is stretched by transfer_fields)
        got_probes = cns.data["probes"].to_numpy()
        np.testing.assert_array_equal(np.sort(got_probes), np.sort(o_hi - o_lo))
        got = cns.data.sort_values(["chromosome", "start"]).reset_index(drop=True)
        o_start = start[fidx[o_lo]]
        o_chrom = chrom[fidx[o_lo]]
        want_tab = pd.DataFrame({"chromosome": o_chrom, "start": o_start, "probes": o_hi - o_lo, "log2": o_mean})
        want_tab = want_tab.sort_values(["chromosome", "start"]).reset_index(drop=True)
        np.testing.assert_array_equal(got["probes"].to_numpy(), want_tab["probes"].to_numpy())
        np.testing.assert_allclose(got["log2"].to_numpy(), want_tab["log2"].to_numpy(), rtol=1e-9, atol=0)


def test_cohort_threaded_workers_give_the_same_tables(world):
    from cnvkit_b200 import cohort

    bins, ref, tabs = world
    T0, A0, R = tabs[0]
    panel = cohort.Panel(T0, A0, R)
    work = [(T, A) for T, A, _ in tabs] * 3
    one = panel.run(work, method="cbs", batch=2, keep_cnr=True)
    many = panel.run(work, method="cbs", batch=2, keep_cnr=True, workers=3)
    assert len(one) == len(many) == len(work)
    for (c1, s1), (c2, s2) in zip(one, many):
        assert_same(c1, c2)
        assert_same(s1, s2)


def test_run_files_writes_the_tables_of_the_per_sample_path(world, tmp_path):
    """File-to-file cohort driver: .cnn files in, .cnr/.cns files out, byte-identical to writing the
    tables of do_fix / do_segmentation sample by sample (one sample has a NaN row -> general path)."""
    import os

    from cnvkit_b200 import cohort, fix, segmentation, tabio

    bins, ref, tabs = world
    tdir, odir, wdir = tmp_path / "in", tmp_path / "out", tmp_path / "want"
    for d in (tdir, odir, wdir):
        os.makedirs(d)
    R = tabs[0][2]
    ref_path = str(tdir / "reference.cnn")
    tabio.write(R, ref_path)
    tf, af = [], []
    for k, (T, A, _) in enumerate(tabs[:3]):
        t = T.data.copy()
        if k == 2:
            t.loc[5, "log2"] = np.nan           # read_tab drops the row; this sample's bins then differ
        tp, ap = str(tdir / f"S{k}.targetcoverage.cnn"), str(tdir / f"S{k}.antitargetcoverage.cnn")
        tabio.write(T.as_dataframe(t), tp)
        tabio.write(A, ap)
        tf.append(tp)
        af.append(ap)
    outs = cohort.run_files(tf, af, ref_path, str(odir), method="cbs", batch=2, workers=2, io_threads=2)
    assert len(outs) == 3
    Rr = tabio.read(ref_path)
    for k, (cnr_path, cns_path) in enumerate(outs):
        T = tabio.read(tf[k], sample_id=f"S{k}")
        A = tabio.read(af[k], sample_id=f"S{k}")
        want_cnr = fix.do_fix(T, A, Rr)
        want_cns = segmentation.do_segmentation(want_cnr, "cbs")
        a, b = str(wdir / f"S{k}.cnr"), str(wdir / f"S{k}.cns")
        tabio.write(want_cnr, a)
        tabio.write(want_cns, b)
        assert open(cnr_path, "rb").read() == open(a, "rb").read()
        assert open(cns_path, "rb").read() == open(b, "rb").read()
        assert os.path.basename(cns_path) == f"S{k}.cns"


def test_cohort_confidence_intervals_equal_do_segmetrics(world, tmp_path):
    """Panel.run(ci=...) adds exactly the columns of do_segmetrics(cnr, cns, interval_stats=['ci'], ...) --
    the step `cnvkit.py batch` runs after segmentation (batch.py:559-566) -- and they agree with the CPU
    oracle on the same bins; run_files writes them into the .cns."""
    from cnvkit_b200 import cohort, segmetrics, tabio
    from oracle import segmetrics as OS

    bins, ref, tabs = world
    T0, A0, R = tabs[0]
    panel = cohort.Panel(T0, A0, R)
    pairs = [(T, A) for T, A, _ in tabs[:3]]
    plain = panel.run(pairs, method="cbs", keep_cnr=True)
    for ci, kw in ((True, dict(alpha=0.5, bootstraps=100, smoothed=True)),
                   (dict(alpha=0.05, smoothed=False), dict(alpha=0.05, bootstraps=100, smoothed=False))):
        got = panel.run(pairs, method="cbs", keep_cnr=True, ci=ci)
        for (cnr, cns), (_, cns_ci) in zip(plain, got):
            want = segmetrics.do_segmetrics(cnr, cns, interval_stats=["ci"], **kw)
            assert_same(cns_ci, want)
            d, s = cnr.data, cns.data
            lo, hi = OS.bin_ranges(d["chromosome"], d["start"], d["end"], s["chromosome"], s["start"], s["end"])
            assert (hi - lo >= s["probes"].to_numpy()).all()          # filtered bins count too
            olo, ohi = OS.ci_table(d["log2"].to_numpy(), d["weight"].to_numpy(), lo, hi, **kw)
            np.testing.assert_allclose(cns_ci["ci_lo"].to_numpy(), olo, rtol=0, atol=1e-9)
            np.testing.assert_allclose(cns_ci["ci_hi"].to_numpy(), ohi, rtol=0, atol=1e-9)
            assert (cns_ci["ci_lo"].to_numpy() <= cns_ci["ci_hi"].to_numpy()).all()
    # file-to-file
    tf, af = [], []
    for k, (T, A, _) in enumerate(tabs[:2]):
        tf.append(str(tmp_path / f"S{k}.targetcoverage.cnn"))
        af.append(str(tmp_path / f"S{k}.antitargetcoverage.cnn"))
        tabio.write(T, tf[-1])
        tabio.write(A, af[-1])
    outs = cohort.run_files(tf, af, R, str(tmp_path / "out"), ci=True, workers=2, batch=1)
    for _, cns_path in outs:
        tab = tabio.read(cns_path)
        assert {"ci_lo", "ci_hi"} <= set(tab.data.columns)
        assert np.isfinite(tab["ci_lo"].to_numpy()).all() and (tab["ci_lo"].to_numpy() <= tab["ci_hi"].to_numpy()).all()


def test_cohort_batch_path_flags_and_redoes_guarded_samples(world):
    """The batch entry point evaluates the "mostly empty" guard of load_adjust_coverages (fix.py:275) speculatively;
    a sample that trips it is redone by the step-by-step path, next to clean samples of the same batch."""
    from cnvkit_b200 import _lib, cohort, fix, segmentation

    bins, ref, tabs = world
    T0, A0, R = tabs[0]
    panel = cohort.Panel(T0, A0, R)
    assert panel._native_ok("cbs")
    T3, A3, _ = tabs[3]
    t3 = T3.data.copy()
    low = np.arange(len(t3)) % 5 != 0                      # 80 % of the targets far below the coverage floor
    t3.loc[low, "log2"] = -24.0
    t3.loc[low, "depth"] = 0.0
    T3low = T3.as_dataframe(t3)
    work = [(tabs[1][0], tabs[1][1]), (T3low, A3), (tabs[2][0], tabs[2][1])]
    res = panel._chunk_native(work, ["a", "b", "c"], None, True, None)
    assert res[0] is not None and res[2] is not None and res[1] is None
    got = panel.run(work, method="cbs", batch=3, keep_cnr=True, sample_ids=["a", "b", "c"])
    for (T, A), (cnr, cns) in zip(work, got):
        want_cnr = fix.do_fix(T, A, R)
        assert_same(cnr, want_cnr)
        assert_same(cns, segmentation.do_segmentation(want_cnr, "cbs"))
    assert _lib.PANEL_LOW_COVERAGE == 2

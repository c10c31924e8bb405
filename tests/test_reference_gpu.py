"""GPU parity for the reference build (cnvkit_b200.reference): bias_correct_logr per normal and
summarize_info (biweight location / midvariance per bin) against the table produced by the
unmodified reference's do_reference (tests/golden/regression.npz, test_regression.py:75-126)."""
import os

import numpy as np
import pytest

from conftest import golden_table
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
NORMALS = ("p2-5_5", "p2-9_5", "p2-20_5")


@pytest.fixture(scope="module")
def Rf():
    from cnvkit_b200 import reference

    return reference


def test_summarize_info_golden(Rf, golden):
    g = golden("regression")
    got = Rf.summarize_info(g["ref_all_logr"], g["ref_all_depths"])
    np.testing.assert_allclose(got["log2"], g["ref_summary_log2"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(got["depth"], g["ref_summary_depth"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(got["spread"], g["ref_summary_spread"], rtol=1e-11, atol=1e-14)


@pytest.mark.parametrize("rows", [1, 2, 3, 8, 33, 257])
def test_summarize_info_vs_oracle_shapes(Rf, rows):
    rng = np.random.default_rng(rows)
    ncols = 301
    logr = rng.normal(0, 0.4, (rows, ncols))
    logr[:, 7] = 0.25                       # MAD == 0 column
    if rows > 4:
        logr[rng.integers(0, rows, 40), rng.integers(0, ncols, 40)] = np.nan   # missing values are dropped
        logr[0, 11] = 9.0                    # gross outlier
    dep = np.exp2(rng.normal(5, 1, (rows, ncols)))
    got = Rf.summarize_info(logr, dep)
    for c in range(ncols):
        loc = O.biweight_location(logr[:, c])
        np.testing.assert_allclose(got["log2"][c], loc, rtol=1e-12, atol=1e-14, equal_nan=True)
        np.testing.assert_allclose(got["spread"][c], O.biweight_midvariance(logr[:, c], initial=loc), rtol=1e-11,
                                   atol=1e-14, equal_nan=True)
        np.testing.assert_allclose(got["depth"][c], O.biweight_location(dep[:, c]), rtol=1e-12, atol=1e-14)


def write_normals(golden, tmp_path):
    from cnvkit_b200 import cnary

    g = golden("regression")
    tfn, afn = [], []
    for s in NORMALS:
        for kind, dest in (("target", tfn), ("antitarget", afn)):
            path = os.path.join(tmp_path, f"{s}.{kind}coverage.cnn")
            cnary.write(golden_table(g, f"normal_{s}_{kind}"), path)
            dest.append(path)
    return g, tfn, afn


def test_do_reference_regression_fixture(Rf, golden, tmp_path):
    g, tfn, afn = write_normals(golden, str(tmp_path))
    sexes = {str(k): bool(v) for k, v in zip(g["ref_sex_ids"], g["ref_sex_is_female"])}
    got = Rf.do_reference(tfn, afn, is_haploid_x_reference=True, sexes=sexes)
    want = golden_table(g, "reference")
    assert list(got.data.columns) == list(want.data.columns)
    for col in ("chromosome", "start", "end", "gene"):
        np.testing.assert_array_equal(got.data[col].to_numpy(), want.data[col].to_numpy())
    np.testing.assert_array_equal(got.data["gc"].to_numpy(), want.data["gc"].to_numpy())
    np.testing.assert_allclose(got.data["log2"].to_numpy(), want.data["log2"].to_numpy(), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(got.data["depth"].to_numpy(), want.data["depth"].to_numpy(), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(got.data["spread"].to_numpy(), want.data["spread"].to_numpy(), rtol=1e-11, atol=1e-13)


def test_do_reference_errors(Rf, golden, tmp_path):
    g, tfn, afn = write_normals(golden, str(tmp_path))
    with pytest.raises(ValueError, match="Unequal number of target and antitarget files"):
        Rf.do_reference(tfn, afn[:2], female_samples=True)
    with pytest.raises(RuntimeError, match="bins do not match"):
        Rf.do_reference([tfn[0], afn[1]], None, female_samples=True)
    ref = Rf.do_reference(tfn, None, female_samples=True)   # targets only (amplicon)
    assert len(ref) == len(golden_table(g, "normal_p2-5_5_target"))


def test_do_reference_accepts_files_that_differ_in_formatting_only(Rf, golden, tmp_path):
    """The fast path compares a hash of the raw coordinate / gene bytes; a file whose rows are in another order (or
    lacks the depth column) parses to the same bins, and the reference -- which compares parsed keys,
    reference.py:594-598 -- accepts it.  So must we: same table as from the files as written."""
    import pandas as pd

    g, tfn, afn = write_normals(golden, str(tmp_path))
    sexes = {str(k): bool(v) for k, v in zip(g["ref_sex_ids"], g["ref_sex_is_female"])}
    want = Rf.do_reference(tfn, afn, is_haploid_x_reference=True, sexes=sexes)
    # second target file: rows reversed; third antitarget file (if any): rows reversed as well
    for path in (tfn[1], afn[-1]):
        df = pd.read_csv(path, sep="\t", dtype={"chromosome": str})
        df.iloc[::-1].to_csv(path, sep="\t", index=False, float_format="%.6g")
    got = Rf.do_reference(tfn, afn, is_haploid_x_reference=True, sexes=sexes)
    pd.testing.assert_frame_equal(got.data, want.data, check_exact=True)


def test_full_pipeline_reference_fix_segment(Rf, golden, tmp_path):
    """test_regression.py end to end on the GPU path: reference -> fix -> segment(haar)."""
    from cnvkit_b200 import fix, segmentation

    g, tfn, afn = write_normals(golden, str(tmp_path))
    sexes = {str(k): bool(v) for k, v in zip(g["ref_sex_ids"], g["ref_sex_is_female"])}
    ref = Rf.do_reference(tfn, afn, is_haploid_x_reference=True, sexes=sexes)
    cnr = fix.do_fix(golden_table(g, "sample_target"), golden_table(g, "sample_antitarget"), ref)
    want_cnr = golden_table(g, "cnr")
    # the frozen fixtures are compared at rtol 1e-5 / atol 1e-6 (test_regression.py:48-49); we are far inside
    np.testing.assert_allclose(cnr.data["log2"].to_numpy(), want_cnr.data["log2"].to_numpy(), rtol=0, atol=1e-10)
    np.testing.assert_allclose(cnr.data["weight"].to_numpy(), want_cnr.data["weight"].to_numpy(), rtol=1e-9)
    cns = segmentation.do_segmentation(cnr, "haar")
    want = golden_table(g, "cns_haar")
    assert len(cns) == len(want)
    np.testing.assert_array_equal(cns.data["start"].to_numpy(), want.data["start"].to_numpy())
    np.testing.assert_array_equal(cns.data["end"].to_numpy(), want.data["end"].to_numpy())
    np.testing.assert_array_equal(cns.data["probes"].to_numpy(), want.data["probes"].to_numpy())

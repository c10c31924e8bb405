"""Sample-sex inference of do_reference(female_samples=None): the host logic on CPU
(reference._reconcile_sex_guesses, reference.py:335-382) and, on the GPU, compare_sex_chromosomes
(cnary.py:484-632; medians by device radix select) against vectors of the unmodified reference
(tests/golden/sex.npz)."""
import numpy as np
import pytest

from conftest import golden_table


def _guess(row):
    return None if row[0] < 0 else (bool(row[0]), float(row[1]))


def test_reconcile_sex_guesses_matches_reference(golden):
    from cnvkit_b200 import reference

    g = golden("sex")
    for k, (t, a, want) in enumerate(zip(g["recon_t"], g["recon_a"], g["recon_out"])):
        t, a = _guess(t), _guess(a)
        got = reference._reconcile_sex_guesses({"s": t} if t else {}, {"s": a} if a else {})
        assert bool(got["s"]) == bool(want), (k, t, a)
    assert reference._chrx_maleness_confidence(float("nan")) == 0.0
    assert reference._chrx_maleness_confidence(-1.0) == 0.0
    assert reference._chrx_maleness_confidence(2.0) == pytest.approx(reference._chrx_maleness_confidence(0.5))


@pytest.mark.gpu
def test_compare_sex_chromosomes_matches_reference(golden):
    from cnvkit_b200 import reference

    g = golden("sex")
    for i in range(int(g["n_cases"])):
        tab = golden_table(g, f"case{i}")
        parx = str(g[f"case{i}_parx"]) or None
        is_xy, stats = reference.compare_sex_chromosomes(tab, bool(g[f"case{i}_haploid"]), parx)
        want = int(g[f"case{i}_is_xy"])
        assert (-1 if is_xy is None else int(is_xy)) == want, i
        got = np.array([stats[k] for k in ("chrx_ratio", "chry_ratio", "chrx_male_lr", "chry_male_lr")])
        np.testing.assert_allclose(got, g[f"case{i}_stats"], rtol=1e-12, atol=1e-13, equal_nan=True)
    # no chrX rows -> not determinable
    tab = golden_table(g, "case0")
    no_x = tab.as_dataframe(tab.data[tab.data["chromosome"] != "chrX"].reset_index(drop=True))
    assert reference.compare_sex_chromosomes(no_x) == (None, {})


@pytest.mark.gpu
def test_do_reference_infers_sexes_like_the_reference(golden, tmp_path):
    """female_samples=None (the default call): the inferred sexes are those the reference inferred for the
    regression normals, and the table equals its do_reference output."""
    import test_reference_gpu as T

    from cnvkit_b200 import reference

    g, tfn, afn = T.write_normals(golden, str(tmp_path))
    want_sex = {str(k): bool(v) for k, v in zip(g["ref_sex_ids"], g["ref_sex_is_female"])}
    tg = reference._guess_sexes(tfn, False, None)
    ag = reference._guess_sexes(afn, False, None)
    got_sex = reference._reconcile_sex_guesses(tg, ag)
    assert {k: bool(v) for k, v in got_sex.items()} == want_sex
    got = reference.do_reference(tfn, afn, is_haploid_x_reference=True)
    want = golden_table(g, "reference")
    assert list(got.data.columns) == list(want.data.columns)
    np.testing.assert_allclose(got.data["log2"].to_numpy(), want.data["log2"].to_numpy(), rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(got.data["spread"].to_numpy(), want.data["spread"].to_numpy(), rtol=1e-11, atol=1e-13)

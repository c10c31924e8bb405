"""GPU: Haar depth + BAF breakpoint union (cnvlib/segmentation/haar.py:188-225 one_chrom_baf; per-segment BAF column,
segmentation/__init__.py:359) against tables the UNMODIFIED reference produced for the same inputs
(tests/golden/haar_baf.npz, written by oracle/make_golden.py gen_haar_baf from test/formats/p2-20_1.cnr plus synthetic
het SNPs).  The VariantArray aggregation is the reference's own host code (the `variants` object a caller passes in);
its class comes from baseline/_ref (or /root/reference), so the test skips where neither exists."""
import numpy as np
import pandas as pd
import pytest

from conftest import golden_table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def VariantArray():
    from oracle import refshim

    if not refshim.available():
        pytest.skip("reference tree not present (VariantArray is the reference's own class)")
    refshim.load()
    from cnvlib.vary import VariantArray as V

    return V


def _variants(g):
    cols = [str(c) for c in g["var__columns"]]
    return pd.DataFrame({c: (g[f"var__{c}"].astype(object) if g[f"var__{c}"].dtype.kind == "U" else g[f"var__{c}"])
                         for c in cols})


@pytest.mark.parametrize("mode", ["counts", "freq"])
def test_depth_baf_union_equals_reference(golden, VariantArray, mode):
    from cnvkit_b200 import segmentation as S

    g = golden("haar_baf")
    cnr = golden_table(golden("haar"), "cnr")
    vdf = _variants(g)
    if mode == "freq":
        vdf = vdf.drop(columns=["alt_count", "depth"])
    got = S.do_segmentation(cnr, "haar", variants=VariantArray(vdf))
    want = golden_table(g, f"cns_{mode}")
    assert len(got) == len(want) > int(g["n_plain"])            # the BAF pass added breakpoints
    for c in ("chromosome", "start", "end", "probes", "gene"):
        assert got.data[c].tolist() == want.data[c].tolist(), c
    np.testing.assert_allclose(got.data["log2"].to_numpy(), want.data["log2"].to_numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got.data["baf"].to_numpy(), want.data["baf"].to_numpy(), rtol=1e-12, equal_nan=True)
    np.testing.assert_allclose(got.data["weight"].to_numpy(), want.data["weight"].to_numpy(), rtol=1e-9)
    assert list(got.data.columns) == list(want.data.columns)

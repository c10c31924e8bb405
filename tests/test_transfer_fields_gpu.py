"""GPU: known-answer tests of transfer_fields ported from the reference's own suite
(test/test_segmentation.py:308-690, TransferFieldsTests) -- same inputs, same assertions, through
cnvkit_b200.segmentation.transfer_fields (the weight / depth reductions run on the device)."""
import warnings

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from cnvkit_b200 import segmentation

    return segmentation


def CNA(d):
    from cnvkit_b200.cnary import CopyNumArray

    return CopyNumArray(pd.DataFrame(d))


def test_nan_gene(S):  # test_segmentation.py: test_transfer_fields_nan_gene (issue #900)
    n = 10
    cnarr = CNA({"chromosome": ["chr1"] * n, "start": np.arange(0, n * 1000, 1000),
                 "end": np.arange(1000, n * 1000 + 1000, 1000),
                 "gene": ["GeneA", float("nan"), "GeneB"] * 3 + [float("nan")], "log2": np.zeros(n),
                 "depth": np.ones(n) * 100.0, "weight": np.ones(n)})
    seg = CNA({"chromosome": ["chr1"], "start": [0], "end": [n * 1000], "gene": ["-"], "log2": [0.0], "probes": [n],
               "weight": [0.0]})
    res = S.transfer_fields(seg, cnarr)
    gene = res["gene"].iat[0]
    assert isinstance(gene, str) and "nan" not in gene.lower() and "GeneA" in gene and "GeneB" in gene
    assert res["weight"].iat[0] == 10.0 and res["depth"].iat[0] == 100.0


def test_genes_near_segment_end(S):  # #688, EGFR geometry
    bins = CNA({"chromosome": ["chr7"] * 6,
                "start": [54246732, 54800000, 55018770, 55019096, 55032092, 55088166],
                "end": [54300000, 54900000, 55019096, 55019423, 55032193, 55088469],
                "gene": ["VSTM2A", "SEC61G", "EGFR", "EGFR", "EGFR", "EGFR"], "log2": np.zeros(6),
                "depth": np.ones(6) * 100.0, "weight": np.ones(6)})
    seg = CNA({"chromosome": ["chr7", "chr7"], "start": [54246732, 55032092], "end": [55031592, 55365525],
               "gene": ["-", "-"], "log2": [0.0, 0.0], "probes": [4, 2], "weight": [0.0, 0.0]})
    res = S.transfer_fields(seg, bins)
    assert "EGFR" in res["gene"].iat[0] and "VSTM2A" in res["gene"].iat[0] and "EGFR" in res["gene"].iat[1]
    assert res["weight"].tolist() == [4.0, 2.0]


def test_nan_weights(S):
    n = 10
    w = [1.0, np.nan, 1.0, np.nan, 1.0, 1.0, np.nan, 1.0, 1.0, 1.0]
    cnarr = CNA({"chromosome": ["chr1"] * n, "start": np.arange(0, n * 1000, 1000),
                 "end": np.arange(1000, n * 1000 + 1000, 1000), "gene": ["GeneA"] * n, "log2": np.zeros(n),
                 "depth": np.ones(n) * 100.0, "weight": w})
    seg = CNA({"chromosome": ["chr1"], "start": [0], "end": [n * 1000], "gene": ["-"], "log2": [0.0], "probes": [n],
               "weight": [0.0]})
    res = S.transfer_fields(seg, cnarr)
    assert res["weight"].iat[0] == pytest.approx(7.0) and not np.isnan(res["depth"].iat[0])
    assert res["depth"].iat[0] == pytest.approx(100.0)
    allnan = cnarr.as_dataframe(cnarr.data.assign(weight=np.nan))
    seg2 = CNA({"chromosome": ["chr1"], "start": [0], "end": [n * 1000], "gene": ["-"], "log2": [0.0], "probes": [n],
                "weight": [0.0]})
    res2 = S.transfer_fields(seg2, allnan)
    assert res2["weight"].iat[0] == 0.0 and res2["depth"].iat[0] == 0.0


def test_stretches_every_chromosome(S):
    n = 10
    origins = {"chr1": 0, "chr2": 1_000_000, "chr3": 2_000_000}
    bins = pd.concat([pd.DataFrame({"chromosome": [c] * n, "start": o + np.arange(0, n * 1000, 1000),
                                    "end": o + np.arange(1000, n * 1000 + 1000, 1000), "gene": ["G"] * n,
                                    "log2": np.zeros(n), "depth": np.ones(n) * 100.0, "weight": np.ones(n)})
                      for c, o in origins.items()], ignore_index=True)
    seg = CNA({"chromosome": ["chr1", "chr1", "chr2", "chr2", "chr3", "chr3"],
               "start": [o + s for o in origins.values() for s in (2000, 5000)],
               "end": [o + e for o in origins.values() for e in (5000, 8000)], "gene": ["-"] * 6, "log2": [0.0] * 6,
               "probes": [3, 3] * 3, "weight": [0.0] * 6})
    res = S.transfer_fields(seg, CNA(bins))
    assert res["start"].tolist() == [o + s for o in origins.values() for s in (0, 5000)]
    assert res["end"].tolist() == [o + e for o in origins.values() for e in (5000, 10000)]
    assert res["weight"].tolist() == [5.0] * 6


def test_rejects_unknown_chromosome(S):
    cnarr = CNA({"chromosome": ["01"] * 4, "start": [0, 1000, 2000, 3000], "end": [1000, 2000, 3000, 4000],
                 "gene": ["G"] * 4, "log2": np.zeros(4), "depth": np.ones(4), "weight": np.ones(4)})
    seg = CNA({"chromosome": ["1"], "start": [0], "end": [4000], "gene": ["-"], "log2": [0.0], "probes": [4]})
    with pytest.raises(ValueError, match="absent from the input bins: 1"):
        S.transfer_fields(seg, cnarr)


def _bin_gap_pair(with_weight):
    d = {"chromosome": ["chr1", "chr1"], "start": [0, 100_000], "end": [1000, 101_000], "gene": ["AAA", "BBB"],
         "log2": [0.0, 0.0], "depth": [100.0, 200.0]}
    if with_weight:
        d["weight"] = [1.0, 2.0]
    seg = CNA({"chromosome": ["chr1"] * 3, "start": [0, 40_000, 100_000], "end": [1000, 60_000, 101_000],
               "gene": ["-"] * 3, "log2": [0.0] * 3, "probes": [1, 0, 1]})
    return CNA(d), seg


def test_segment_spanning_no_bins(S):
    cnarr, seg = _bin_gap_pair(True)
    res = S.transfer_fields(seg, cnarr)
    assert res["gene"].tolist() == ["AAA", "-", "BBB"]
    assert res["weight"].tolist() == [1.0, 0.0, 2.0] and res["depth"].tolist() == [100.0, 0.0, 200.0]


def test_no_weight_column_spanning_no_bins(S):
    cnarr, seg = _bin_gap_pair(False)
    with warnings.catch_warnings():
        warnings.simplefilter("error", RuntimeWarning)
        res = S.transfer_fields(seg, cnarr)
    assert res["depth"].tolist() == [100.0, 0.0, 200.0] and res["weight"].tolist() == [1.0, 0.0, 1.0]
    assert res.data["weight"].dtype == np.float64 and res.data["depth"].dtype == np.float64


def test_non_contiguous_chromosomes(S):
    cnarr = CNA({"chromosome": ["chr2", "chr2", "chr10", "chr10"], "start": [0, 1000, 0, 1000],
                 "end": [1000, 2000, 1000, 2000], "gene": ["AAA", "BBB", "CCC", "DDD"], "log2": [0.0] * 4,
                 "depth": [100.0, 200.0, 300.0, 400.0], "weight": [1.0, 2.0, 3.0, 4.0]})
    seg = CNA({"chromosome": ["chr2", "chr10", "chr2", "chr10"], "start": [0, 0, 1000, 1000],
               "end": [1000, 1000, 2000, 2000], "gene": ["-"] * 4, "log2": [0.0] * 4})
    res = S.transfer_fields(seg, cnarr)
    assert res["gene"].tolist() == ["AAA", "CCC", "BBB", "DDD"]
    assert res["weight"].tolist() == [1.0, 3.0, 2.0, 4.0] and res["depth"].tolist() == [100.0, 300.0, 200.0, 400.0]


def test_filtered_bins_gapped_index(S):
    cnarr = CNA({"chromosome": ["chr1"] * 4, "start": [0, 1000, 2000, 3000], "end": [1000, 2000, 3000, 4000],
                 "gene": ["AAA", "BBB", "CCC", "DDD"], "log2": [0.0] * 4, "depth": [100.0, 200.0, 300.0, 400.0],
                 "weight": [1.0, 2.0, 3.0, 4.0]})
    kept = cnarr[np.array([False, False, True, True])]
    assert kept.data.index.tolist() == [2, 3]
    seg = CNA({"chromosome": ["chr1", "chr1"], "start": [2000, 3000], "end": [3000, 4000], "gene": ["-"] * 2,
               "log2": [0.0] * 2, "probes": [1, 1]})
    res = S.transfer_fields(seg, kept)
    assert res["gene"].tolist() == ["CCC", "DDD"]
    assert res["depth"].tolist() == [300.0, 400.0] and res["weight"].tolist() == [3.0, 4.0]


def test_nested_bins_use_the_mask_of_irange_nested(S):
    """skgenome/intersect.py _irange_nested, mode 'outer': with a bin nested inside a longer one (`end` not
    ascending) a segment takes the rows before the first start >= its end whose end lies beyond its start --
    a mask, not a run of rows.  Checked against a direct numpy statement of that rule."""
    start = np.array([0, 100, 150, 400, 600, 900])
    end = np.array([1000, 200, 500, 700, 800, 1200])          # bin 0 spans bins 1-4: nested
    w = np.array([1.0, 2.0, 4.0, 8.0, 16.0, 32.0])
    d = np.array([10.0, 20.0, 30.0, 40.0, 50.0, 60.0])
    cnarr = CNA({"chromosome": ["chr1"] * 6, "start": start, "end": end, "gene": list("ABCDEF"), "log2": np.zeros(6),
                 "depth": d, "weight": w})
    seg = CNA({"chromosome": ["chr1"] * 2, "start": [0, 550], "end": [550, 1200], "gene": ["-"] * 2,
               "log2": [0.0, 0.1], "probes": [3, 3], "weight": [0.0, 0.0]})
    res = S.transfer_fields(seg, cnarr)
    for k, (s0, s1) in enumerate([(0, 550), (550, 1200)]):
        mask = (end > s0) & (np.arange(6) < np.searchsorted(start, s1))
        assert res["weight"].iat[k] == w[mask].sum()
        assert res["depth"].iat[k] == pytest.approx((d[mask] * w[mask]).sum() / w[mask].sum(), rel=1e-12)
        assert res["gene"].iat[k] == ",".join(np.array(list("ABCDEF"))[mask])

"""CPU: the native .cnn/.cnr/.cns reader / writer (csrc/tsv.cpp via cnvkit_b200.tabio) against pandas
-- which is what the reference uses (skgenome/tabio/tab.py:18-33, tabio/__init__.py:175-195) -- on
synthetic tables and, in the authoring container, on the reference's own test/formats files."""
import glob
import os

import numpy as np
import pandas as pd
import pytest

from cnvkit_b200 import cnary, tabio


def synthetic_table(n=5000, seed=3):
    rng = np.random.default_rng(seed)
    chrom = np.repeat(np.array(["chr1", "chr2", "chrX"], dtype=object), [n // 2, n // 4, n - n // 2 - n // 4])
    start = np.concatenate([np.sort(rng.integers(0, 10**8, m)) for m in (n // 2, n // 4, n - n // 2 - n // 4)])
    log2 = rng.normal(0, 1, n)
    log2[::97] = np.nan
    log2[5] = 1e-7
    log2[6] = -123456789.0
    log2[7] = 0.0
    log2[8] = -0.0
    log2[9] = 2.5e-5
    depth = np.abs(rng.normal(100, 30, n))
    depth[11] = np.nan
    gene = np.array([f"G{i // 7},H{i // 11}" if i % 5 else "Antitarget" for i in range(n)], dtype=object)
    gene[3] = "-"
    return pd.DataFrame({"chromosome": chrom, "start": start, "end": start + 200, "gene": gene, "log2": log2,
                         "depth": depth, "probes": rng.integers(1, 50, n), "weight": rng.uniform(0, 1, n)})


def frames_equal(got, want):
    assert list(got.columns) == list(want.columns)
    assert len(got) == len(want)
    for c in want.columns:
        w, g = want[c], got[c]
        if pd.api.types.is_float_dtype(w.dtype):
            assert pd.api.types.is_float_dtype(g.dtype), c
            np.testing.assert_array_equal(g.to_numpy(dtype=float), w.to_numpy(dtype=float))
        elif pd.api.types.is_integer_dtype(w.dtype):
            assert pd.api.types.is_integer_dtype(g.dtype), c
            np.testing.assert_array_equal(g.to_numpy(), w.to_numpy())
        else:
            wl = [None if (isinstance(v, float) and v != v) or v is None or v is pd.NA else str(v) for v in w.tolist()]
            gl = [None if (isinstance(v, float) and v != v) or v is None or v is pd.NA else str(v) for v in g.tolist()]
            assert gl == wl, c


def test_writer_is_byte_identical_to_pandas(tmp_path):
    df = synthetic_table()
    a, b = tmp_path / "pandas.tsv", tmp_path / "native.tsv"
    df.to_csv(a, sep="\t", index=False, float_format="%.6g")
    tabio.write(df, b)
    assert a.read_bytes() == b.read_bytes()
    empty = df.iloc[:0]
    empty.to_csv(a, sep="\t", index=False, float_format="%.6g")
    tabio.write(empty, b)
    assert a.read_bytes() == b.read_bytes()


def test_reader_matches_pandas_on_synthetic(tmp_path):
    df = synthetic_table()
    p = tmp_path / "t.cnr"
    df.to_csv(p, sep="\t", index=False, float_format="%.6g")
    want = pd.read_csv(p, sep="\t", dtype={"chromosome": "str"})
    want = want.dropna(subset=["log2"])
    got = tabio.read_frame(p)
    frames_equal(got.reset_index(drop=True), want.reset_index(drop=True))
    # full read(): same table as the pandas-based reader of cnary
    a = tabio.read(p)
    b = cnary.read(p)
    frames_equal(a.data.reset_index(drop=True), b.data.reset_index(drop=True))
    assert a.sample_id == b.sample_id == "t"


def test_reader_edge_cases(tmp_path):
    p = tmp_path / "e.cnn"
    p.write_bytes(b'chromosome\tstart\tend\tgene\tlog2\n1\t10\t20\t"a\tb"\t0.5\r\n\n2\t30\t40\tNA\t\n3\t50\t60\tx\t-inf\n')
    got = tabio.read_frame(p)           # row 2 has no log2 -> dropped; blank line skipped; CRLF ok
    assert got["chromosome"].tolist() == ["1", "3"]
    assert got["gene"].tolist() == ["a\tb", "x"]
    assert got["log2"].tolist() == [0.5, -np.inf]
    assert got["start"].dtype == np.int64
    bad = tmp_path / "bad.cnn"
    bad.write_bytes(b"chromosome\tstart\n1\t2\t3\n")
    from cnvkit_b200._lib import CnvkError

    with pytest.raises(CnvkError):
        tabio.read_frame(bad)
    blank = tmp_path / "blank.cnn"
    blank.write_bytes(b"")
    assert len(tabio.read(blank)) == 0


def test_read_columns_and_hash(tmp_path):
    df = synthetic_table(2000)
    df["log2"] = df["log2"].fillna(0.25)
    a, b, c = tmp_path / "A.cnn", tmp_path / "B.cnn", tmp_path / "C.cnn"
    df.to_csv(a, sep="\t", index=False, float_format="%.6g")
    df2 = df.assign(log2=df["log2"] + 1.0)
    df2.to_csv(b, sep="\t", index=False, float_format="%.6g")
    df3 = df.copy()
    df3.loc[100, "end"] += 1
    df3.to_csv(c, sep="\t", index=False, float_format="%.6g")
    ca, ha = tabio.read_columns(a)
    cb, hb = tabio.read_columns(b)
    cc, hc = tabio.read_columns(c)
    assert ha == hb and ha != hc
    want = pd.read_csv(a, sep="\t")
    np.testing.assert_array_equal(ca["log2"], want["log2"].to_numpy())
    np.testing.assert_array_equal(ca["depth"], want["depth"].to_numpy())
    np.testing.assert_array_equal(cb["log2"], pd.read_csv(b, sep="\t")["log2"].to_numpy())
    with pytest.raises(KeyError):
        tabio.read_columns(a, columns=("nope",))


@pytest.mark.needs_reference
def test_reader_and_writer_on_reference_fixtures(tmp_path):
    files = sorted(glob.glob("/root/reference/test/formats/*.cn[nrs]"))
    assert len(files) > 20
    for f in files:
        try:
            want = pd.read_csv(f, sep="\t", dtype={"chromosome": "str"})
        except pd.errors.EmptyDataError:
            continue
        if "log2" in want.columns:
            want = want.dropna(subset=["log2"])
        got = tabio.read_frame(f)
        frames_equal(got.reset_index(drop=True), want.reset_index(drop=True))
        a, b = tmp_path / "p.tsv", tmp_path / "n.tsv"
        want.to_csv(a, sep="\t", index=False, float_format="%.6g")
        tabio.write(got, b)
        assert a.read_bytes() == b.read_bytes(), f


def test_writer_dictionary_strings_quoting_and_special_floats(tmp_path):
    """String columns go through dictionary codes (cnvk_tsv_dict): missing values, fields that need
    csv quoting, many distinct values, and %.6g of inf / tiny / huge / negative-zero floats."""
    n = 70_000                                   # more than one formatting chunk
    rng = np.random.default_rng(3)
    genes = np.array([f"G{k}" for k in rng.integers(0, 9000, n)], dtype=object)
    genes[5] = 'has"quote'
    genes[6] = "has\ttab"
    genes[7] = None
    genes[8] = "multi\nline"
    vals = rng.normal(0, 1, n)
    vals[:6] = [np.inf, -np.inf, np.nan, 1e-300, 1e21, -0.0]
    df = pd.DataFrame({"chromosome": [f"chr{1 + k % 22}" for k in range(n)], "start": np.arange(n) * 100,
                       "end": np.arange(n) * 100 + 50, "gene": genes, "log2": vals,
                       "depth": rng.uniform(0, 1e6, n)})
    ours, theirs = tmp_path / "ours.tsv", tmp_path / "pandas.tsv"
    tabio.write(df, str(ours))
    df.to_csv(theirs, sep="\t", index=False, float_format="%.6g")
    assert ours.read_bytes() == theirs.read_bytes()

"""GPU parity for GC / RepeatMasker content from a FASTA genome (csrc/fasta.cu through the C ABI):
bit-exact against the unmodified reference's calculate_gc_lo / get_fasta_stats vectors
(tests/golden/fasta.npz), and do_reference(fa_fname=...) against the reference's own table."""
import os

import numpy as np
import pytest

from conftest import golden_table
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    from cnvkit_b200 import fasta

    return fasta


def write_genome(g, tmp_path, crlf=False):
    text = g["fasta_bytes"].tobytes()
    if crlf:
        text = text.replace(b"\n", b"\r\n")
    path = os.path.join(str(tmp_path), "genome.fa")
    with open(path, "wb") as fh:
        fh.write(text)
    return path


def test_calculate_gc_lo_known_answers(F, golden):
    g = golden("fasta")
    got = np.array([F.calculate_gc_lo(str(q)) for q in g["kat_seqs"]])
    np.testing.assert_array_equal(got, g["kat_gc_lo"])


def test_ragged_ranges_equal_oracle(F):
    rng = np.random.default_rng(5)
    data = rng.choice(np.frombuffer(b"ACGTacgtNn\n\rRYxX*-", dtype=np.uint8), size=200_003)
    lo = np.concatenate([rng.integers(0, len(data), 400), np.arange(0, 40), [0, 0, len(data), len(data) - 1]])
    size = np.concatenate([rng.integers(0, 70, 200), rng.integers(0, 5000, 200), np.arange(0, 40), [len(data), 0, 10, 1]])
    hi = np.minimum(lo + size, len(data) + 3)           # a few ranges run past the buffer: clamped
    for shift in (0, 1, 7):                             # unaligned buffer starts
        gc, rm = F.gc_lo_of_ranges(data[shift:], lo, hi)
        want = np.array([O.calculate_gc_lo(data[shift:][a:b].tobytes()) for a, b in zip(lo, hi)])
        np.testing.assert_array_equal(gc, want[:, 0])
        np.testing.assert_array_equal(rm, want[:, 1])


@pytest.mark.parametrize("crlf", [False, True])
def test_get_fasta_stats_matches_reference(F, golden, tmp_path, crlf):
    g = golden("fasta")
    fa = write_genome(g, tmp_path, crlf)
    for tag in ("tgt", "anti"):
        gc, rm = F.get_fasta_stats(golden_table(g, tag), fa)
        np.testing.assert_array_equal(gc, g[f"{tag}_gc"])
        np.testing.assert_array_equal(rm, g[f"{tag}_rmask"])
    import pandas as pd

    from cnvkit_b200.cnary import CopyNumArray

    missing = CopyNumArray(pd.DataFrame({"chromosome": ["chr9"], "start": [0], "end": [10], "gene": ["-"], "log2": [0.0]}), {})
    with pytest.raises(KeyError, match="chr9 not in"):
        F.get_fasta_stats(missing, fa)


def test_do_reference_with_fasta_matches_reference(golden, tmp_path):
    from cnvkit_b200 import reference, tabio

    g = golden("fasta")
    fa = write_genome(g, tmp_path)
    tfn, afn = [], []
    for k in range(3):
        for tag, kind, dest in (("tgt", "targetcoverage", tfn), ("anti", "antitargetcoverage", afn)):
            tab = golden_table(g, tag)
            tab = tab.as_dataframe(tab.data.assign(log2=g[f"N{k}_{kind}_log2"], depth=g[f"N{k}_{kind}_depth"]))
            path = os.path.join(str(tmp_path), f"N{k}.{kind}.cnn")
            tabio.write(tab, path)
            dest.append(path)
    got = reference.do_reference(tfn, afn, fa, female_samples=True)
    want = golden_table(g, "ref")
    assert list(got.data.columns) == list(want.data.columns)
    assert "gc" in got.data.columns and "rmask" in got.data.columns
    for col in ("chromosome", "start", "end", "gene"):
        np.testing.assert_array_equal(got.data[col].to_numpy(), want.data[col].to_numpy())
    np.testing.assert_array_equal(got.data["gc"].to_numpy(), want.data["gc"].to_numpy())
    np.testing.assert_array_equal(got.data["rmask"].to_numpy(), want.data["rmask"].to_numpy())   # NaN on target rows
    np.testing.assert_allclose(got.data["log2"].to_numpy(), want.data["log2"].to_numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got.data["depth"].to_numpy(), want.data["depth"].to_numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(got.data["spread"].to_numpy(), want.data["spread"].to_numpy(), rtol=1e-11, atol=1e-12)


def write_beds(g, tmp_path):
    t_bed, a_bed = os.path.join(str(tmp_path), "targets.bed"), os.path.join(str(tmp_path), "antitargets.bed")
    with open(t_bed, "wb") as fh:
        fh.write(g["targets_bed"].tobytes())
    with open(a_bed, "wb") as fh:
        fh.write(g["antitargets_bed"].tobytes())
    return t_bed, a_bed


def assert_tables_equal(got, want):
    assert list(got.data.columns) == list(want.data.columns)
    assert len(got) == len(want)
    for col in want.data.columns:
        np.testing.assert_array_equal(got.data[col].to_numpy(), want.data[col].to_numpy(), err_msg=col)


def test_do_reference_flat_matches_reference(golden, tmp_path):
    from cnvkit_b200 import reference

    g = golden("fasta")
    fa = write_genome(g, tmp_path)
    t_bed, a_bed = write_beds(g, tmp_path)
    assert_tables_equal(reference.do_reference_flat(t_bed, a_bed, fa, is_haploid_x_reference=True), golden_table(g, "flat_hx"))
    assert_tables_equal(reference.do_reference_flat(t_bed, a_bed, fa), golden_table(g, "flat"))

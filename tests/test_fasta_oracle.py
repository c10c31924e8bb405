"""CPU-only: FASTA GC / RepeatMasker content -- the oracle (oracle/np_oracle.py) against vectors of the
unmodified reference (tests/golden/fasta.npz), and the host logic of cnvkit_b200.fasta (index scan,
.fai parsing, base -> file-offset ranges) checked by counting the byte ranges with numpy."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O


def count_ranges(data, lo, hi):
    return np.array([O.calculate_gc_lo(data[a:b].tobytes()) for a, b in zip(lo, hi)])


def test_calculate_gc_lo_known_answers(golden):
    g = golden("fasta")
    got = np.array([O.calculate_gc_lo(str(q)) for q in g["kat_seqs"]])
    np.testing.assert_array_equal(got, g["kat_gc_lo"])


def test_fasta_stats_oracle_matches_reference(golden):
    g = golden("fasta")
    text = g["fasta_bytes"].tobytes().decode()
    for tag in ("tgt", "anti"):
        gc, rm = O.fasta_stats(text, g[f"{tag}__chromosome"], g[f"{tag}__start"], g[f"{tag}__end"])
        np.testing.assert_array_equal(gc, g[f"{tag}_gc"])
        np.testing.assert_array_equal(rm, g[f"{tag}_rmask"])


@pytest.mark.parametrize("variant", ["scan", "fai", "crlf", "no_final_newline"])
def test_host_index_and_byte_ranges(golden, tmp_path, variant):
    from cnvkit_b200 import fasta

    g = golden("fasta")
    text = g["fasta_bytes"].tobytes().decode()
    if variant == "crlf":
        text = text.replace("\n", "\r\n")
    if variant == "no_final_newline":
        text = text.rstrip("\n")
    path = str(tmp_path / "genome.fa")
    with open(path, "wb") as fh:
        fh.write(text.encode())
    seqs = O.parse_fasta(text)
    index = fasta.build_index(path)
    assert list(index) == list(seqs)
    for name, seq in seqs.items():
        assert index[name].length == len(seq), (name, index[name])
        assert index[name].linebases == 60
        assert index[name].linewidth == (62 if variant == "crlf" else 61)
    if variant == "fai":
        with open(path + ".fai", "w") as fh:
            for name, e in index.items():
                fh.write(f"{name}\t{e.length}\t{e.offset}\t{e.linebases}\t{e.linewidth}\n")
        assert fasta.read_index(path) == index
    else:
        assert fasta.read_index(path) == index
    data = np.frombuffer(text.encode(), dtype=np.uint8)
    for tag in ("tgt", "anti"):
        chrom, start, end = g[f"{tag}__chromosome"], g[f"{tag}__start"], g[f"{tag}__end"]
        gc = np.zeros(len(chrom))
        rm = np.zeros(len(chrom))
        for name in dict.fromkeys(chrom.tolist()):
            rows = np.flatnonzero(chrom == name)
            lo, hi = fasta.byte_ranges(index[name], start[rows], end[rows])
            assert (hi >= lo).all() and lo.min() >= index[name].offset
            vals = count_ranges(data, lo, hi)
            gc[rows], rm[rows] = vals[:, 0], vals[:, 1]
        np.testing.assert_array_equal(gc, g[f"{tag}_gc"])
        np.testing.assert_array_equal(rm, g[f"{tag}_rmask"])
    # slices outside the sequence are empty, as in Python
    e = index["chr2"]
    lo, hi = fasta.byte_ranges(e, [e.length + 5, 0, e.length - 1], [e.length + 50, 0, e.length + 7])
    assert hi[0] == lo[0] and hi[1] == lo[1] and hi[2] == lo[2] + 1


def test_compressed_or_bad_input_is_refused(tmp_path):
    from cnvkit_b200 import fasta

    gz = tmp_path / "g.fa.gz"
    gz.write_bytes(b"\x1f\x8b\x08\x00rest")
    with pytest.raises(NotImplementedError):
        fasta.build_index(str(gz))
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b"not a fasta\nACGT\n")
    with pytest.raises(ValueError):
        fasta.build_index(str(bad))


def test_flat_reference_without_fasta_matches_reference(golden, tmp_path):
    """do_reference_flat(targets.bed) (reference.py:26-75) needs no device work without a FASTA: BED
    parsing (track line, name column), flat log2 / depth / spread columns, column order."""
    from conftest import golden_table

    from cnvkit_b200 import reference

    g = golden("fasta")
    t_bed = str(tmp_path / "targets.bed")
    with open(t_bed, "wb") as fh:
        fh.write(g["targets_bed"].tobytes())
    got = reference.do_reference_flat(t_bed)
    want = golden_table(g, "flat_nofa")
    assert list(got.data.columns) == list(want.data.columns)
    for col in want.data.columns:
        np.testing.assert_array_equal(got.data[col].to_numpy(), want.data[col].to_numpy(), err_msg=col)
    assert got.sample_id == "targets"
    other = tmp_path / "regions.txt"
    other.write_text("chr1:100-200\n")
    with pytest.raises(NotImplementedError):
        reference.bed2probes(str(other))


def test_index_scan_edge_cases(tmp_path):
    from cnvkit_b200 import fasta

    text = ">a first record\nACGTAC\nGTACGT\nAC\n\n>b\nGGGG\n>c desc\n" + "T" * 7 + "\n" + "T" * 3 + "\n\n\n"
    path = tmp_path / "g.fa"
    path.write_text(text)
    idx = fasta.build_index(str(path))
    assert {k: v.length for k, v in idx.items()} == {"a": 14, "b": 4, "c": 10}
    assert (idx["a"].linebases, idx["a"].linewidth) == (6, 7) and (idx["c"].linebases, idx["c"].linewidth) == (7, 8)
    data = np.frombuffer(text.encode(), dtype=np.uint8)
    seqs = O.parse_fasta(text)
    for name, seq in seqs.items():
        for s, e in ((0, len(seq)), (1, 5), (5, 9), (len(seq) - 1, len(seq) + 4), (3, 3)):
            lo, hi = fasta.byte_ranges(idx[name], [s], [e])
            got = bytes(data[lo[0]:hi[0]]).replace(b"\n", b"").decode()
            assert got == seq[s:e], (name, s, e, got)
    # a stale .fai (older than the FASTA) is ignored, a fresh one is used
    fai = tmp_path / "g.fa.fai"
    fai.write_text("a\t999\t16\t6\t7\n")
    os.utime(fai, (1, 1))
    assert fasta.read_index(str(path))["a"].length == 14
    os.utime(fai, None)
    os.utime(path, (2, 2))
    assert fasta.read_index(str(path))["a"].length == 999

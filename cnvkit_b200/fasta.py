"""GC and RepeatMasker content of bins from a FASTA genome on the GPU -- host mirror of
cnvlib/reference.py:859-895 (get_fasta_stats, calculate_gc_lo, fasta_extract_regions).

The reference slices each bin out of the genome with pyfaidx and counts letters in Python.  Here a
chromosome's block of the FASTA *file* (line breaks included) is copied to the device once and every
bin becomes a byte range under the ``.fai`` line geometry; `cnvk_fasta_gc_lo_dev` counts a/c/g/t in
either case over the ranges, so nothing is sliced or stripped on the host.  Slicing semantics are
Python's (pyfaidx's): coordinates are clamped to the chromosome, an empty slice gives (0.0, 0.0); a
chromosome missing from the FASTA raises ``KeyError`` as pyfaidx does.
"""
from __future__ import annotations

import logging
import mmap
import os
from typing import NamedTuple

import numpy as np

from . import _lib


class FaiEntry(NamedTuple):
    length: int      # bases
    offset: int      # file offset of the first base
    linebases: int   # bases per line
    linewidth: int   # bytes per line, line terminator included


def read_index(fa_fname: str) -> dict:
    """{sequence name: FaiEntry} from ``<fasta>.fai`` (samtools / pyfaidx format) when it exists and is
    not older than the FASTA, else by scanning the file."""
    fai = fa_fname + ".fai"
    if os.path.exists(fai) and os.path.getmtime(fai) >= os.path.getmtime(fa_fname):
        index = {}
        with open(fai) as handle:
            for line in handle:
                f = line.rstrip("\n").split("\t")
                if len(f) >= 5:
                    index[f[0]] = FaiEntry(int(f[1]), int(f[2]), int(f[3]), int(f[4]))
        return index
    return build_index(fa_fname)


def build_index(fa_fname: str) -> dict:
    """Scan a plain-text FASTA for its record geometry (what ``samtools faidx`` writes), without
    creating a file next to the user's genome."""
    index = {}
    size = os.path.getsize(fa_fname)
    if size == 0:
        return index
    with open(fa_fname, "rb") as fh, mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ) as mm:
        if mm[:2] == b"\x1f\x8b":
            raise NotImplementedError("compressed FASTA is not supported; decompress it (or provide bgzip + .fai via cnvlib)")
        pos = 0 if mm[:1] == b">" else mm.find(b"\n>") + 1
        if pos == 0 and mm[:1] != b">":
            raise ValueError(f"{fa_fname} does not look like a FASTA file")
        while 0 <= pos < size:
            eol = mm.find(b"\n", pos)
            if eol < 0:
                break
            name = mm[pos + 1:eol].split()[0].decode() if eol > pos + 1 else ""
            seq0 = eol + 1
            nxt = mm.find(b"\n>", seq0 - 1)
            seq1 = size if nxt < 0 else nxt + 1           # one past the record's last byte
            first_eol = mm.find(b"\n", seq0, seq1)
            if first_eol < 0:                              # single unterminated line
                line = mm[seq0:seq1]
                linebases = len(line.rstrip(b"\r"))
                linewidth = linebases + 1
            else:
                linewidth = first_eol + 1 - seq0
                linebases = linewidth - 1 - (1 if mm[first_eol - 1:first_eol] == b"\r" and first_eol > seq0 else 0)
            while seq1 > seq0 and mm[seq1 - 1:seq1] in (b"\n", b"\r"):   # trailing terminators / blank lines
                seq1 -= 1
            span = seq1 - seq0
            if linebases <= 0:
                length = 0
                linebases, linewidth = 1, 2
            else:
                full, rem = divmod(span, linewidth)
                length = full * linebases + min(rem, linebases)
            index[name] = FaiEntry(length, seq0, linebases, linewidth)
            pos = -1 if nxt < 0 else nxt + 1
    return index


def byte_ranges(entry: FaiEntry, start, end):
    """File byte range [lo, hi) holding bases [start, end) of one sequence (Python slice clamping)."""
    s = np.clip(np.asarray(start, dtype=np.int64), 0, entry.length)
    e = np.clip(np.asarray(end, dtype=np.int64), 0, entry.length)
    e = np.maximum(e, s)

    def off(p):
        return entry.offset + (p // entry.linebases) * entry.linewidth + (p % entry.linebases)

    lo = off(s)
    hi = np.where(e > s, off(np.maximum(e, 1) - 1) + 1, lo)
    return lo.astype(np.int64), hi.astype(np.int64)


def gc_lo_of_ranges(data: np.ndarray, lo, hi):
    """(gc, rmask) of byte ranges of a host uint8 buffer (host-buffer C ABI)."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    lo, hi = _lib.i64(lo), _lib.i64(hi)
    gc = np.zeros(len(lo))
    rm = np.zeros(len(lo))
    if len(lo):
        _lib.call("cnvk_fasta_gc_lo", _lib.ptr(data), len(data), _lib.ptr(lo), _lib.ptr(hi), len(lo), _lib.ptr(gc),
                  _lib.ptr(rm))
    return gc, rm


def calculate_gc_lo(subseq) -> tuple:
    """cnvlib/reference.py:868-880 for one sequence string."""
    raw = subseq.encode("latin-1") if isinstance(subseq, str) else bytes(subseq)
    gc, rm = gc_lo_of_ranges(np.frombuffer(raw, dtype=np.uint8), [0], [len(raw)])
    return float(gc[0]), float(rm[0])


def get_fasta_stats(cnarr, fa_fname: str):
    """Calculate GC and RepeatMasker content of each bin in the FASTA genome (reference.py:859-866).
    Returns (gc, rmask) float arrays in the table's row order."""
    import torch

    from . import _dev

    logging.info("Calculating GC and RepeatMasker content in %s ...", fa_fname)
    index = read_index(fa_fname)
    chrom = cnarr.data["chromosome"].to_numpy()
    start = cnarr.data["start"].to_numpy(dtype=np.int64)
    end = cnarr.data["end"].to_numpy(dtype=np.int64)
    n = len(chrom)
    gc = np.zeros(n)
    rm = np.zeros(n)
    if n == 0:
        return gc, rm
    names = list(dict.fromkeys(chrom.tolist()))
    with open(fa_fname, "rb") as fh:
        for name in names:
            logging.info("Extracting sequences from chromosome %s", name)
            if name not in index:
                raise KeyError(f"{name} not in {fa_fname}.")
            rows = np.flatnonzero(chrom == name)
            lo, hi = byte_ranges(index[name], start[rows], end[rows])
            base, top = int(lo.min()), int(hi.max())
            if top <= base:
                continue
            fh.seek(base)
            block = np.frombuffer(fh.read(top - base), dtype=np.uint8)
            d_block = torch.from_numpy(block.copy()).to(_dev.device())
            d_lo, d_hi = _dev.up(lo - base), _dev.up(hi - base)
            d_gc, d_rm = _dev.empty(len(rows), torch.float64), _dev.empty(len(rows), torch.float64)
            _lib.call("cnvk_fasta_gc_lo_dev", d_block.data_ptr(), len(block), d_lo.data_ptr(), d_hi.data_ptr(), len(rows),
                      d_gc.data_ptr(), d_rm.data_ptr(), _dev.stream())
            gc[rows] = d_gc.cpu().numpy()
            rm[rows] = d_rm.cpu().numpy()
    return gc, rm

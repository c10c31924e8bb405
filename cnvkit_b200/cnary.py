"""Boundary container: a minimal CopyNumArray compatible with the reference's
(cnvlib/cnary.py:39-56 on top of skgenome/gary.py:37-94).

The hot-path entry points accept *either* this class or the reference's own
``cnvlib.cnary.CopyNumArray`` -- both expose ``.data`` (a pandas DataFrame with at least
chromosome/start/end/gene/log2), ``.meta`` and ``as_dataframe()`` -- and return the same class they
were given.  Only what the hot path needs is implemented; range algebra, gene grouping and sex
inference stay with the reference.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import genome

REQUIRED = ("chromosome", "start", "end", "gene", "log2")
_DTYPES = (str, int, int, str, float)


class CopyNumArray:
    _required_columns = REQUIRED
    _required_dtypes = _DTYPES

    def __init__(self, data_table=None, meta_dict=None):
        if data_table is None or (isinstance(data_table, pd.DataFrame) and not len(data_table.columns)):
            data_table = pd.DataFrame({c: pd.Series(dtype=("object" if t is str else t))
                                       for c, t in zip(REQUIRED, _DTYPES)})
        elif not isinstance(data_table, pd.DataFrame):
            data_table = pd.DataFrame(data_table)
        missing = [c for c in REQUIRED if c not in data_table.columns]
        if missing:
            raise ValueError(f"data table must have at least columns {REQUIRED!r}; got {tuple(data_table.columns)!r}")
        fix = {}
        if len(data_table):
            for c, t in zip(REQUIRED, _DTYPES):
                v = data_table[c].iat[0]
                if t is int:
                    ok = isinstance(v, (int, np.integer))
                elif t is float:
                    ok = isinstance(v, (float, np.floating))
                else:
                    ok = isinstance(v, str)
                if not ok:
                    fix[c] = t
        if fix:
            data_table = data_table.astype(fix)
        self.data = data_table
        self.meta = dict(meta_dict) if meta_dict else {}

    # -- basics
    def __len__(self):
        return len(self.data)

    def __bool__(self):
        return bool(len(self.data))

    def __contains__(self, key):
        return key in self.data.columns

    def __iter__(self):
        return self.data.itertuples(index=False)

    @property
    def sample_id(self):
        return self.meta.get("sample_id")

    @property
    def chromosome(self):
        return self.data["chromosome"]

    @property
    def start(self):
        return self.data["start"]

    @property
    def end(self):
        return self.data["end"]

    @property
    def log2(self):
        return self.data["log2"]

    def __getitem__(self, index):
        if isinstance(index, str):
            return self.data[index]
        if isinstance(index, (int, np.integer)):
            return self.data.iloc[index]
        if isinstance(index, slice):
            return self.as_dataframe(self.data.iloc[index])
        idx = np.asarray(index)
        if idx.dtype == bool:
            return self.as_dataframe(self.data[idx])
        return self.as_dataframe(self.data.iloc[idx])

    def __setitem__(self, key, value):
        if isinstance(key, str):
            self.data[key] = value
        else:
            raise NotImplementedError("only column assignment is supported")

    def __eq__(self, other):
        return isinstance(other, self.__class__) and self.data.equals(other.data)

    def as_dataframe(self, dframe, reset_index=False):
        if reset_index:
            dframe = dframe.reset_index(drop=True)
        return self.__class__(dframe, self.meta.copy())

    def copy(self):
        return self.as_dataframe(self.data.copy())

    def add_columns(self, **columns):
        return self.as_dataframe(self.data.assign(**columns))

    def keep_columns(self, colnames):
        cols = [c for c in self.data.columns if c in colnames]
        return self.__class__(self.data.loc[:, cols], self.meta.copy())

    def coords(self):
        return list(zip(self.data["chromosome"], self.data["start"], self.data["end"]))

    # -- ordering
    def sort(self):
        order = genome.genomic_order(self.data["chromosome"].to_numpy(), self.data["start"].to_numpy(),
                                     self.data["end"].to_numpy())
        n = len(order)
        if n and np.array_equal(order, np.arange(n)) and isinstance(self.data.index, pd.RangeIndex) \
                and self.data.index.start == 0 and self.data.index.step == 1:
            return                                   # already in genomic order with a clean index
        self.data = self.data.iloc[order].reset_index(drop=True)

    def sort_columns(self):
        extra = sorted(c for c in self.data.columns if c not in REQUIRED)
        want = list(REQUIRED) + extra
        if list(self.data.columns) != want:
            self.data = self.data.reindex(columns=want)

    def concat(self, others):
        table = pd.concat([o.data for o in others], ignore_index=True)
        result = self.as_dataframe(table)
        result.sort()
        return result

    def add(self, other):
        if len(other.data):
            self.data = pd.concat([self.data, other.data], ignore_index=True)
            self.sort()

    # -- traversal
    def by_chromosome(self):
        for chrom, sub in self.data.groupby("chromosome", sort=False):
            yield chrom, self.as_dataframe(sub)

    def by_arm(self, min_gap_size=1e5, min_arm_bins=50):
        for chrom, lo, hi in arm_ranges(self.data, min_gap_size, min_arm_bins):
            yield chrom, self.as_dataframe(self.data.iloc[lo:hi])


def arm_ranges(df: pd.DataFrame, min_gap_size=1e5, min_arm_bins=50):
    """(chromosome, lo, hi) row ranges of GenomicArray.by_arm (skgenome/gary.py:309-355): per
    chromosome, the largest inter-bin gap inside the middle 80 % (margin >= min_arm_bins) is the
    centromere if it is at least min_gap_size wide.  Rows of a chromosome must be contiguous."""
    names, offsets, _ = genome.chrom_runs(df["chromosome"])
    return arm_ranges_from_runs(names, offsets, df["start"].to_numpy(), df["end"].to_numpy(), min_gap_size,
                                min_arm_bins)


def arm_cut(start, end, lo, hi, min_gap_size=1e5, min_arm_bins=50):
    """Row offset (relative to lo) of the centromere split of the chromosome run [lo, hi), 0 = no split
    (skgenome/gary.py:309-355)."""
    n = hi - lo
    margin = max(min_arm_bins, round(0.1 * n))
    if n > 2 * margin + 1:
        gaps = start[lo + margin + 1 : hi - margin] - end[lo + margin : hi - margin - 1]
        k = int(np.argmax(gaps))
        if gaps[k] >= min_gap_size:
            return k + margin + 1
    return 0


def arm_ranges_from_runs(names, offsets, start, end, min_gap_size=1e5, min_arm_bins=50):
    out = []
    for c, name in enumerate(names):
        lo, hi = int(offsets[c]), int(offsets[c + 1])
        cut = arm_cut(start, end, lo, hi, min_gap_size, min_arm_bins)
        if cut:
            out.append((name, lo, lo + cut))
            out.append((name, lo + cut, hi))
        else:
            out.append((name, lo, hi))
    return out


def read(path, sample_id=None):
    """Read a .cnn/.cnr/.cns table: tab format with a header row, rows without log2 dropped
    (skgenome/tabio/tab.py:18-33), columns and rows sorted (tabio/__init__.py:116-119), sample id
    from the file name (cnvlib/core.py fbase)."""
    import logging
    import os

    try:
        df = pd.read_csv(path, sep="\t", dtype={"chromosome": "str"})
    except pd.errors.EmptyDataError:
        df = None
    if df is not None and "log2" in df.columns:
        d2 = df.dropna(subset=["log2"])
        if len(d2) < len(df):
            logging.warning("Dropped %d rows with missing log2 values", len(df) - len(d2))
            df = d2.copy()
    if sample_id is None:
        base = os.path.basename(str(path))
        sample_id = base.split(".")[0]
    arr = CopyNumArray(df, {"sample_id": sample_id, "filename": str(path)})
    arr.sort_columns()
    arr.sort()
    return arr


def write(cnarr, path):
    """Write a table the way tabio.write does: tab-separated, floats as %.6g
    (skgenome/tabio/__init__.py:175-195)."""
    cnarr.data.to_csv(path, sep="\t", index=False, float_format="%.6g")

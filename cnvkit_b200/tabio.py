"""Tab-separated .cnn / .cnr / .cns tables through the native reader / writer (csrc/tsv.cpp).

Host mirror of skgenome/tabio/__init__.py:39-119 (read, fmt="tab") + tab.py:18-33 (read_tab) and of
tabio/__init__.py:175-195 (write): same columns, dtypes, row dropping, sorting and "%.6g" text, without
pandas.read_csv / DataFrame.to_csv on the path.  `read_columns` is the cohort form: numeric columns
only, plus a hash of the coordinate columns so that a shared bin set is verified without
materialising a single Python string.
"""
from __future__ import annotations

import ctypes
import logging
import os

import numpy as np
import pandas as pd

from . import _lib

COORD_COLUMNS = ("chromosome", "start", "end", "gene")


class _Handle:
    def __init__(self, path):
        self.lib = _lib.load()
        self.h = ctypes.c_void_p()
        _lib.check(self.lib.cnvk_tsv_open(os.fsencode(str(path)), ctypes.byref(self.h)), "cnvk_tsv_open")
        nrows, ncols = ctypes.c_int64(), ctypes.c_int32()
        self.lib.cnvk_tsv_shape(self.h, ctypes.byref(nrows), ctypes.byref(ncols))
        self.nrows, self.ncols = nrows.value, ncols.value
        self.names = [self.lib.cnvk_tsv_colname(self.h, c).decode() for c in range(self.ncols)]

    def close(self):
        if self.h:
            self.lib.cnvk_tsv_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def kind(self, col):
        return self.lib.cnvk_tsv_col_kind(self.h, col)

    def f64(self, col, out=None):
        out = np.empty(self.nrows, dtype=np.float64) if out is None else out
        _lib.check(self.lib.cnvk_tsv_read_f64(self.h, col, out.ctypes.data), "cnvk_tsv_read_f64")
        return out

    def i64(self, col):
        out = np.empty(self.nrows, dtype=np.int64)
        _lib.check(self.lib.cnvk_tsv_read_i64(self.h, col, out.ctypes.data), "cnvk_tsv_read_i64")
        return out

    def strings(self, col, na_values=()):
        """Column as an object array of str; values listed in `na_values` become NaN.  The native side
        dictionary-codes the column, so only the distinct values are turned into Python strings."""
        n_unique, nbytes = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.cnvk_tsv_read_str_dict(self.h, col, None, None, None, 0, ctypes.byref(n_unique),
                                                   ctypes.byref(nbytes)), "cnvk_tsv_read_str_dict")
        if self.nrows == 0:
            return np.zeros(0, dtype=object)
        codes = np.empty(self.nrows, dtype=np.int32)
        off = np.empty(n_unique.value + 1, dtype=np.int64)
        buf = ctypes.create_string_buffer(max(nbytes.value, 1))
        _lib.check(self.lib.cnvk_tsv_read_str_dict(self.h, col, codes.ctypes.data, off.ctypes.data, buf, nbytes.value,
                                                   ctypes.byref(n_unique), ctypes.byref(nbytes)),
                   "cnvk_tsv_read_str_dict")
        raw = buf.raw[: nbytes.value]
        o = off.tolist()
        uniq = np.empty(len(o) - 1, dtype=object)
        for k, (a, b) in enumerate(zip(o[:-1], o[1:])):
            text = raw[a:b].decode("utf-8")
            uniq[k] = np.nan if text in na_values else text
        return uniq[codes]

    def hash_cols(self, cols):
        arr = (ctypes.c_int32 * len(cols))(*cols)
        return int(self.lib.cnvk_tsv_hash_cols(self.h, arr, len(cols)))


def read_frame(path) -> pd.DataFrame:
    """read_tab (tab.py:18-33): the table as a DataFrame with pandas' inferred dtypes, chromosome as
    str, rows without log2 dropped (with the reference's warning)."""
    with _Handle(path) as t:
        if t.ncols == 0:
            raise pd.errors.EmptyDataError("No columns to parse from file")
        cols = {}
        for c, name in enumerate(t.names):
            kind = 0 if name == "chromosome" else t.kind(c)
            if kind == 1:
                cols[name] = t.i64(c)
            elif kind == 2:
                cols[name] = t.f64(c)
            else:
                cols[name] = t.strings(c, () if name == "chromosome" else _NA_SET)
        df = pd.DataFrame(cols, columns=t.names)
    if "log2" in df.columns:
        keep = df["log2"].notna().to_numpy()
        if not keep.all():
            logging.warning("Dropped %d rows with missing log2 values", int((~keep).sum()))
            df = df[keep].copy()
    return df


_NA_STRINGS = ("", "#N/A", "#N/A N/A", "#NA", "-1.#IND", "-1.#QNAN", "-NaN", "-nan", "1.#IND", "1.#QNAN", "<NA>",
               "N/A", "NA", "NULL", "NaN", "None", "n/a", "nan", "null")
_NA_SET = frozenset(_NA_STRINGS)


def read(path, sample_id=None, into=None):
    """tabio.read(path, "tab", into=CopyNumArray) (tabio/__init__.py:39-119): sample id from the file
    name, columns and rows sorted."""
    from .cnary import CopyNumArray

    try:
        df = read_frame(path)
    except pd.errors.EmptyDataError:
        logging.info("Blank %s file?: %s", "tab", path)
        df = None
    if sample_id is None:
        sample_id = os.path.basename(str(path)).split(".")[0]
    arr = (into or CopyNumArray)(df, {"sample_id": sample_id, "filename": str(path)})
    arr.sort_columns()
    arr.sort()
    return arr


def read_columns(path, columns=("log2", "depth"), hash_columns=COORD_COLUMNS, pinned=False):
    """Cohort reader: the named float columns of a coverage table as numpy arrays (optionally views of
    pinned torch tensors, ready for an async H2D copy) and the FNV hash of `hash_columns`.  No strings,
    no DataFrame.  Returns (dict name -> float64[nrows], hash)."""
    with _Handle(path) as t:
        idx = {n: c for c, n in enumerate(t.names)}
        out = {}
        for name in columns:
            if name not in idx:
                raise KeyError(f"{path}: no column {name!r}")
            buf = None
            if pinned:
                import torch

                ten = torch.empty(t.nrows, dtype=torch.float64).pin_memory()
                buf = ten.numpy()
                out["_pinned_" + name] = ten
            out[name] = t.f64(idx[name], buf)
        h = t.hash_cols([idx[n] for n in hash_columns if n in idx]) if hash_columns else None
    return out, h


def write(cnarr, path):
    """tabio.write(garr, path) (tabio/__init__.py:175-195): tab-separated, header row, floats as %.6g,
    missing values as empty fields."""
    df = cnarr.data if hasattr(cnarr, "data") else cnarr
    n = len(df)
    names, kinds, ptrs, ptrs2, keep = [], [], [], [], []
    for name in df.columns:
        col = df[name]
        names.append(str(name).encode())
        if pd.api.types.is_bool_dtype(col.dtype):
            raise NotImplementedError("boolean columns are not part of the .cnn/.cnr/.cns schema")
        if pd.api.types.is_integer_dtype(col.dtype):
            a = np.ascontiguousarray(col.to_numpy(dtype=np.int64))
            kinds.append(1); ptrs.append(a.ctypes.data); ptrs2.append(0); keep.append(a)
        elif pd.api.types.is_float_dtype(col.dtype):
            a = np.ascontiguousarray(col.to_numpy(dtype=np.float64))
            kinds.append(2); ptrs.append(a.ctypes.data); ptrs2.append(0); keep.append(a)
        else:
            # dictionary-code the strings: a table has few distinct chromosomes / gene labels
            codes, uniques = pd.factorize(col, use_na_sentinel=True)
            enc = [str(v).encode("utf-8") for v in np.asarray(uniques, dtype=object).tolist()]
            off = np.zeros(len(enc) + 1, dtype=np.int64)
            if enc:
                np.cumsum([len(b) for b in enc], out=off[1:])
            blob = b"".join(enc)
            buf = ctypes.create_string_buffer(blob, max(len(blob), 1))
            codes = np.ascontiguousarray(codes, dtype=np.int32)
            d = _lib.TsvDict(ctypes.cast(off.ctypes.data, ctypes.POINTER(ctypes.c_int64)),
                             ctypes.cast(buf, ctypes.c_char_p), len(enc))
            kinds.append(3); ptrs.append(codes.ctypes.data); ptrs2.append(ctypes.addressof(d)); keep.append((codes, off, buf, d))
    k = len(names)
    c_names = (ctypes.c_char_p * k)(*names)
    c_kinds = (ctypes.c_int32 * k)(*kinds)
    c_ptrs = (ctypes.c_void_p * k)(*ptrs)
    c_ptrs2 = (ctypes.c_void_p * k)(*[p or None for p in ptrs2])
    _lib.call("cnvk_tsv_write", os.fsencode(str(path)), k, c_names, c_kinds, c_ptrs, c_ptrs2, n)

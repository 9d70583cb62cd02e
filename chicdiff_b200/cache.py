"""Columnar cache of the input tables (Arrow IPC files, read back memory-mapped).

setChicdiffExperiment() re-reads every CHiCAGO .Rds (gunzip + XDR decode of the whole object) and every .chinput text
file on every run (chicdiff.R:517-534, 614-623, 828, 1272); once the compute takes tens of milliseconds that is the wall
time.  A table that has been decoded once is written as an uncompressed Arrow IPC file of exactly the columns the
pipeline uses; the next run maps it and gets NumPy views of the file's pages (no decode, no copy: numeric columns are
stored without nulls -- NA_integer_ stays the INT_MIN sentinel, NA_real_ stays a NaN -- so the buffers are the arrays).
Character columns (the tblb / tlb bin labels) are dictionary-encoded.

    cols = load_or_build("rep1.Rds")             # decodes rep1.Rds and writes rep1.Rds.arrow the first time, maps it later
"""
import os

import numpy as np
import pyarrow as pa
import pyarrow.ipc as ipc

from . import rds

# the columns of a CHiCAGO interaction table that getFullRegionData1() reads (chicdiff.R:614-702, 820-853)
CHICAGO_COLUMNS = ("baitID", "otherEndID", "s_j", "s_i", "tblb", "tlb", "Tmean", "N", "distSign", "score")


def save_columns(path, columns, metadata=None):
    """dict name -> 1-D NumPy array (numeric, bool or object array of str / None) -> one Arrow IPC file, atomically"""
    arrays, names = [], []
    for name, col in columns.items():
        col = np.asarray(col)
        if col.dtype == object or col.dtype.kind in "US":
            arr = pa.array([None if v is None else str(v) for v in col.tolist()], type=pa.string()).dictionary_encode()
        else:
            arr = pa.array(np.ascontiguousarray(col))           # zero-copy wrap, no null bitmap
        arrays.append(arr)
        names.append(name)
    meta = {str(k): str(v) for k, v in (metadata or {}).items()}
    table = pa.Table.from_arrays(arrays, names=names).replace_schema_metadata(meta)
    tmp = path + ".tmp%d" % os.getpid()
    with pa.OSFile(tmp, "wb") as sink:
        with ipc.new_file(sink, table.schema) as writer:
            writer.write_table(table)
    os.replace(tmp, path)


def load_columns(path, columns=None):
    """-> (dict name -> NumPy array, metadata dict).  Numeric columns are read-only views of the memory-mapped file."""
    source = pa.memory_map(path, "r")
    table = ipc.open_file(source).read_all()
    meta = {k.decode(): v.decode() for k, v in (table.schema.metadata or {}).items()}
    out = {}
    for name in (columns or table.column_names):
        col = table.column(name)
        if pa.types.is_dictionary(col.type):
            col = col.combine_chunks()
            codes = col.indices.to_numpy(zero_copy_only=False)
            levels = np.asarray(col.dictionary.to_pylist(), dtype=object)
            vals = np.empty(len(col), dtype=object)
            valid = ~np.asarray(col.is_null().to_numpy(zero_copy_only=False), dtype=bool)
            vals[valid] = levels[np.asarray(codes[valid], dtype=np.int64)]
            vals[~valid] = None
            out[name] = vals
        else:
            chunks = col.chunks
            out[name] = chunks[0].to_numpy(zero_copy_only=True) if len(chunks) == 1 else \
                np.concatenate([c.to_numpy(zero_copy_only=True) for c in chunks])
    return out, meta


def _stamp(path):
    st = os.stat(path)
    return "%d:%d" % (st.st_size, st.st_mtime_ns)


def load_or_build(path, columns=CHICAGO_COLUMNS, cache_path=None, reader=None):
    """The table in `path` (.Rds by default; any reader(path) -> dict of columns) through its Arrow cache: the cache is
    used when it exists and carries the source file's size and modification time, rebuilt otherwise.  Only the requested
    columns that the table has are kept."""
    cache_path = cache_path or path + ".arrow"
    stamp = _stamp(path)
    if os.path.exists(cache_path):
        try:
            cols, meta = load_columns(cache_path)
            if meta.get("source_stamp") == stamp:
                return cols
        except (pa.ArrowInvalid, OSError):
            pass                                                  # unreadable cache: rebuild it
    table = reader(path) if reader is not None else rds.chicago_table(path)["columns"]
    keep = {k: v for k, v in table.items() if columns is None or k in columns}
    save_columns(cache_path, keep, metadata={"source": os.path.basename(path), "source_stamp": stamp})
    return load_columns(cache_path)[0]

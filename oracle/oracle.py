"""CPU oracle for the covering-index hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; the product package ``hyperspace_b200`` never does.

Two independent restatements of Spark's bucket function live here so they can check each other:

* ``libhs_oracle.so`` (``hs_oracle.c``, bound through ctypes) -- the C restatement, also the timed CPU baseline;
* the vectorised numpy functions ``np_hash_int`` / ``np_hash_long`` / ``np_pmod`` below.

``create_index`` restates the whole reference write path (CoveringIndex.createIndexData + CoveringIndex.write,
``src/main/scala/com/microsoft/hyperspace/index/covering/CoveringIndex.scala:56-71,140-192`` and
``index/DataFrameWriterExtensions.scala:50-68``) on the CPU, with pyarrow doing Parquet decode/encode so that the
oracle shares no codec code with the CUDA path.

Parity pin: BucketUnionTest.scala:101-123 golden vector + Spark's ``hash(1L) == -1712319331`` (tests/test_oracle.py).
Unpinned in the reference: page encodings, compression, row-group sizing, order among equal keys.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import uuid
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhs_oracle.so")

HSO_INT32, HSO_INT64, HSO_FLOAT, HSO_DOUBLE, HSO_BOOL, HSO_STRING = range(6)

_NP_TYPE = {
    np.dtype("int32"): HSO_INT32,
    np.dtype("int64"): HSO_INT64,
    np.dtype("float32"): HSO_FLOAT,
    np.dtype("float64"): HSO_DOUBLE,
    np.dtype("bool"): HSO_BOOL,
    np.dtype("uint8"): HSO_BOOL,
}


class _Column(ctypes.Structure):
    _fields_ = [
        ("type", ctypes.c_int32),
        ("data", ctypes.c_void_p),
        ("aux", ctypes.c_void_p),
        ("valid", ctypes.c_void_p),
    ]


def build(force: bool = False) -> str:
    """Compile hs_oracle.c -> libhs_oracle.so (gcc); returns the library path."""
    src = os.path.join(_HERE, "hs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libhs_oracle.so"])
    return _LIB_PATH


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.hso_hash_int.restype = ctypes.c_int32
        L.hso_hash_int.argtypes = [ctypes.c_int32, ctypes.c_int32]
        L.hso_hash_long.restype = ctypes.c_int32
        L.hso_hash_long.argtypes = [ctypes.c_int64, ctypes.c_int32]
        L.hso_hash_bytes.restype = ctypes.c_int32
        L.hso_hash_bytes.argtypes = [ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32]
        L.hso_bucket_ids.restype = None
        L.hso_bucket_ids.argtypes = [ctypes.POINTER(_Column), ctypes.c_int32, ctypes.c_int64, ctypes.c_int32,
                                     ctypes.c_void_p, ctypes.c_int32]
        L.hso_sort_perm.restype = None
        L.hso_sort_perm.argtypes = [ctypes.POINTER(_Column), ctypes.c_int32, ctypes.c_int64, ctypes.c_int32,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
        L.hso_range_select_i64.restype = None
        L.hso_range_select_i64.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                           ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
        L.hso_merge_join_i64.restype = ctypes.c_int64
        L.hso_merge_join_i64.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.hso_splitmix64.restype = ctypes.c_uint64
        L.hso_splitmix64.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        L.hso_synth_bucket_rows.restype = ctypes.c_int64
        L.hso_synth_bucket_rows.argtypes = [ctypes.c_uint64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p,
                                            ctypes.c_int64, ctypes.c_int32]
        _lib = L
    return _lib


# ----------------------------------------------------------------------------------------------
# numpy mirror of Spark's Murmur3_x86_32 (independent of the C code)
# ----------------------------------------------------------------------------------------------

def _u32(x):
    return np.asarray(x).astype(np.uint32)


def _rotl(x, r):
    return (x << np.uint32(r)) | (x >> np.uint32(32 - r))


def _mix_k1(k1):
    k1 = k1 * np.uint32(0xCC9E2D51)
    k1 = _rotl(k1, 15)
    return k1 * np.uint32(0x1B873593)


def _mix_h1(h1, k1):
    h1 = h1 ^ k1
    h1 = _rotl(h1, 13)
    return h1 * np.uint32(5) + np.uint32(0xE6546B64)


def _fmix(h1, length):
    h1 = h1 ^ np.uint32(length)
    h1 = h1 ^ (h1 >> np.uint32(16))
    h1 = h1 * np.uint32(0x85EBCA6B)
    h1 = h1 ^ (h1 >> np.uint32(13))
    h1 = h1 * np.uint32(0xC2B2AE35)
    return h1 ^ (h1 >> np.uint32(16))


def np_hash_int(v, seed=42) -> np.ndarray:
    """Murmur3_x86_32.hashInt over an int32 array; ``seed`` may be a scalar or a per-row array."""
    with np.errstate(over="ignore"):
        v = np.asarray(v).astype(np.int32).view(np.uint32)
        s = np.broadcast_to(np.asarray(seed).astype(np.int32).view(np.uint32), v.shape)
        return _fmix(_mix_h1(s, _mix_k1(v)), 4).view(np.int32)


def np_hash_long(v, seed=42) -> np.ndarray:
    """Murmur3_x86_32.hashLong over an int64 array (low word first, then high word)."""
    with np.errstate(over="ignore"):
        u = np.asarray(v).astype(np.int64).view(np.uint64)
        lo = (u & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        hi = (u >> np.uint64(32)).astype(np.uint32)
        s = np.broadcast_to(np.asarray(seed).astype(np.int32).view(np.uint32), lo.shape)
        h1 = _mix_h1(s, _mix_k1(lo))
        h1 = _mix_h1(h1, _mix_k1(hi))
        return _fmix(h1, 8).view(np.int32)


def py_hash_bytes(b: bytes, seed: int = 42) -> int:
    """Murmur3_x86_32.hashUnsafeBytes (pure Python; small cases only)."""
    with np.errstate(over="ignore"):
        h1 = np.uint32(seed & 0xFFFFFFFF)
        n = len(b)
        aligned = n - n % 4
        for i in range(0, aligned, 4):
            w = np.uint32(int.from_bytes(b[i:i + 4], "little"))
            h1 = _mix_h1(h1, _mix_k1(w))
        for i in range(aligned, n):
            v = b[i] - 256 if b[i] >= 128 else b[i]
            h1 = _mix_h1(h1, _mix_k1(np.uint32(v & 0xFFFFFFFF)))
        return int(np.asarray(_fmix(h1, n)).astype(np.uint32).view(np.int32))


def np_pmod(h, n: int) -> np.ndarray:
    """Spark Pmod on the signed 32-bit hash."""
    h = np.asarray(h).astype(np.int64)
    return (((h % n) + n) % n).astype(np.int32)


def np_hash_column(values: np.ndarray, seed, valid: Optional[np.ndarray] = None) -> np.ndarray:
    """One step of Spark's Murmur3Hash fold for one key column (null leaves the running hash unchanged)."""
    dt = values.dtype
    if dt == np.int32:
        h = np_hash_int(values, seed)
    elif dt == np.int64:
        h = np_hash_long(values, seed)
    elif dt == np.float32:
        v = np.where(values == 0.0, np.float32(0.0), values)
        bits = v.view(np.int32).copy()
        bits[np.isnan(v)] = 0x7FC00000
        h = np_hash_int(bits, seed)
    elif dt == np.float64:
        v = np.where(values == 0.0, 0.0, values)
        bits = v.view(np.int64).copy()
        bits[np.isnan(v)] = 0x7FF8000000000000
        h = np_hash_long(bits, seed)
    elif dt == np.bool_:
        h = np_hash_int(values.astype(np.int32), seed)
    else:
        raise TypeError(f"unsupported key dtype {dt}")
    if valid is not None:
        s = np.broadcast_to(np.asarray(seed).astype(np.int32), h.shape)
        h = np.where(valid.astype(bool), h, s)
    return h


def np_bucket_ids(key_columns: Sequence[np.ndarray], num_buckets: int,
                  valids: Optional[Sequence[Optional[np.ndarray]]] = None) -> np.ndarray:
    h = np.int32(42)
    for i, col in enumerate(key_columns):
        h = np_hash_column(col, h, None if valids is None else valids[i])
    return np_pmod(h, num_buckets)


# ----------------------------------------------------------------------------------------------
# C oracle wrappers
# ----------------------------------------------------------------------------------------------

def _as_columns(cols: Sequence[np.ndarray], valids: Optional[Sequence[Optional[np.ndarray]]]):
    keep = []
    arr = (_Column * len(cols))()
    for i, c in enumerate(cols):
        c = np.asarray(c)
        if c.dtype.kind in "OUS":  # strings / binary: int64 offsets[n + 1] + the bytes (utf-8 for str)
            raw = [(x if isinstance(x, (bytes, bytearray)) else ("" if x is None else str(x)).encode("utf-8")) for x in c.tolist()]
            offs = np.zeros(len(raw) + 1, dtype=np.int64)
            np.cumsum([len(b) for b in raw], out=offs[1:])
            blob = np.frombuffer(b"".join(raw) + b"\0", dtype=np.uint8).copy()
            keep += [offs, blob]
            arr[i].type = HSO_STRING
            arr[i].data = offs.ctypes.data
            arr[i].aux = blob.ctypes.data
        else:
            c = np.ascontiguousarray(c)
            keep.append(c)
            arr[i].type = _NP_TYPE[c.dtype]
            arr[i].data = c.ctypes.data
            arr[i].aux = None
        v = None if valids is None else valids[i]
        if v is not None:
            v = np.ascontiguousarray(np.asarray(v).astype(np.uint8))
            keep.append(v)
            arr[i].valid = v.ctypes.data
        else:
            arr[i].valid = None
    return arr, keep


def bucket_ids(key_columns: Sequence[np.ndarray], num_buckets: int, valids=None, nthreads: int = 1) -> np.ndarray:
    n = len(key_columns[0])
    arr, keep = _as_columns(key_columns, valids)
    out = np.empty(n, dtype=np.int32)
    lib().hso_bucket_ids(arr, len(key_columns), n, num_buckets, out.ctypes.data, nthreads)
    return out


def sort_perm(key_columns: Sequence[np.ndarray], num_buckets: int, buckets: np.ndarray, valids=None,
              nthreads: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """Returns (perm, bucket_offsets): rows ordered by (bucket, keys asc nulls-first, row index)."""
    n = len(key_columns[0])
    arr, keep = _as_columns(key_columns, valids)
    perm = np.empty(n, dtype=np.int64)
    offs = np.empty(num_buckets + 1, dtype=np.int64)
    b = np.ascontiguousarray(buckets.astype(np.int32))
    lib().hso_sort_perm(arr, len(key_columns), n, num_buckets, b.ctypes.data, perm.ctypes.data, offs.ctypes.data,
                        nthreads)
    return perm, offs


def range_select(keys: np.ndarray, lo: int, hi: int) -> Tuple[int, int]:
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    a, b = ctypes.c_int64(), ctypes.c_int64()
    lib().hso_range_select_i64(keys.ctypes.data, len(keys), lo, hi, ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def merge_join(lk: np.ndarray, rk: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    lk = np.ascontiguousarray(lk, dtype=np.int64)
    rk = np.ascontiguousarray(rk, dtype=np.int64)
    total = lib().hso_merge_join_i64(lk.ctypes.data, len(lk), rk.ctypes.data, len(rk), None, None, 0)
    li = np.empty(total, dtype=np.int64)
    ri = np.empty(total, dtype=np.int64)
    lib().hso_merge_join_i64(lk.ctypes.data, len(lk), rk.ctypes.data, len(rk), li.ctypes.data, ri.ctypes.data, total)
    return li, ri


def splitmix64(seed: int, idx: np.ndarray) -> np.ndarray:
    """Vectorised splitmix64(seed, i) -- the generator of SURVEY.md section 8d's synthetic tables."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (np.asarray(idx).astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synthetic_table(first_row: int, nrows: int, ncols: int = 5) -> Dict[str, np.ndarray]:
    """Synthetic table T of SURVEY.md section 8d: k:int64, v1:int64, v2:float64, v3:int32, v4:float32."""
    i = np.arange(first_row, first_row + nrows, dtype=np.uint64)
    cols = {"k": splitmix64(42, i).view(np.int64)}
    if ncols >= 2:
        cols["v1"] = (splitmix64(43, i) % np.uint64(1000)).astype(np.int64)
    if ncols >= 3:
        cols["v2"] = i.astype(np.float64) * 1e-3
    if ncols >= 4:
        cols["v3"] = (i % np.uint64(100)).astype(np.int32)
    if ncols >= 5:
        cols["v4"] = (i % np.uint64(4096)).astype(np.float32) * np.float32(0.25)
    return cols


def synthetic_rows_at(rows: np.ndarray, ncols: int = 5) -> Dict[str, np.ndarray]:
    """Columns of table T at the given global row numbers (any order)."""
    i = np.asarray(rows).astype(np.uint64)
    cols = {"k": splitmix64(42, i).view(np.int64)}
    if ncols >= 2:
        cols["v1"] = (splitmix64(43, i) % np.uint64(1000)).astype(np.int64)
    if ncols >= 3:
        cols["v2"] = i.astype(np.float64) * 1e-3
    if ncols >= 4:
        cols["v3"] = (i % np.uint64(100)).astype(np.int32)
    if ncols >= 5:
        cols["v4"] = (i % np.uint64(4096)).astype(np.float32) * np.float32(0.25)
    return cols


def synthetic_bucket(first_row: int, nrows: int, num_buckets: int, bucket: int, ncols: int = 5,
                     nthreads: int = 1) -> Dict[str, np.ndarray]:
    """What the index file of ``bucket`` must hold for createIndex(T[first_row : first_row + nrows], indexed = k):
    the table's rows with pmod(hash(k), num_buckets) == bucket, ordered by (k, source row) -- without materialising T."""
    n = lib().hso_synth_bucket_rows(first_row, nrows, num_buckets, bucket, None, 0, nthreads)
    rows = np.empty(n, dtype=np.int64)
    lib().hso_synth_bucket_rows(first_row, nrows, num_buckets, bucket, rows.ctypes.data, n, nthreads)
    k = splitmix64(42, rows.astype(np.uint64)).view(np.int64)
    order = np.argsort(k, kind="stable")
    return synthetic_rows_at(rows[order], ncols)


# ----------------------------------------------------------------------------------------------
# whole-path restatement (pyarrow codec)
# ----------------------------------------------------------------------------------------------

def bucket_file_name(bucket: int, job_uuid: str, task: Optional[int] = None, codec: str = "") -> str:
    """Spark FileFormatWriter / BucketingUtils naming: part-<task>-<uuid>_<bucket>.c000[.codec].parquet
    (asserted by T/index/IndexManagerTest.scala:259,738 and parsed by OptimizeAction.scala:110)."""
    task = bucket if task is None else task
    ext = f".{codec}" if codec else ""
    return f"part-{task:05d}-{job_uuid}_{bucket:05d}.c000{ext}.parquet"


def index_rows(table_cols: Dict[str, np.ndarray], indexed: Sequence[str], included: Sequence[str], num_buckets: int,
               valids: Optional[Dict[str, np.ndarray]] = None, nthreads: int = 1):
    """Project + bucket + sort decoded columns.  Returns (perm, bucket_offsets, column order)."""
    order = list(indexed) + list(included)  # CoveringIndex.scala:149
    keys = [table_cols[c] for c in indexed]
    kvalid = None if valids is None else [valids.get(c) for c in indexed]
    b = bucket_ids(keys, num_buckets, kvalid, nthreads)
    perm, offs = sort_perm(keys, num_buckets, b, kvalid, nthreads)
    return perm, offs, order


def create_index(src_files: Sequence[str], indexed: Sequence[str], included: Sequence[str], num_buckets: int,
                 out_dir: str, job_uuid: Optional[str] = None, nthreads: int = 1, compression: str = "NONE") -> List[str]:
    """CPU createIndex data path: scan -> project -> hash-repartition -> sort -> one Parquet file per non-empty bucket."""
    import pyarrow as pa
    import pyarrow.parquet as pq

    order = list(indexed) + list(included)
    tables = [pq.read_table(f, columns=order, use_threads=nthreads > 1) for f in src_files]
    t = pa.concat_tables(tables).combine_chunks()
    cols, valids = {}, {}
    for name in order:
        arr = t.column(name).chunk(0) if t.num_rows else pa.array([], type=t.schema.field(name).type)
        if arr.null_count:
            valids[name] = np.asarray(arr.is_valid())
            arr = arr.fill_null(0)
        cols[name] = np.asarray(arr)
    perm, offs, _ = index_rows(cols, indexed, included, num_buckets, valids or None, nthreads)
    os.makedirs(out_dir, exist_ok=True)
    job_uuid = job_uuid or str(uuid.uuid4())
    t = t.select(order)
    out = []
    for b in range(num_buckets):
        lo, hi = int(offs[b]), int(offs[b + 1])
        if hi == lo:
            continue  # no file for an empty bucket
        part = t.take(pa.array(perm[lo:hi]))
        path = os.path.join(out_dir, bucket_file_name(b, job_uuid, codec="" if compression == "NONE" else compression.lower()))
        pq.write_table(part, path, compression=compression, use_dictionary=False, data_page_version="1.0")
        out.append(path)
    return out

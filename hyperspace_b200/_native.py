"""ctypes binding of libhs_gpu.so (the C ABI in include/hs_gpu.h).

This is the Python counterpart of the JNI stub shown in INTEGRATION.md.  There is no CPU fallback: if the shared
library is missing, or no CUDA device is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HS_GPU_LIB: load another build of the same library (kernel A/B experiments); there is still no CPU fallback
LIB_PATH = os.environ.get("HS_GPU_LIB") or os.path.join(_HERE, "lib", "libhs_gpu.so")

HS_OK, HS_EINVAL, HS_ENODEVICE, HS_ECUDA, HS_EFORMAT, HS_EIO, HS_EUNSUPPORTED, HS_ENOMEM, HS_ECOMM = 0, -1, -2, -3, -4, -5, -6, -7, -8
HS_TYPE_INT32, HS_TYPE_INT64, HS_TYPE_FLOAT, HS_TYPE_DOUBLE, HS_TYPE_BOOL, HS_TYPE_STRING = range(6)
HS_SAVE_OVERWRITE, HS_SAVE_APPEND = 0, 1
HS_OUT_FILES, HS_OUT_HOST, HS_OUT_DEVICE = 0, 1, 2
HS_CODEC_UNCOMPRESSED, HS_CODEC_SNAPPY = 0, 1

_NP_OF_TYPE = {HS_TYPE_INT32: np.int32, HS_TYPE_INT64: np.int64, HS_TYPE_FLOAT: np.float32, HS_TYPE_DOUBLE: np.float64,
               HS_TYPE_BOOL: np.uint8}
_TYPE_OF_NP = {np.dtype(np.int32): HS_TYPE_INT32, np.dtype(np.int64): HS_TYPE_INT64, np.dtype(np.float32): HS_TYPE_FLOAT,
               np.dtype(np.float64): HS_TYPE_DOUBLE, np.dtype(np.bool_): HS_TYPE_BOOL, np.dtype(np.uint8): HS_TYPE_BOOL}


class HyperspaceGpuError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libhs_gpu error {code}: {message}")
        self.code = code
        self.message = message


class SourceFile(C.Structure):
    _fields_ = [("path", C.c_char_p), ("data", C.c_void_p), ("size", C.c_uint64), ("file_id", C.c_int64),
                ("on_device", C.c_int32), ("reserved", C.c_int32)]


class IndexSpec(C.Structure):
    _fields_ = [("files", C.POINTER(SourceFile)), ("n_files", C.c_int32),
                ("indexed_columns", C.POINTER(C.c_char_p)), ("n_indexed", C.c_int32),
                ("included_columns", C.POINTER(C.c_char_p)), ("n_included", C.c_int32),
                ("num_buckets", C.c_int32), ("save_mode", C.c_int32), ("output", C.c_int32), ("lineage", C.c_int32),
                ("out_dir", C.c_char_p), ("job_uuid", C.c_char_p),
                ("rows_per_page", C.c_int64), ("rows_per_row_group", C.c_int64),
                ("deleted_file_ids", C.POINTER(C.c_int64)), ("n_deleted_file_ids", C.c_int32), ("disable_dictionary", C.c_int32),
                ("compression", C.c_int32), ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("rows_in", C.c_int64), ("rows_out", C.c_int64), ("bytes_in", C.c_int64), ("bytes_out", C.c_int64),
                ("bytes_exchanged", C.c_int64), ("files_out", C.c_int32), ("gpu_launches", C.c_int32),
                ("ms_total", C.c_float), ("ms_h2d", C.c_float), ("ms_plan", C.c_float), ("ms_decode", C.c_float),
                ("ms_hash", C.c_float), ("ms_partition", C.c_float), ("ms_exchange", C.c_float), ("ms_sort", C.c_float),
                ("ms_gather", C.c_float), ("ms_encode", C.c_float), ("ms_d2h", C.c_float), ("ms_write", C.c_float)]

    def as_dict(self) -> Dict[str, float]:
        return {n: getattr(self, n) for n, _ in self._fields_}


class ScanSpec(C.Structure):
    _fields_ = [("files", C.POINTER(SourceFile)), ("n_files", C.c_int32), ("sorted_on_key", C.c_int32),
                ("key_column", C.c_char_p), ("projected_columns", C.POINTER(C.c_char_p)), ("n_projected", C.c_int32),
                ("has_lo", C.c_int32), ("has_hi", C.c_int32), ("lo", C.c_int64), ("hi", C.c_int64),
                ("deleted_file_ids", C.POINTER(C.c_int64)), ("n_deleted_file_ids", C.c_int32), ("output", C.c_int32),
                ("lo_bytes", C.c_char_p), ("hi_bytes", C.c_char_p), ("lo_len", C.c_uint32), ("hi_len", C.c_uint32)]


class JoinSpec(C.Structure):
    _fields_ = [("left_files", C.POINTER(SourceFile)), ("n_left", C.c_int32),
                ("right_files", C.POINTER(SourceFile)), ("n_right", C.c_int32),
                ("left_buckets", C.POINTER(C.c_int32)), ("right_buckets", C.POINTER(C.c_int32)),
                ("num_buckets", C.c_int32), ("output", C.c_int32),
                ("left_key", C.c_char_p), ("right_key", C.c_char_p),
                ("left_columns", C.POINTER(C.c_char_p)), ("n_left_columns", C.c_int32),
                ("right_columns", C.POINTER(C.c_char_p)), ("n_right_columns", C.c_int32)]


class VerifyReport(C.Structure):
    _fields_ = [("rows", C.c_int64), ("bucket_mismatches", C.c_int64), ("order_violations", C.c_int64),
                ("row_checksum", C.c_uint64), ("column_checksum", C.c_uint64 * 16), ("n_columns", C.c_int32),
                ("reserved", C.c_int32)]

    def as_dict(self) -> Dict[str, object]:
        return {"rows": self.rows, "bucket_mismatches": self.bucket_mismatches, "order_violations": self.order_violations,
                "row_checksum": int(self.row_checksum),
                "column_checksum": [int(self.column_checksum[i]) for i in range(self.n_columns)]}


class HostColumn(C.Structure):
    _fields_ = [("type", C.c_int32), ("reserved", C.c_int32), ("data", C.c_void_p), ("valid", C.c_void_p)]


# every symbol include/hs_gpu.h declares; tests/test_abi.py checks the library exports all of them
EXPORTED_SYMBOLS = [
    "hs_abi_version", "hs_build_info", "hs_init", "hs_shutdown", "hs_trim", "hs_host_alloc", "hs_host_free", "hs_profile_enable", "hs_profile_report",
    "hs_comm_unique_id", "hs_comm_init", "hs_create_index", "hs_result_num_files", "hs_result_file", "hs_result_free",
    "hs_filter_scan", "hs_bucket_join", "hs_batch_num_rows", "hs_batch_on_device", "hs_batch_num_columns", "hs_batch_column", "hs_batch_free",
    "hs_k_bucket_ids", "hs_k_sort_perm", "hs_synth_table",
    "hs_stage_sources", "hs_staged_num_files", "hs_staged_file", "hs_staged_wait", "hs_staged_free",
    "hs_create_index_async", "hs_pending_wait", "hs_pending_cancel", "hs_verify_index", "hs_synth_checksum",
    "hs_synth_table_ex", "hs_k_snappy_compress", "hs_k_snappy_decompress", "hs_batch_string_offsets",
]

_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Loads libhs_gpu.so; raises if it has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HyperspaceGpuError(HS_ENODEVICE, f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                                               "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    err = (C.c_char_p, C.c_size_t)
    L.hs_abi_version.restype = C.c_int
    L.hs_build_info.restype = C.c_char_p
    L.hs_init.restype = C.c_int
    L.hs_init.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p), *err]
    L.hs_shutdown.restype = None
    L.hs_shutdown.argtypes = [C.c_void_p]
    L.hs_trim.restype = None
    L.hs_trim.argtypes = [C.c_void_p]
    L.hs_host_alloc.restype = C.c_void_p
    L.hs_host_alloc.argtypes = [C.c_void_p, C.c_size_t]
    L.hs_host_free.restype = None
    L.hs_host_free.argtypes = [C.c_void_p, C.c_void_p]
    L.hs_profile_enable.restype = None
    L.hs_profile_enable.argtypes = [C.c_void_p, C.c_int]
    L.hs_profile_report.restype = C.c_int
    L.hs_profile_report.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.hs_comm_unique_id.restype = C.c_int
    L.hs_comm_unique_id.argtypes = [C.c_void_p, *err]
    L.hs_comm_init.restype = C.c_int
    L.hs_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, *err]
    L.hs_create_index.restype = C.c_int
    L.hs_create_index.argtypes = [C.c_void_p, C.POINTER(IndexSpec), C.POINTER(C.c_void_p), C.POINTER(Stats), *err]
    L.hs_result_num_files.restype = C.c_int32
    L.hs_result_num_files.argtypes = [C.c_void_p]
    L.hs_result_file.restype = C.c_int
    L.hs_result_file.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_char_p), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]
    L.hs_result_free.restype = None
    L.hs_result_free.argtypes = [C.c_void_p]
    L.hs_filter_scan.restype = C.c_int
    L.hs_filter_scan.argtypes = [C.c_void_p, C.POINTER(ScanSpec), C.POINTER(C.c_void_p), C.POINTER(Stats), *err]
    L.hs_bucket_join.restype = C.c_int
    L.hs_bucket_join.argtypes = [C.c_void_p, C.POINTER(JoinSpec), C.POINTER(C.c_void_p), C.POINTER(Stats), *err]
    L.hs_batch_num_rows.restype = C.c_int64
    L.hs_batch_num_rows.argtypes = [C.c_void_p]
    L.hs_batch_on_device.restype = C.c_int32
    L.hs_batch_on_device.argtypes = [C.c_void_p]
    L.hs_batch_num_columns.restype = C.c_int32
    L.hs_batch_num_columns.argtypes = [C.c_void_p]
    L.hs_batch_column.restype = C.c_int
    L.hs_batch_column.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_void_p)]
    L.hs_batch_free.restype = None
    L.hs_batch_free.argtypes = [C.c_void_p]
    L.hs_k_bucket_ids.restype = C.c_int
    L.hs_k_bucket_ids.argtypes = [C.c_void_p, C.POINTER(HostColumn), C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, *err]
    L.hs_k_sort_perm.restype = C.c_int
    L.hs_k_sort_perm.argtypes = [C.c_void_p, C.POINTER(HostColumn), C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, *err]
    L.hs_synth_table.restype = C.c_int
    L.hs_synth_table.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.POINTER(C.c_void_p), *err]
    L.hs_stage_sources.restype = C.c_int
    L.hs_stage_sources.argtypes = [C.c_void_p, C.POINTER(SourceFile), C.c_int32, C.POINTER(C.c_void_p), *err]
    L.hs_staged_num_files.restype = C.c_int32
    L.hs_staged_num_files.argtypes = [C.c_void_p]
    L.hs_staged_file.restype = C.c_int
    L.hs_staged_file.argtypes = [C.c_void_p, C.c_int32, C.POINTER(SourceFile)]
    L.hs_staged_wait.restype = C.c_int
    L.hs_staged_wait.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.hs_staged_free.restype = None
    L.hs_staged_free.argtypes = [C.c_void_p]
    L.hs_create_index_async.restype = C.c_int
    L.hs_create_index_async.argtypes = [C.c_void_p, C.POINTER(IndexSpec), C.POINTER(C.c_void_p), *err]
    L.hs_pending_wait.restype = C.c_int
    L.hs_pending_wait.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(Stats), *err]
    L.hs_pending_cancel.restype = None
    L.hs_pending_cancel.argtypes = [C.c_void_p]
    L.hs_verify_index.restype = C.c_int
    L.hs_verify_index.argtypes = [C.c_void_p, C.POINTER(SourceFile), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_char_p), C.c_int32,
                                  C.POINTER(C.c_char_p), C.c_int32, C.c_int32, C.POINTER(VerifyReport), *err]
    L.hs_synth_checksum.restype = C.c_int
    L.hs_synth_checksum.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.POINTER(VerifyReport), *err]
    L.hs_synth_table_ex.restype = C.c_int
    L.hs_synth_table_ex.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                    C.POINTER(C.c_void_p), *err]
    L.hs_k_snappy_compress.restype = C.c_int
    L.hs_k_snappy_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), *err]
    L.hs_batch_string_offsets.restype = C.c_int
    L.hs_batch_string_offsets.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
    L.hs_k_snappy_decompress.restype = C.c_int
    L.hs_k_snappy_decompress.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_int32), *err]
    if L.hs_abi_version() != 1:
        raise HyperspaceGpuError(HS_EINVAL, f"ABI version mismatch: library {L.hs_abi_version()}, binding 1")
    _lib = L
    return L


def _check(rc: int, err) -> None:
    if rc != HS_OK:
        raise HyperspaceGpuError(rc, err.value.decode("utf-8", "replace"))


def _cstr_array(names: Sequence[str]):
    arr = (C.c_char_p * max(1, len(names)))()
    for i, n in enumerate(names):
        arr[i] = n.encode()
    return arr


@dataclass
class FileImage:
    """One Parquet file handed to the engine: a path, or an in-memory image (host bytes / numpy uint8 / device pointer)."""
    path: Optional[str] = None
    data: Optional[object] = None   # bytes, numpy uint8 array, or int (raw pointer)
    size: int = 0
    file_id: int = -1
    on_device: bool = False


def _source_array(files: Sequence[FileImage]):
    keep = []
    arr = (SourceFile * max(1, len(files)))()
    for i, f in enumerate(files):
        arr[i].path = f.path.encode() if f.path else None
        arr[i].file_id = f.file_id
        arr[i].on_device = 1 if f.on_device else 0
        if f.data is None:
            arr[i].data = None
            arr[i].size = 0
        elif isinstance(f.data, int):
            arr[i].data = f.data
            arr[i].size = f.size
        elif isinstance(f.data, np.ndarray):
            a = np.ascontiguousarray(f.data).view(np.uint8)
            keep.append(a)
            arr[i].data = a.ctypes.data
            arr[i].size = a.nbytes
        else:
            b = bytes(f.data)
            buf = C.create_string_buffer(b, len(b))
            keep.append(buf)
            arr[i].data = C.addressof(buf)
            arr[i].size = len(b)
    return arr, keep


@dataclass
class ResultFile:
    bucket: int
    name: str
    ptr: Optional[int]
    size: int
    rows: int


class IndexResult:
    """Owns an hs_index_result handle."""

    def __init__(self, ctx: "Context", handle: int, output: int):
        self._ctx, self._h, self.output = ctx, handle, output
        ctx._results.add(self)  # a result must not outlive its context: Context.close() frees the stragglers
        L = load_library()
        self.files: List[ResultFile] = []
        for i in range(L.hs_result_num_files(handle)):
            b, nm, p, sz, rows = C.c_int32(), C.c_char_p(), C.c_void_p(), C.c_uint64(), C.c_int64()
            L.hs_result_file(handle, i, C.byref(b), C.byref(nm), C.byref(p), C.byref(sz), C.byref(rows))
            self.files.append(ResultFile(b.value, nm.value.decode(), p.value, sz.value, rows.value))

    def host_bytes(self, i: int) -> bytes:
        assert self.output == HS_OUT_HOST
        f = self.files[i]
        return C.string_at(f.ptr, f.size)

    def host_view(self, i: int) -> np.ndarray:
        assert self.output == HS_OUT_HOST
        f = self.files[i]
        return np.ctypeslib.as_array((C.c_uint8 * f.size).from_address(f.ptr))

    def as_sources(self) -> List[FileImage]:
        """The result's in-memory images as engine inputs (host or device)."""
        return [FileImage(path=f.name, data=f.ptr, size=f.size, on_device=self.output == HS_OUT_DEVICE) for f in self.files]

    def free(self) -> None:
        if self._h:
            load_library().hs_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Staged:
    """Source file images on their way to device memory (hs_stage_sources); ``as_sources()`` feeds create_index[_async]."""

    def __init__(self, ctx: "Context", handle: int, keep):
        self._ctx, self._h, self._keep = ctx, handle, keep
        ctx._results.add(self)
        L = load_library()
        self.files: List[FileImage] = []
        for i in range(L.hs_staged_num_files(handle)):
            sf = SourceFile()
            L.hs_staged_file(handle, i, C.byref(sf))
            self.files.append(FileImage(path=sf.path.decode() if sf.path else None, data=sf.data, size=sf.size,
                                        file_id=sf.file_id, on_device=True))

    def as_sources(self) -> List[FileImage]:
        return list(self.files)

    def wait(self) -> float:
        """Blocks until the copies have completed; returns their duration on the H2D stream in ms."""
        ms = C.c_float(0)
        if self._h and load_library().hs_staged_wait(self._h, C.byref(ms)) != HS_OK:
            raise HyperspaceGpuError(HS_ECUDA, "hs_staged_wait failed")
        return float(ms.value)

    def free(self) -> None:
        if self._h:
            load_library().hs_staged_free(self._h)
            self._h = None
            self._keep = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Pending:
    """A createIndex whose index files are draining to the host (hs_create_index_async); ``wait()`` yields the result."""

    def __init__(self, ctx: "Context", handle: int, output: int, keep):
        self._ctx, self._h, self.output, self._keep = ctx, handle, output, keep
        ctx._results.add(self)

    def wait(self) -> Tuple[IndexResult, Dict[str, float]]:
        assert self._h, "already waited for"
        res, st = C.c_void_p(), Stats()
        err = C.create_string_buffer(1024)
        h, self._h = self._h, None
        _check(load_library().hs_pending_wait(h, C.byref(res), C.byref(st), err, len(err)), err)
        self._keep = None
        return IndexResult(self._ctx, res.value, self.output), st.as_dict()

    def free(self) -> None:
        if self._h:
            load_library().hs_pending_cancel(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Batch:
    """An hs_batch: result columns in pinned host memory, exposed as zero-copy numpy views.  The views are valid until
    ``free()`` (or garbage collection of the Batch); copy them if they must outlive it."""

    def __init__(self, handle: int, ctx: "Context" = None):
        L = load_library()
        self._h = handle
        if ctx is not None:
            ctx._results.add(self)
        self.num_rows = L.hs_batch_num_rows(handle)
        self.on_device = bool(L.hs_batch_on_device(handle))
        self.columns: List[Tuple[str, np.ndarray, Optional[np.ndarray]]] = []
        self.device_columns: List[Tuple[str, int, int]] = []  # (name, HS_TYPE_*, device pointer) when on_device
        n = self.num_rows
        for i in range(L.hs_batch_num_columns(handle)):
            nm, ty, d, v = C.c_char_p(), C.c_int32(), C.c_void_p(), C.c_void_p()
            L.hs_batch_column(handle, i, C.byref(nm), C.byref(ty), C.byref(d), C.byref(v))
            if self.on_device:
                self.device_columns.append((nm.value.decode(), ty.value, d.value))
                continue
            if ty.value == HS_TYPE_STRING:  # bytes back to back + num_rows + 1 offsets -> an object array of bytes
                off_p, total = C.c_void_p(), C.c_uint64()
                L.hs_batch_string_offsets(handle, i, C.byref(off_p), C.byref(total))
                offs = np.ctypeslib.as_array((C.c_uint64 * (n + 1)).from_address(off_p.value)) if n else np.zeros(1, np.uint64)
                blob = C.string_at(d.value, total.value) if total.value else b""
                data = np.empty(n, dtype=object)
                o = offs.tolist()
                for r in range(n):
                    data[r] = blob[o[r]:o[r + 1]]
                valid = None
                if v.value:
                    valid = np.ctypeslib.as_array((C.c_uint8 * n).from_address(v.value)).copy() if n else np.empty(0, np.uint8)
                self.columns.append((nm.value.decode(), data, valid))
                continue
            dt = np.dtype(_NP_OF_TYPE[ty.value])
            if n:
                data = np.ctypeslib.as_array((C.c_uint8 * (n * dt.itemsize)).from_address(d.value)).view(dt)
            else:
                data = np.empty(0, dt)
            valid = None
            if v.value:
                valid = np.ctypeslib.as_array((C.c_uint8 * n).from_address(v.value)) if n else np.empty(0, np.uint8)
            self.columns.append((nm.value.decode(), data, valid))

    def column(self, name: str) -> np.ndarray:
        for n, d, _ in self.columns:
            if n == name:
                return d
        raise KeyError(name)

    def free(self) -> None:
        if self._h:
            self.columns = []
            load_library().hs_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One GPU, one stream (hs_ctx)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        L = load_library()
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        _check(L.hs_init(device, stream, C.byref(h), err, len(err)), err)
        self._h = h.value
        self.device = device
        self.rank, self.world = 0, 1
        self._results = weakref.WeakSet()

    def close(self) -> None:
        if self._h:
            for r in list(self._results):
                r.free()
            load_library().hs_shutdown(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def trim(self) -> None:
        load_library().hs_trim(self._h)

    def profile_enable(self, on: bool = True) -> None:
        load_library().hs_profile_enable(self._h, 1 if on else 0)

    def profile_report(self) -> Dict[str, Dict[str, float]]:
        import json

        buf = C.create_string_buffer(1 << 16)
        rc = load_library().hs_profile_report(self._h, buf, len(buf))
        if rc != HS_OK:
            raise HyperspaceGpuError(rc, "hs_profile_report failed")
        return json.loads(buf.value.decode())

    def host_alloc(self, nbytes: int) -> np.ndarray:
        """Pinned host buffer (owned by the context's pool) as a numpy uint8 view."""
        p = load_library().hs_host_alloc(self._h, nbytes)
        if not p:
            raise HyperspaceGpuError(HS_ENOMEM, f"cannot pin {nbytes} bytes")
        return np.ctypeslib.as_array((C.c_uint8 * nbytes).from_address(p))

    def host_free(self, arr: np.ndarray) -> None:
        load_library().hs_host_free(self._h, arr.ctypes.data)

    # ---- multi-GPU --------------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        err = C.create_string_buffer(1024)
        _check(load_library().hs_comm_unique_id(buf, err, len(err)), err)
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes) -> None:
        err = C.create_string_buffer(1024)
        idbuf = C.create_string_buffer(unique_id, 128)
        _check(load_library().hs_comm_init(self._h, rank, world, idbuf, err, len(err)), err)
        self.rank, self.world = rank, world

    # ---- write side ----------------------------------------------------------------------------------
    def stage_sources(self, files: Sequence[FileImage]) -> Staged:
        """Starts copying host / file-system Parquet images to the device on the ctx's H2D copy stream."""
        src, keep = _source_array(files)
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        _check(load_library().hs_stage_sources(self._h, src, len(files), C.byref(h), err, len(err)), err)
        return Staged(self, h.value, (src, keep, list(files)))

    def create_index_async(self, files: Sequence[FileImage], indexed: Sequence[str], included: Sequence[str], num_buckets: int,
                           **kw) -> Pending:
        """create_index up to the encoded files in device memory; the device->host copy continues on the D2H copy stream."""
        spec, keep = self._index_spec(files, indexed, included, num_buckets, **kw)
        h = C.c_void_p()
        err = C.create_string_buffer(1024)
        _check(load_library().hs_create_index_async(self._h, C.byref(spec), C.byref(h), err, len(err)), err)
        return Pending(self, h.value, spec.output, keep)

    def verify_index(self, files: Sequence[FileImage], buckets: Sequence[int], indexed: Sequence[str], included: Sequence[str],
                     num_buckets: int) -> Dict[str, object]:
        src, keep = _source_array(files)
        ic, nc = _cstr_array(indexed), _cstr_array(included)
        bk = (C.c_int32 * max(1, len(buckets)))(*buckets)
        rep = VerifyReport()
        err = C.create_string_buffer(1024)
        _check(load_library().hs_verify_index(self._h, src, bk, len(files), ic, len(indexed), nc, len(included), num_buckets,
                                              C.byref(rep), err, len(err)), err)
        return rep.as_dict()

    def synth_checksum(self, first_row: int, nrows: int, ncols: int = 5) -> Dict[str, object]:
        rep = VerifyReport()
        err = C.create_string_buffer(1024)
        _check(load_library().hs_synth_checksum(self._h, first_row, nrows, ncols, C.byref(rep), err, len(err)), err)
        return rep.as_dict()

    def _index_spec(self, files, indexed, included, num_buckets, out_dir=None, output=HS_OUT_FILES, job_uuid=None,
                    save_mode=HS_SAVE_OVERWRITE, lineage=False, deleted_file_ids=(), rows_per_page=0, rows_per_row_group=0,
                    dictionary=True, compression=HS_CODEC_UNCOMPRESSED):
        src, keep = _source_array(files)
        ic, nc = _cstr_array(indexed), _cstr_array(included)
        spec = IndexSpec()
        spec.files, spec.n_files = src, len(files)
        spec.indexed_columns, spec.n_indexed = ic, len(indexed)
        spec.included_columns, spec.n_included = nc, len(included)
        spec.num_buckets, spec.save_mode, spec.output, spec.lineage = num_buckets, save_mode, output, 1 if lineage else 0
        spec.out_dir = out_dir.encode() if out_dir else None
        spec.job_uuid = job_uuid.encode() if job_uuid else None
        spec.rows_per_page, spec.rows_per_row_group = rows_per_page, rows_per_row_group
        dl = (C.c_int64 * max(1, len(deleted_file_ids)))(*deleted_file_ids)
        spec.deleted_file_ids, spec.n_deleted_file_ids = dl, len(deleted_file_ids)
        spec.disable_dictionary = 0 if dictionary else 1
        spec.compression = compression
        return spec, (src, keep, ic, nc, dl)

    def create_index(self, files: Sequence[FileImage], indexed: Sequence[str], included: Sequence[str], num_buckets: int,
                     **kw) -> Tuple[IndexResult, Dict[str, float]]:
        """hs_create_index.  Keywords: out_dir, output (HS_OUT_*), job_uuid, save_mode, lineage, deleted_file_ids,
        rows_per_page, rows_per_row_group, dictionary, compression (HS_CODEC_*)."""
        spec, keep = self._index_spec(files, indexed, included, num_buckets, **kw)
        res, st = C.c_void_p(), Stats()
        err = C.create_string_buffer(1024)
        _check(load_library().hs_create_index(self._h, C.byref(spec), C.byref(res), C.byref(st), err, len(err)), err)
        return IndexResult(self, res.value, spec.output), st.as_dict()

    def synth_table(self, first_row: int, nrows: int, ncols: int = 5, n_files: int = 1, row_groups_per_file: int = 1,
                    output: int = HS_OUT_HOST, dictionary: bool = True, compression: int = HS_CODEC_UNCOMPRESSED) -> IndexResult:
        L = load_library()
        res = C.c_void_p()
        err = C.create_string_buffer(1024)
        _check(L.hs_synth_table_ex(self._h, first_row, nrows, ncols, n_files, row_groups_per_file, 1 if dictionary else 0,
                                   compression, output, C.byref(res), err, len(err)), err)
        return IndexResult(self, res.value, output)

    def k_snappy_compress(self, data: bytes) -> bytes:
        """The GPU page compressor on a host buffer (parity tests: any Snappy decoder must give `data` back)."""
        cap = 64 + len(data) + len(data) // 5
        out = C.create_string_buffer(cap)
        n = C.c_uint64(0)
        err = C.create_string_buffer(1024)
        _check(load_library().hs_k_snappy_compress(self._h, data, len(data), out, cap, C.byref(n), err, len(err)), err)
        return out.raw[:n.value]

    def k_snappy_decompress(self, stream: bytes, uncompressed_len: int) -> Tuple[bytes, bool]:
        """The GPU page decompressor on one raw Snappy stream; returns (bytes, decoded front-to-back by one warp?)."""
        out = C.create_string_buffer(max(1, uncompressed_len))
        seq = C.c_int32(0)
        err = C.create_string_buffer(1024)
        _check(load_library().hs_k_snappy_decompress(self._h, stream, len(stream), out, uncompressed_len, C.byref(seq), err,
                                                     len(err)), err)
        return out.raw[:uncompressed_len], bool(seq.value)

    # ---- read side ----------------------------------------------------------------------------------
    def filter_scan(self, files: Sequence[FileImage], key: str, projected: Sequence[str], lo=None, hi=None, sorted_on_key: bool = True, deleted_file_ids: Sequence[int] = (),
                    output: int = HS_OUT_HOST) -> Tuple[Batch, Dict[str, float]]:
        L = load_library()
        src, keep = _source_array(files)
        pc = _cstr_array(projected)
        spec = ScanSpec()
        spec.files, spec.n_files, spec.sorted_on_key = src, len(files), 1 if sorted_on_key else 0
        spec.key_column = key.encode()
        spec.projected_columns, spec.n_projected = pc, len(projected)
        spec.has_lo, spec.has_hi = int(lo is not None), int(hi is not None)
        if isinstance(lo, (str, bytes)) or isinstance(hi, (str, bytes)):  # string / binary key: bounds as bytes
            lob = (lo.encode() if isinstance(lo, str) else lo) if lo is not None else b""
            hib = (hi.encode() if isinstance(hi, str) else hi) if hi is not None else b""
            spec.lo_bytes, spec.lo_len, spec.hi_bytes, spec.hi_len = lob, len(lob), hib, len(hib)
            spec.lo, spec.hi = 0, 0
        else:
            spec.lo, spec.hi = lo or 0, hi or 0
        dl = (C.c_int64 * max(1, len(deleted_file_ids)))(*deleted_file_ids)
        spec.deleted_file_ids, spec.n_deleted_file_ids = dl, len(deleted_file_ids)
        spec.output = output
        res, st = C.c_void_p(), Stats()
        err = C.create_string_buffer(1024)
        _check(L.hs_filter_scan(self._h, C.byref(spec), C.byref(res), C.byref(st), err, len(err)), err)
        return Batch(res.value, self), st.as_dict()

    def bucket_join(self, left: Sequence[FileImage], left_buckets: Sequence[int], right: Sequence[FileImage],
                    right_buckets: Sequence[int], num_buckets: int, left_key: str, right_key: str,
                    left_columns: Sequence[str], right_columns: Sequence[str], output: int = HS_OUT_HOST
                    ) -> Tuple[Batch, Dict[str, float]]:
        L = load_library()
        ls, k1 = _source_array(left)
        rs, k2 = _source_array(right)
        lc, rc = _cstr_array(left_columns), _cstr_array(right_columns)
        lb = (C.c_int32 * max(1, len(left_buckets)))(*left_buckets)
        rb = (C.c_int32 * max(1, len(right_buckets)))(*right_buckets)
        spec = JoinSpec()
        spec.left_files, spec.n_left, spec.right_files, spec.n_right = ls, len(left), rs, len(right)
        spec.left_buckets, spec.right_buckets, spec.num_buckets = lb, rb, num_buckets
        spec.left_key, spec.right_key = left_key.encode(), right_key.encode()
        spec.left_columns, spec.n_left_columns = lc, len(left_columns)
        spec.right_columns, spec.n_right_columns = rc, len(right_columns)
        spec.output = output
        res, st = C.c_void_p(), Stats()
        err = C.create_string_buffer(1024)
        _check(L.hs_bucket_join(self._h, C.byref(spec), C.byref(res), C.byref(st), err, len(err)), err)
        return Batch(res.value, self), st.as_dict()

    # ---- kernel-level entry points ----------------------------------------------------------------------------------
    @staticmethod
    def _host_columns(cols: Sequence[np.ndarray], valids):
        keep = []
        arr = (HostColumn * len(cols))()
        for i, c in enumerate(cols):
            c = np.ascontiguousarray(c)
            keep.append(c)
            arr[i].type = _TYPE_OF_NP[c.dtype]
            arr[i].data = c.ctypes.data
            v = None if valids is None else valids[i]
            if v is not None:
                v = np.ascontiguousarray(v.astype(np.uint8))
                keep.append(v)
                arr[i].valid = v.ctypes.data
            else:
                arr[i].valid = None
        return arr, keep

    def k_bucket_ids(self, keys: Sequence[np.ndarray], num_buckets: int, valids=None) -> Tuple[np.ndarray, np.ndarray]:
        n = len(keys[0])
        arr, keep = self._host_columns(keys, valids)
        out = np.empty(n, dtype=np.int32)
        hist = np.zeros(num_buckets, dtype=np.int64)
        err = C.create_string_buffer(1024)
        _check(load_library().hs_k_bucket_ids(self._h, arr, len(keys), n, num_buckets, out.ctypes.data, hist.ctypes.data,
                                              err, len(err)), err)
        return out, hist

    def k_sort_perm(self, keys: Sequence[np.ndarray], num_buckets: int, valids=None) -> Tuple[np.ndarray, np.ndarray]:
        n = len(keys[0])
        arr, keep = self._host_columns(keys, valids)
        perm = np.empty(n, dtype=np.int64)
        offs = np.empty(num_buckets + 1, dtype=np.int64)
        err = C.create_string_buffer(1024)
        _check(load_library().hs_k_sort_perm(self._h, arr, len(keys), n, num_buckets, perm.ctypes.data, offs.ctypes.data,
                                             err, len(err)), err)
        return perm, offs

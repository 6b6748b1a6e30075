"""Minimal session / DataFrame layer the Hyperspace API needs around the GPU engine.

The reference plugs into Spark: ``spark.read.parquet`` gives the source relation, the optimizer hook
(``ApplyHyperspace``, src/main/scala/com/microsoft/hyperspace/index/rules/ApplyHyperspace.scala:45-66) swaps relations for
index scans, Spark executes.  Spark is not available here, so this module provides just enough of that surface for
notebooks of the shape used in the reference's docs/tests to run unchanged against the GPU engine:

    session = HyperspaceSession()
    df = session.read.parquet("/data/t")
    hs = Hyperspace(session); hs.createIndex(df, IndexConfig("idx", ["k"], ["v1"]))
    session.enableHyperspace()
    df.filter(col("k").between(0, 100)).select("k", "v1").collect()
    a.join(b, on="k").select(...).collect()

Every scan, filter and join runs on the GPU through the C ABI (no CPU fallback); the plan layer only decides which
files the native call reads -- the same decision FilterIndexRule / JoinIndexRule make (hyperspace_b200/rules.py).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import log_entry as LE

# conf keys and defaults: src/main/scala/com/microsoft/hyperspace/index/IndexConstants.scala:21-170
INDEX_SYSTEM_PATH = "spark.hyperspace.system.path"
INDEX_NUM_BUCKETS = "spark.hyperspace.index.numBuckets"
INDEX_NUM_BUCKETS_LEGACY = "spark.hyperspace.index.num.buckets"
INDEX_NUM_BUCKETS_DEFAULT = 200
INDEX_LINEAGE_ENABLED = "spark.hyperspace.index.lineage.enabled"
INDEX_HYBRID_SCAN_ENABLED = "spark.hyperspace.index.hybridscan.enabled"
INDEX_HYBRID_SCAN_APPENDED_RATIO_THRESHOLD = "spark.hyperspace.index.hybridscan.maxAppendedRatio"
INDEX_HYBRID_SCAN_DELETED_RATIO_THRESHOLD = "spark.hyperspace.index.hybridscan.maxDeletedRatio"
OPTIMIZE_FILE_SIZE_THRESHOLD = "spark.hyperspace.index.optimize.fileSizeThreshold"
OPTIMIZE_FILE_SIZE_THRESHOLD_DEFAULT = 256 * 1024 * 1024
HYPERSPACE_ENABLED = "spark.hyperspace.enabled"  # session flag toggled by enableHyperspace()/disableHyperspace()


class RuntimeConf:
    """String key/value conf like ``spark.conf`` (typed getters: util/HyperspaceConf.scala:27-238)."""

    def __init__(self, values: Optional[Dict[str, str]] = None):
        self._v: Dict[str, str] = dict(values or {})

    def set(self, key: str, value) -> None:
        self._v[key] = str(value).lower() if isinstance(value, bool) else str(value)

    def get(self, key: str, default=None):
        return self._v.get(key, default)

    def unset(self, key: str) -> None:
        self._v.pop(key, None)

    def get_bool(self, key: str, default: bool) -> bool:
        return str(self._v.get(key, default)).lower() == "true"

    @property
    def num_buckets(self) -> int:
        """HyperspaceConf.numBucketsForIndex (util/HyperspaceConf.scala:88-93): new key, legacy key, then 200."""
        return int(self._v.get(INDEX_NUM_BUCKETS, self._v.get(INDEX_NUM_BUCKETS_LEGACY, INDEX_NUM_BUCKETS_DEFAULT)))

    @property
    def lineage_enabled(self) -> bool:
        return self.get_bool(INDEX_LINEAGE_ENABLED, False)

    @property
    def hybrid_scan_enabled(self) -> bool:
        return self.get_bool(INDEX_HYBRID_SCAN_ENABLED, False)

    @property
    def hybrid_scan_appended_ratio(self) -> float:
        return float(self._v.get(INDEX_HYBRID_SCAN_APPENDED_RATIO_THRESHOLD, 0.3))

    @property
    def hybrid_scan_deleted_ratio(self) -> float:
        return float(self._v.get(INDEX_HYBRID_SCAN_DELETED_RATIO_THRESHOLD, 0.2))

    @property
    def optimize_file_size_threshold(self) -> int:
        return int(self._v.get(OPTIMIZE_FILE_SIZE_THRESHOLD, OPTIMIZE_FILE_SIZE_THRESHOLD_DEFAULT))


# ---------------------------------------------------------------------------------------------------------------------
# expressions
# ---------------------------------------------------------------------------------------------------------------------

@dataclass
class Predicate:
    """Conjunction of inclusive integer bounds per column: {column: (lo or None, hi or None)}."""
    bounds: Dict[str, Tuple[Optional[int], Optional[int]]]

    def __and__(self, other: "Predicate") -> "Predicate":
        out = dict(self.bounds)
        for c, (lo, hi) in other.bounds.items():
            if c in out:
                l0, h0 = out[c]
                lo = l0 if lo is None else (lo if l0 is None else max(lo, l0))
                hi = h0 if hi is None else (hi if h0 is None else min(hi, h0))
            out[c] = (lo, hi)
        return Predicate(out)

    @property
    def columns(self) -> List[str]:
        return list(self.bounds)


def _as_bytes(v) -> bytes:
    return v.encode("utf-8") if isinstance(v, str) else bytes(v)


class Column:
    def __init__(self, name: str):
        self.name = name

    # integer key columns: a non-integral literal is rounded in the direction that keeps the predicate's meaning
    # (k < 1.5  <=>  k <= 1;  k >= 1.5  <=>  k >= 2).  String / binary literals give byte bounds in UTF8String order --
    # `col("Query") == "facebook"` is the predicate of the reference's own filter-rule tests (T/index/E2EHyperspaceRulesTest.scala).
    def __ge__(self, v):
        if isinstance(v, (str, bytes)):
            return Predicate({self.name: (_as_bytes(v), None)})
        return Predicate({self.name: (math.ceil(v), None)})

    def __gt__(self, v):
        if isinstance(v, (str, bytes)):
            return Predicate({self.name: (_as_bytes(v) + b"\x00", None)})  # the smallest value above v
        return Predicate({self.name: (math.floor(v) + 1, None)})

    def __le__(self, v):
        if isinstance(v, (str, bytes)):
            return Predicate({self.name: (None, _as_bytes(v))})
        return Predicate({self.name: (None, math.floor(v))})

    def __lt__(self, v):
        if isinstance(v, (str, bytes)):
            raise ValueError("a strict upper bound on a string column has no inclusive form: use <= or between")
        return Predicate({self.name: (None, math.ceil(v) - 1)})

    def __eq__(self, v):  # noqa: A003
        if isinstance(v, (str, bytes)):
            return Predicate({self.name: (_as_bytes(v), _as_bytes(v))})
        if v != math.floor(v):
            return Predicate({self.name: (1, 0)})  # an integer never equals a fraction: empty range
        return Predicate({self.name: (int(v), int(v))})

    def between(self, lo, hi):
        if isinstance(lo, (str, bytes)) or isinstance(hi, (str, bytes)):
            return Predicate({self.name: (_as_bytes(lo), _as_bytes(hi))})
        return Predicate({self.name: (math.ceil(lo), math.floor(hi))})


def col(name: str) -> Column:
    return Column(name)


# ---------------------------------------------------------------------------------------------------------------------
# logical plan
# ---------------------------------------------------------------------------------------------------------------------

@dataclass
class RelationNode:
    """A file-based Parquet relation (DefaultFileBasedRelation, index/sources/default/DefaultFileBasedRelation.scala:38-242)."""
    root_paths: List[str]
    files: List[Tuple[str, int, int]]  # (uri, size, mtime) of every data file, DataPathFilter applied
    schema: List[Tuple[str, str]]      # (name, spark type name)

    @property
    def signature(self) -> str:
        """md5 fold over len + mtime + path of the files sorted by path (DefaultFileBasedRelation.scala:45-53,193-196)."""
        acc = ""
        for uri, size, mtime in sorted(self.files, key=lambda f: f[0]):
            acc = LE.md5_hex(acc + f"{size}{mtime}{uri}")
        return acc

    @property
    def column_names(self) -> List[str]:
        return [n for n, _ in self.schema]


@dataclass
class FilterNode:
    child: object
    predicate: Predicate


@dataclass
class ProjectNode:
    child: object
    columns: List[str]


@dataclass
class JoinNode:
    left: object
    right: object
    left_key: str
    right_key: str


_SPARK_TYPE_OF_ARROW = {"int32": "integer", "int64": "long", "float": "float", "double": "double", "bool": "boolean",
                        "string": "string", "large_string": "string", "date32[day]": "date", "timestamp[us]": "timestamp",
                        "int8": "byte", "int16": "short"}


def list_data_files(path: str) -> List[Tuple[str, int, int]]:
    p = LE.from_uri(path)
    out = []
    if os.path.isdir(p):
        for dirpath, dirnames, filenames in os.walk(p):
            dirnames[:] = sorted(d for d in dirnames if not d.startswith("_") and not d.startswith("."))
            for fn in sorted(filenames):
                if fn.startswith("_") or fn.startswith("."):
                    continue
                out.append(LE.file_status(os.path.join(dirpath, fn)))
    elif os.path.isfile(p):
        out.append(LE.file_status(p))
    else:
        raise LE.HyperspaceException(f"Path does not exist: {path}")
    return out


def read_parquet_schema(path: str) -> List[Tuple[str, str]]:
    """Footer-only read (driver-side metadata, like Spark's schema inference)."""
    import pyarrow.parquet as pq

    sch = pq.ParquetFile(LE.from_uri(path)).schema_arrow
    return [(f.name, _SPARK_TYPE_OF_ARROW.get(str(f.type), str(f.type))) for f in sch]


class DataFrameReader:
    def __init__(self, session: "HyperspaceSession"):
        self._s = session

    def parquet(self, *paths: str) -> "DataFrame":
        files: List[Tuple[str, int, int]] = []
        for p in paths:
            files.extend(list_data_files(p))
        if not files:
            raise LE.HyperspaceException(f"No Parquet data files under {paths}")
        schema = read_parquet_schema(files[0][0])
        return DataFrame(self._s, RelationNode([LE.to_uri(p) for p in paths], files, schema))


class DataFrame:
    def __init__(self, session: "HyperspaceSession", plan):
        self.session = session
        self.plan = plan

    # ---- transformations ------------------------------------------------------------------------------------
    def _resolve(self, name: str) -> str:
        """Column names resolve case-insensitively (Spark's default, spark.sql.caseSensitive=false; the reference does the
        same for index configs in util/ResolverUtils.scala) and come out in the schema's own spelling."""
        hits = [c for c in self.columns if c.lower() == name.lower()]
        if not hits:
            raise LE.HyperspaceException(f"cannot resolve column '{name}' among ({', '.join(self.columns)})")
        if len(hits) > 1 and name not in hits:
            raise LE.HyperspaceException(f"Reference '{name}' is ambiguous, could be: {', '.join(hits)}")
        return name if name in hits else hits[0]

    def filter(self, predicate: Predicate) -> "DataFrame":
        resolved = Predicate({self._resolve(c): b for c, b in predicate.bounds.items()})
        return DataFrame(self.session, FilterNode(self.plan, resolved))

    where = filter

    def select(self, *columns: str) -> "DataFrame":
        cols = list(columns[0]) if len(columns) == 1 and isinstance(columns[0], (list, tuple)) else list(columns)
        return DataFrame(self.session, ProjectNode(self.plan, [self._resolve(c) for c in cols]))

    def join(self, other: "DataFrame", on, how: str = "inner") -> "DataFrame":
        if how != "inner":
            raise LE.HyperspaceException("only inner equi-joins are handled by the GPU path")
        lk, rk = (on, on) if isinstance(on, str) else on
        return DataFrame(self.session, JoinNode(self.plan, other.plan, self._resolve(lk), other._resolve(rk)))

    # ---- introspection ------------------------------------------------------------------------------------
    @property
    def columns(self) -> List[str]:
        return output_columns(self.plan)

    def explain(self) -> str:
        from .rules import plan_query

        return plan_query(self.session, self.plan).describe()

    # ---- actions ------------------------------------------------------------------------------------
    def collect(self) -> Dict[str, np.ndarray]:
        """Executes the plan on the GPU and returns the result columns as numpy arrays (row order unspecified)."""
        from .rules import plan_query

        return plan_query(self.session, self.plan).execute()

    def count(self) -> int:
        res = self.collect()
        return len(next(iter(res.values()))) if res else 0


def output_columns(plan) -> List[str]:
    if isinstance(plan, RelationNode):
        return plan.column_names
    if isinstance(plan, FilterNode):
        return output_columns(plan.child)
    if isinstance(plan, ProjectNode):
        return list(plan.columns)
    if isinstance(plan, JoinNode):
        return output_columns(plan.left) + [c for c in output_columns(plan.right)]
    raise TypeError(plan)


class HyperspaceSession:
    """Stand-in for the SparkSession a Hyperspace object is constructed with (src/main/scala/.../Hyperspace.scala:27)."""

    def __init__(self, conf: Optional[Dict[str, str]] = None, device: int = 0):
        self.conf = RuntimeConf(conf)
        self.device = device
        self._ctx = None

    @property
    def read(self) -> DataFrameReader:
        return DataFrameReader(self)

    @property
    def gpu(self):
        """The native context (created on first use; raises without a CUDA device -- there is no CPU fallback)."""
        if self._ctx is None:
            from . import _native

            self._ctx = _native.Context(self.device)
        return self._ctx

    # S/package.scala:40-93
    def enableHyperspace(self) -> "HyperspaceSession":
        self.conf.set(HYPERSPACE_ENABLED, True)
        return self

    def disableHyperspace(self) -> "HyperspaceSession":
        self.conf.set(HYPERSPACE_ENABLED, False)
        return self

    def isHyperspaceEnabled(self) -> bool:
        return self.conf.get_bool(HYPERSPACE_ENABLED, False)

    def stop(self) -> None:
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None

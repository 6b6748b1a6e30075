"""On-disk metadata contract of a Hyperspace index: IndexLogEntry JSON, operation log, versioned data directories.

Restates (host-side, no data-parallel work):
  * ``IndexLogEntry`` / ``Content`` / ``Directory`` / ``FileInfo`` / ``Relation`` / ``Hdfs`` / ``Update`` / ``Source`` /
    ``SparkPlan`` / ``Signature`` / ``FileIdTracker``   -- src/main/scala/com/microsoft/hyperspace/index/IndexLogEntry.scala:34-703
  * ``IndexLogManagerImpl``                               -- index/IndexLogManager.scala:57-195
  * ``IndexDataManagerImpl``                              -- index/IndexDataManager.scala:50-108
  * ``PathResolver``                                      -- index/PathResolver.scala:30-70
  * ``JsonUtils`` pretty-printed Jackson output           -- util/JsonUtils.scala:35-50
The JSON field names, nesting, the ``type`` discriminator (``com.microsoft.hyperspace.index.covering.CoveringIndex``)
and the directory-tree encoding of file lists follow the spec example in
src/test/scala/com/microsoft/hyperspace/index/IndexLogEntryTest.scala:74-188 (golden copy in tests/golden/).
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import time
import uuid
from dataclasses import dataclass, field
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

HYPERSPACE_LOG = "_hyperspace_log"                       # index/IndexConstants.scala
INDEX_VERSION_DIRECTORY_PREFIX = "v__"                   # index/IndexConstants.scala
LATEST_STABLE_LOG_NAME = "latestStable"                  # index/IndexLogManager.scala:67
HYPERSPACE_VERSION_PROPERTY = "hyperspaceVersion"
HYPERSPACE_VERSION = "0.5.0-SNAPSHOT"
INDEX_LOG_VERSION = "indexLogVersion"
LINEAGE_PROPERTY = "lineage"
HAS_PARQUET_AS_SOURCE_FORMAT_PROPERTY = "hasParquetAsSourceFormat"
DATA_FILE_NAME_ID = "_data_file_id"
COVERING_INDEX_TYPE = "com.microsoft.hyperspace.index.covering.CoveringIndex"
UNKNOWN_FILE_ID = -1


class HyperspaceException(Exception):
    """src/main/scala/com/microsoft/hyperspace/HyperspaceException.scala"""


class States:  # actions/Constants.scala:19-35
    ACTIVE = "ACTIVE"
    CREATING = "CREATING"
    DELETING = "DELETING"
    DELETED = "DELETED"
    REFRESHING = "REFRESHING"
    VACUUMING = "VACUUMING"
    VACUUMINGOUTDATED = "VACUUMINGOUTDATED"
    RESTORING = "RESTORING"
    OPTIMIZING = "OPTIMIZING"
    DOESNOTEXIST = "DOESNOTEXIST"
    CANCELLING = "CANCELLING"


STABLE_STATES = {States.ACTIVE, States.DELETED, States.DOESNOTEXIST}


def md5_hex(s: str) -> str:
    """util/HashingUtils.scala:32-34"""
    return hashlib.md5(s.encode("utf-8")).hexdigest()


def to_uri(path: str) -> str:
    """Hadoop ``Path.toString`` of a local file: ``file:/abs/path``."""
    if path.startswith("file:"):
        return path
    return "file:" + os.path.abspath(path)


def from_uri(uri: str) -> str:
    return uri[5:] if uri.startswith("file:") else uri


# ---------------------------------------------------------------------------------------------------------------------
# file lists as directory trees
# ---------------------------------------------------------------------------------------------------------------------

@dataclass(frozen=True)
class FileInfo:
    """IndexLogEntry.scala:321-349.  Equality ignores ``id`` (name, size, modifiedTime identify a file version)."""
    name: str
    size: int
    modifiedTime: int
    id: int = field(compare=False, default=UNKNOWN_FILE_ID)

    def to_json(self):
        return {"name": self.name, "size": self.size, "modifiedTime": self.modifiedTime, "id": self.id}

    @staticmethod
    def from_json(j):
        return FileInfo(j["name"], j["size"], j["modifiedTime"], j.get("id", UNKNOWN_FILE_ID))


class FileIdTracker:
    """IndexLogEntry.scala:627-703: unique ids per (full path, size, modifiedTime), assigned from 0 in arrival order."""

    def __init__(self):
        self._max_id = -1
        self._map: Dict[Tuple[str, int, int], int] = {}

    @property
    def max_file_id(self) -> int:
        return self._max_id

    def get_file_id(self, path: str, size: int, mtime: int) -> Optional[int]:
        return self._map.get((path, size, mtime))

    def add_file_info(self, files: Iterable[FileInfo]) -> None:
        for f in files:
            if f.id == UNKNOWN_FILE_ID:
                raise HyperspaceException(f"Cannot add file info with unknown id. (file: {f.name}).")
            key = (f.name, f.size, f.modifiedTime)
            old = self._map.get(key)
            if old is not None and old != f.id:
                raise HyperspaceException(f"Adding file info with a conflicting id. (existing id: {old}, new id: {f.id}, file: {f.name}).")
            if old is None:
                self._map[key] = f.id
                self._max_id = max(self._max_id, f.id)

    def add_file(self, path: str, size: int, mtime: int) -> int:
        key = (path, size, mtime)
        if key not in self._map:
            self._max_id += 1
            self._map[key] = self._max_id
        return self._map[key]

    def id_to_file(self) -> Dict[int, str]:
        return {v: k[0] for k, v in self._map.items()}


def file_status(path: str) -> Tuple[str, int, int]:
    """(uri, length, modification time in ms) -- the fields of Hadoop's FileStatus Hyperspace records."""
    st = os.stat(from_uri(path))
    return to_uri(path), st.st_size, int(st.st_mtime * 1000)


@dataclass
class Directory:
    """IndexLogEntry.scala:136-319"""
    name: str
    files: List[FileInfo] = field(default_factory=list)
    subDirs: List["Directory"] = field(default_factory=list)

    def to_json(self):
        return {"name": self.name, "files": [f.to_json() for f in self.files], "subDirs": [d.to_json() for d in self.subDirs]}

    @staticmethod
    def from_json(j):
        return Directory(j["name"], [FileInfo.from_json(f) for f in j.get("files", [])],
                         [Directory.from_json(d) for d in j.get("subDirs", [])])

    def merge(self, that: "Directory") -> "Directory":
        """IndexLogEntry.scala:149-171"""
        if self.name != that.name:
            raise HyperspaceException(f"Merging directories with names {self.name} and {that.name} failed. "
                                      "Directory names must be same for merging directories.")
        mine = {d.name: d for d in self.subDirs}
        theirs = {d.name: d for d in that.subDirs}
        merged = []
        for n in list(dict.fromkeys(list(mine) + list(theirs))):
            if n in mine and n in theirs:
                merged.append(mine[n].merge(theirs[n]))
            else:
                merged.append(mine.get(n) or theirs[n])
        return Directory(self.name, self.files + that.files, merged)

    @staticmethod
    def _split(uri: str) -> List[str]:
        """['file:/', 'a', 'b', 'f.parquet'] for file:/a/b/f.parquet (the root directory is named by its URI)."""
        p = from_uri(uri)
        parts = [x for x in p.split("/") if x]
        return ["file:/"] + parts

    @staticmethod
    def from_leaf_files(files: Sequence[Tuple[str, int, int]], tracker: FileIdTracker) -> "Directory":
        """IndexLogEntry.scala:232-293: a tree rooted at the file-system root containing exactly the given leaf files."""
        if not files:
            raise HyperspaceException("Empty files list found while creating a Directory.")
        root = Directory("file:/")
        for uri, size, mtime in files:
            parts = Directory._split(uri)
            cur = root
            for name in parts[1:-1]:
                nxt = next((d for d in cur.subDirs if d.name == name), None)
                if nxt is None:
                    nxt = Directory(name)
                    cur.subDirs.append(nxt)
                cur = nxt
            cur.files.append(FileInfo(parts[-1], size, mtime, tracker.add_file(to_uri(uri), size, mtime)))
        return root

    @staticmethod
    def create_empty(path: str) -> "Directory":
        """IndexLogEntry.scala:205-214"""
        parts = Directory._split(to_uri(path))
        cur = None
        for name in reversed(parts[1:]):
            cur = Directory(name, subDirs=[cur] if cur else [])
        return Directory("file:/", subDirs=[cur] if cur else [])

    @staticmethod
    def from_directory(path: str, tracker: FileIdTracker) -> "Directory":
        """IndexLogEntry.scala:186-203 with PathUtils.DataPathFilter (util/PathUtils.scala:34-39)."""
        leaves = []
        p = from_uri(path)
        if os.path.isdir(p):
            for dirpath, dirnames, filenames in os.walk(p):
                dirnames.sort()
                for fn in sorted(filenames):
                    if fn.startswith("_") or fn.startswith("."):
                        continue
                    leaves.append(file_status(os.path.join(dirpath, fn)))
        return Directory.from_leaf_files(leaves, tracker) if leaves else Directory.create_empty(path)


@dataclass
class Content:
    """IndexLogEntry.scala:40-113"""
    root: Directory

    def to_json(self):
        return {"root": self.root.to_json(), "fingerprint": {"kind": "NoOp", "properties": {}}}

    @staticmethod
    def from_json(j):
        return Content(Directory.from_json(j["root"])) if j is not None else None

    def _rec(self, prefix: str, d: Directory, out: List[FileInfo]):
        for f in d.files:
            out.append(FileInfo(_join(prefix, f.name), f.size, f.modifiedTime, f.id))
        for s in d.subDirs:
            self._rec(_join(prefix, s.name), s, out)

    @property
    def file_infos(self) -> List[FileInfo]:
        """Fully qualified FileInfo of every leaf file."""
        out: List[FileInfo] = []
        self._rec(self.root.name, self.root, out)
        return out

    @property
    def files(self) -> List[str]:
        return [f.name for f in self.file_infos]

    @staticmethod
    def from_directory(path: str, tracker: FileIdTracker) -> "Content":
        return Content(Directory.from_directory(path, tracker))

    @staticmethod
    def from_leaf_files(files: Sequence[Tuple[str, int, int]], tracker: FileIdTracker) -> Optional["Content"]:
        return Content(Directory.from_leaf_files(files, tracker)) if files else None


def _join(prefix: str, name: str) -> str:
    if prefix.endswith("/"):
        return prefix + name
    return prefix + "/" + name


# ---------------------------------------------------------------------------------------------------------------------
# source relation
# ---------------------------------------------------------------------------------------------------------------------

@dataclass
class Update:
    """IndexLogEntry.scala:377-379: files appended to / deleted from the source since the index data was built."""
    appendedFiles: Optional[Content] = None
    deletedFiles: Optional[Content] = None

    def to_json(self):
        return {"deletedFiles": self.deletedFiles.to_json() if self.deletedFiles else None,
                "appendedFiles": self.appendedFiles.to_json() if self.appendedFiles else None}

    @staticmethod
    def from_json(j):
        if j is None:
            return None
        return Update(Content.from_json(j.get("appendedFiles")), Content.from_json(j.get("deletedFiles")))


@dataclass
class Relation:
    """IndexLogEntry.scala:395-406"""
    rootPaths: List[str]
    content: Content
    dataSchema: dict
    fileFormat: str
    options: Dict[str, str] = field(default_factory=dict)
    update: Optional[Update] = None

    def to_json(self):
        return {"rootPaths": self.rootPaths,
                "data": {"properties": {"content": self.content.to_json(),
                                        "update": self.update.to_json() if self.update else None},
                         "kind": "HDFS"},
                "dataSchema": self.dataSchema, "fileFormat": self.fileFormat, "options": self.options}

    @staticmethod
    def from_json(j):
        p = j["data"]["properties"]
        return Relation(j["rootPaths"], Content.from_json(p["content"]), j["dataSchema"], j["fileFormat"], j.get("options", {}),
                        Update.from_json(p.get("update")))


@dataclass
class Signature:
    provider: str
    value: str


INDEX_SIGNATURE_PROVIDER = "com.microsoft.hyperspace.index.IndexSignatureProvider"


@dataclass
class IndexLogEntry:
    """IndexLogEntry.scala:408-622"""
    name: str
    indexedColumns: List[str]
    includedColumns: List[str]
    schema: dict                       # Spark StructType JSON of the index data
    numBuckets: int
    derived_properties: Dict[str, str]
    content: Content
    relations: List[Relation]
    signatures: List[Signature]
    properties: Dict[str, str] = field(default_factory=dict)
    id: int = 0
    state: str = States.DOESNOTEXIST
    timestamp: int = 0
    enabled: bool = True
    version: str = "0.1"

    # ---- derived views ------------------------------------------------------------------------------------
    @property
    def source_file_infos(self) -> List[FileInfo]:
        return self.relations[0].content.file_infos

    @property
    def source_files_size_in_bytes(self) -> int:
        return sum(f.size for f in self.source_file_infos)

    @property
    def index_files(self) -> List[str]:
        return self.content.files

    @property
    def index_files_size_in_bytes(self) -> int:
        return sum(f.size for f in self.content.file_infos)

    @property
    def has_lineage_column(self) -> bool:
        return self.derived_properties.get(LINEAGE_PROPERTY, "false").lower() == "true"

    @property
    def appended_files(self) -> List[FileInfo]:
        u = self.relations[0].update
        return u.appendedFiles.file_infos if u and u.appendedFiles else []

    @property
    def deleted_files(self) -> List[FileInfo]:
        u = self.relations[0].update
        return u.deletedFiles.file_infos if u and u.deletedFiles else []

    def file_id_tracker(self) -> FileIdTracker:
        t = FileIdTracker()
        t.add_file_info(self.source_file_infos)
        return t

    def index_version_dirs(self) -> List[int]:
        out = set()
        for f in self.index_files:
            for part in from_uri(f).split("/"):
                if part.startswith(INDEX_VERSION_DIRECTORY_PREFIX + "="):
                    out.add(int(part.split("=", 1)[1]))
        return sorted(out)

    # ---- JSON ------------------------------------------------------------------------------------
    def to_json_obj(self):
        return {
            "name": self.name,
            "derivedDataset": {"type": COVERING_INDEX_TYPE, "indexedColumns": self.indexedColumns,
                               "includedColumns": self.includedColumns, "schema": self.schema, "numBuckets": self.numBuckets,
                               "properties": self.derived_properties},
            "content": self.content.to_json(),
            "source": {"plan": {"properties": {"relations": [r.to_json() for r in self.relations], "rawPlan": None, "sql": None,
                                               "fingerprint": {"properties": {"signatures": [{"provider": s.provider, "value": s.value}
                                                                                               for s in self.signatures]},
                                                               "kind": "LogicalPlan"}},
                                "kind": "Spark"}},
            "properties": self.properties,
            "version": self.version, "id": self.id, "state": self.state, "timestamp": self.timestamp, "enabled": self.enabled,
        }

    def to_json(self) -> str:
        """Pretty-printed like Jackson's writerWithDefaultPrettyPrinter (util/JsonUtils.scala:48-50)."""
        return json.dumps(self.to_json_obj(), indent=2, separators=(",", " : "))

    @staticmethod
    def from_json(text: str) -> "IndexLogEntry":
        j = json.loads(text)
        dd = j["derivedDataset"]
        if dd.get("type") != COVERING_INDEX_TYPE:
            raise HyperspaceException(f"Unsupported index type {dd.get('type')}: only covering indexes are on the GPU path")
        plan = j["source"]["plan"]["properties"]
        return IndexLogEntry(
            name=j["name"], indexedColumns=list(dd["indexedColumns"]), includedColumns=list(dd["includedColumns"]),
            schema=dd["schema"], numBuckets=int(dd["numBuckets"]), derived_properties=dict(dd.get("properties", {})),
            content=Content.from_json(j["content"]), relations=[Relation.from_json(r) for r in plan["relations"]],
            signatures=[Signature(s["provider"], s["value"]) for s in plan["fingerprint"]["properties"]["signatures"]],
            properties=dict(j.get("properties", {})), id=int(j.get("id", 0)), state=j.get("state", States.DOESNOTEXIST),
            timestamp=int(j.get("timestamp", 0)), enabled=bool(j.get("enabled", True)), version=j.get("version", "0.1"))

    def copy(self, **changes) -> "IndexLogEntry":
        import dataclasses

        return dataclasses.replace(self, **changes)


# ---------------------------------------------------------------------------------------------------------------------
# operation log + data directories
# ---------------------------------------------------------------------------------------------------------------------

class IndexLogManager:
    """index/IndexLogManager.scala:57-195: optimistic concurrency through create-temp-then-rename of ``<id>`` files."""

    def __init__(self, index_path: str):
        self.index_path = index_path
        self.log_path = os.path.join(index_path, HYPERSPACE_LOG)

    def _path(self, id_: int) -> str:
        return os.path.join(self.log_path, str(id_))

    def _read(self, path: str) -> Optional[IndexLogEntry]:
        if not os.path.exists(path):
            return None
        try:
            return IndexLogEntry.from_json(open(path).read())
        except HyperspaceException:
            raise
        except Exception as e:
            raise HyperspaceException(f"Cannot parse JSON in {path}: {e}")

    def get_log(self, id_: int) -> Optional[IndexLogEntry]:
        return self._read(self._path(id_))

    def get_latest_id(self) -> Optional[int]:
        if not os.path.isdir(self.log_path):
            return None
        ids = [int(n) for n in os.listdir(self.log_path) if n.isdigit()]
        return max(ids) if ids else None

    def get_latest_log(self) -> Optional[IndexLogEntry]:
        i = self.get_latest_id()
        return self.get_log(i) if i is not None else None

    def get_latest_stable_log(self) -> Optional[IndexLogEntry]:
        log = self._read(os.path.join(self.log_path, LATEST_STABLE_LOG_NAME))
        if log is not None:
            assert log.state in STABLE_STATES
            return log
        latest = self.get_latest_id()
        if latest is None:
            return None
        for i in range(latest, -1, -1):
            e = self.get_log(i)
            if e and e.state in STABLE_STATES:
                return e
            if e and e.state in (States.CREATING, States.VACUUMING):
                return None
        return None

    def get_index_versions(self, states: Sequence[str]) -> List[int]:
        latest = self.get_latest_id()
        if latest is None:
            return []
        out = []
        for i in range(latest, -1, -1):
            e = self.get_log(i)
            if e and e.state in states:
                out.append(i)
        return out

    def create_latest_stable_log(self, id_: int) -> bool:
        e = self.get_log(id_)
        if e is None or e.state not in STABLE_STATES:
            return False
        try:
            shutil.copyfile(self._path(id_), os.path.join(self.log_path, LATEST_STABLE_LOG_NAME))
            return True
        except OSError:
            return False

    def delete_latest_stable_log(self) -> bool:
        p = os.path.join(self.log_path, LATEST_STABLE_LOG_NAME)
        try:
            if os.path.exists(p):
                os.remove(p)
            return True
        except OSError:
            return False

    def write_log(self, id_: int, entry: IndexLogEntry) -> bool:
        """False when someone else already wrote ``id_`` (the caller then fails with 'Could not acquire proper state')."""
        if os.path.exists(self._path(id_)):
            return False
        os.makedirs(self.log_path, exist_ok=True)
        tmp = os.path.join(self.log_path, "temp" + str(uuid.uuid4()))
        with open(tmp, "w") as f:
            f.write(entry.to_json())
        try:
            os.link(tmp, self._path(id_))  # atomic, fails if the target exists (rename would overwrite on POSIX)
            os.remove(tmp)
            return True
        except OSError:
            os.remove(tmp)
            return False


class IndexDataManager:
    """index/IndexDataManager.scala:50-108: index data lives in ``<indexPath>/v__=<N>``."""

    def __init__(self, index_path: str):
        self.index_path = index_path

    def get_all_version_ids(self) -> List[int]:
        if not os.path.isdir(self.index_path):
            return []
        out = []
        for n in os.listdir(self.index_path):
            if n.startswith(INDEX_VERSION_DIRECTORY_PREFIX + "="):
                try:
                    out.append(int(n.split("=", 1)[1]))
                except ValueError:
                    pass
        return sorted(out)

    def get_latest_version_id(self) -> Optional[int]:
        ids = self.get_all_version_ids()
        return ids[-1] if ids else None

    def get_path(self, id_: int) -> str:
        return os.path.join(self.index_path, f"{INDEX_VERSION_DIRECTORY_PREFIX}={id_}")

    def delete(self, id_: int) -> None:
        shutil.rmtree(self.get_path(id_), ignore_errors=True)


class PathResolver:
    """index/PathResolver.scala:30-70: ``spark.hyperspace.system.path`` + case-insensitive index directory lookup."""

    def __init__(self, conf):
        self.conf = conf

    @property
    def system_path(self) -> str:
        p = self.conf.get("spark.hyperspace.system.path")
        if not p:
            p = os.path.join(self.conf.get("spark.sql.warehouse.dir", os.path.abspath("spark-warehouse")), "indexes")
        return from_uri(p)

    def get_index_path(self, name: str) -> str:
        root = self.system_path
        if os.path.isdir(root):
            matches = [d for d in os.listdir(root) if d.lower() == name.lower()]
            if len(matches) > 1:
                raise HyperspaceException(f"There are multiple directories with the index name '{name}' in {root}")
            if matches:
                return os.path.join(root, matches[0])
        return os.path.join(root, name)


def now_ms() -> int:
    return int(time.time() * 1000)

"""The reference's data-free unit suites, restated case by case against the Python host layer (CPU only).

Every test names the reference test it follows (``T/`` = ``src/test/scala/com/microsoft/hyperspace/``).  Where the reference
mocks ``IndexLogManager`` / ``IndexDataManager`` with Mockito, small recording fakes play the same role here.
"""
import os
import uuid

import pytest

from hyperspace_b200 import log_entry as LE
from hyperspace_b200.hyperspace import CancelAction, VacuumAction, _Action, _StateFlip
from hyperspace_b200.index_config import CoveringIndexConfig, IndexConfig
from hyperspace_b200.log_entry import FileIdTracker, FileInfo, HyperspaceException, States

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _entry(state: str) -> LE.IndexLogEntry:
    e = LE.IndexLogEntry.from_json(open(os.path.join(GOLDEN, "index_log_entry_spec.json")).read())
    e.state = state
    return e


def _put(index_path, id_, content: str):
    d = os.path.join(index_path, LE.HYPERSPACE_LOG)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, str(id_)), "w") as f:
        f.write(content)


# ---------------------------------------------------------------------------------------------------------------------
# T/index/IndexLogManagerImplTest.scala
# ---------------------------------------------------------------------------------------------------------------------

def test_get_log_returns_none_if_log_not_found(tmp_path):
    assert LE.IndexLogManager(str(tmp_path / "testPath")).get_log(0) is None


def test_get_log_returns_entry_if_id_found(tmp_path):
    path = str(tmp_path / "testPath")
    _put(path, 0, _entry("ACTIVE").to_json())
    assert LE.IndexLogManager(path).get_log(0).to_json() == _entry("ACTIVE").to_json()


def test_get_log_fails_if_json_is_not_in_proper_form(tmp_path):
    path = str(tmp_path / "testPath")
    js = _entry("ACTIVE").to_json()
    i = js.index('"source"') + 8
    _put(path, 0, js[:i] + "\x00" + js[i:])
    with pytest.raises(HyperspaceException):
        LE.IndexLogManager(path).get_log(0)


def test_write_log_passes_only_if_no_other_file_exists_with_same_name(tmp_path):
    path = str(tmp_path / str(uuid.uuid4()))
    assert LE.IndexLogManager(path).write_log(0, _entry("ACTIVE"))
    assert not LE.IndexLogManager(path).write_log(0, _entry("ACTIVE"))


def test_get_latest_id_ignores_non_numeric_names(tmp_path):
    path = str(tmp_path / str(uuid.uuid4()))
    for name in ("0", "1", "abc", "20"):
        _put(path, name, "file contents")
    assert LE.IndexLogManager(path).get_latest_id() == 20


def test_get_latest_stable_log_returns_latest_stable_log(tmp_path):
    path = str(tmp_path / str(uuid.uuid4()))
    for id_, state in ((0, "CREATING"), (1, "ACTIVE"), (3, "ACTIVE"), (4, "REFRESHING"), (20, "CANCELLING")):
        _put(path, id_, _entry(state).to_json())
    lm = LE.IndexLogManager(path)
    assert lm.get_latest_stable_log().to_json() == _entry("ACTIVE").to_json()
    assert lm.get_index_versions(["ACTIVE"]) == [3, 1]


def test_get_latest_stable_log_does_not_return_irrelevant_previous_log(tmp_path):
    path = str(tmp_path / str(uuid.uuid4()))
    _put(path, 8, _entry("ACTIVE").to_json())
    _put(path, 10, _entry("VACUUMING").to_json())
    assert LE.IndexLogManager(path).get_latest_stable_log() is None  # VACUUMING cuts the history off
    _put(path, 12, _entry("CREATING").to_json())
    assert LE.IndexLogManager(path).get_latest_stable_log() is None


def test_create_latest_stable_log(tmp_path):
    path = str(tmp_path / str(uuid.uuid4()))
    _put(path, 0, _entry("ACTIVE").to_json())
    assert LE.IndexLogManager(path).create_latest_stable_log(0) is True
    assert os.path.exists(os.path.join(path, LE.HYPERSPACE_LOG, "latestStable"))
    # ... fails if the log state is not stable
    path2 = str(tmp_path / str(uuid.uuid4()))
    _put(path2, 0, _entry("CANCELLING").to_json())
    assert LE.IndexLogManager(path2).create_latest_stable_log(0) is False
    assert not os.path.exists(os.path.join(path2, LE.HYPERSPACE_LOG, "latestStable"))
    # ... fails with an exception if it cannot find a valid log entry
    path3 = str(tmp_path / str(uuid.uuid4()))
    _put(path3, 0, "Invalid Log Entry")
    with pytest.raises(HyperspaceException):
        LE.IndexLogManager(path3).create_latest_stable_log(0)


# ---------------------------------------------------------------------------------------------------------------------
# T/actions/ActionTest.scala, DeleteActionTest, RestoreActionTest, VacuumActionTest, CancelActionTest
# ---------------------------------------------------------------------------------------------------------------------

class _RecordingLogManager:
    """The Mockito mock of the reference's action tests: canned answers, calls recorded."""

    def __init__(self, latest_id=None, log=None, stable=None):
        self.latest_id, self.log, self.stable = latest_id, log, stable
        self.calls = []

    def get_latest_id(self):
        return self.latest_id

    def get_log(self, id_):
        return self.log

    def get_latest_stable_log(self):
        return self.stable

    def write_log(self, id_, entry):
        self.calls.append(("writeLog", id_, entry.state))
        return True

    def delete_latest_stable_log(self):
        self.calls.append(("deleteLatestStableLog",))
        return True

    def create_latest_stable_log(self, id_):
        self.calls.append(("createLatestStableLog", id_))
        return True


def test_action_run_protocol():
    """ActionTest 'verify run()': writeLog(0, CREATING), deleteLatestStableLog, writeLog(1, ACTIVE), createLatestStableLog(1)."""
    lm = _RecordingLogManager()

    class A(_Action):
        transient_state, final_state = States.CREATING, States.ACTIVE

        def log_entry(self):
            return _entry(States.DOESNOTEXIST)

    A(lm).run()
    assert lm.calls == [("writeLog", 0, "CREATING"), ("deleteLatestStableLog",), ("writeLog", 1, "ACTIVE"),
                        ("createLatestStableLog", 1)]


def test_action_fails_when_the_log_slot_is_taken():
    """Action.scala:66-70: losing the optimistic write means 'Could not acquire proper state'."""
    lm = _RecordingLogManager()
    lm.write_log = lambda id_, entry: False

    class A(_Action):
        transient_state, final_state = States.CREATING, States.ACTIVE

        def log_entry(self):
            return _entry(States.DOESNOTEXIST)

    with pytest.raises(HyperspaceException, match="Could not acquire proper state"):
        A(lm).run()


def test_delete_action_validate():
    """DeleteActionTest: passes from ACTIVE, fails otherwise with 'Delete is only supported in ACTIVE state'."""
    _StateFlip(_RecordingLogManager(log=_entry("ACTIVE")), "ACTIVE", "DELETING", "DELETED", "Delete").validate()
    with pytest.raises(HyperspaceException, match="Delete is only supported in ACTIVE state"):
        _StateFlip(_RecordingLogManager(log=_entry("CREATING")), "ACTIVE", "DELETING", "DELETED", "Delete").validate()


def test_restore_action_validate():
    """RestoreActionTest: passes from DELETED, fails otherwise with 'Restore is only supported in DELETED state'."""
    _StateFlip(_RecordingLogManager(log=_entry("DELETED")), "DELETED", "RESTORING", "ACTIVE", "Restore").validate()
    with pytest.raises(HyperspaceException, match="Restore is only supported in DELETED state"):
        _StateFlip(_RecordingLogManager(log=_entry("ACTIVE")), "DELETED", "RESTORING", "ACTIVE", "Restore").validate()


class _RecordingDataManager:
    def __init__(self, versions):
        self.versions, self.deleted = list(versions), []

    def get_all_version_ids(self):
        return list(self.versions)

    def get_latest_version_id(self):
        return max(self.versions) if self.versions else None

    def delete(self, id_):
        self.deleted.append(id_)


def test_vacuum_action_validate_and_op():
    """VacuumActionTest: validate() passes only from DELETED; op() deletes every data version (0, 1, 2 and nothing else)."""
    dm = _RecordingDataManager([0, 1, 2])
    VacuumAction(_RecordingLogManager(log=_entry("DELETED")), dm).validate()
    with pytest.raises(HyperspaceException):
        VacuumAction(_RecordingLogManager(log=_entry("CREATING")), dm).validate()
    VacuumAction(_RecordingLogManager(log=_entry("DELETED")), dm).op()
    assert sorted(dm.deleted) == [0, 1, 2]


@pytest.mark.parametrize("current,stable,final", [
    ("ACTIVE", "ACTIVE", "ACTIVE"),            # 'Cancel leads to ACTIVE from ACTIVE state'
    ("REFRESHING", "ACTIVE", "ACTIVE"),        # '... to last stable state from transient state if stable state exists'
    ("VACUUMING", None, "DOESNOTEXIST"),       # '... to DoesNotExist state from VACUUMING'
    ("REFRESHING", None, "DOESNOTEXIST"),      # '... to DoesNotExist from transient state if no stable state exists'
])
def test_cancel_action_final_state(current, stable, final):
    lm = _RecordingLogManager(log=_entry(current), stable=_entry(stable) if stable else None)
    assert CancelAction(lm).final_state == final


def test_cancel_action_validate_rejects_stable_states():
    """CancelAction.scala:44-52."""
    with pytest.raises(HyperspaceException, match="Cancel\\(\\) is not supported in stable states"):
        CancelAction(_RecordingLogManager(log=_entry("ACTIVE"), stable=_entry("ACTIVE"))).validate()
    CancelAction(_RecordingLogManager(log=_entry("REFRESHING"), stable=_entry("ACTIVE"))).validate()


# ---------------------------------------------------------------------------------------------------------------------
# T/index/IndexConfigTest.scala
# ---------------------------------------------------------------------------------------------------------------------

def test_index_config_empty_names_and_columns_are_not_allowed():
    with pytest.raises(ValueError):
        IndexConfig("", ["c1"], ["c2"])
    with pytest.raises(ValueError):
        IndexConfig.builder().indexName("")
    with pytest.raises(ValueError):
        IndexConfig("name", [], ["c1"])
    with pytest.raises(ValueError):
        IndexConfig.builder().indexName("name").include("c1").create()


def test_index_config_same_column_names_case_insensitive_are_not_allowed():
    with pytest.raises(ValueError):
        IndexConfig("name", ["c1", "C1"], ["c2"])
    with pytest.raises(ValueError):
        IndexConfig.builder().indexName("name").indexBy("c1", "C1").include("c2").create()
    with pytest.raises(ValueError):
        IndexConfig("name", ["c1"], ["C1", "c2"])
    with pytest.raises(ValueError):
        IndexConfig.builder().indexName("name").indexBy("c1").include("C1", "c2").create()


def test_index_config_equals_and_hash():
    base = IndexConfig("name", ["c1", "c2"], ["c3", "c4"])
    assert base != object()
    assert base != IndexConfig("name", ["c2", "c1"], ["c3", "c4"])       # indexed column order matters
    assert base != IndexConfig("name", ["c1", "c5"], ["c3", "c4"])
    assert base != IndexConfig("name", ["c1", "c2"], ["c3", "c5"])
    assert IndexConfig("Name1", ["c1", "c2"], ["c3", "c4"]) != IndexConfig("Name2", ["c1", "c2"], ["c3", "c4"])
    assert base == IndexConfig("name", ["c1", "c2"], ["c3", "c4"])
    assert base == IndexConfig("name", ["c1", "c2"], ["c4", "c3"])       # included columns are a set
    assert base == IndexConfig("Name", ["C1", "C2"], ["C3", "C4"])       # everything is case-insensitive
    a, b = IndexConfig("name1", ["c1"], ["c2"]), IndexConfig("name1", ["C1"], ["c2"])
    assert a == b and hash(a) == hash(b)
    c, d = IndexConfig("name3", ["c1"], ["c3", "c4"]), IndexConfig("name3", ["C1"], ["c4", "c3"])
    assert c == d and hash(c) == hash(d)


def test_index_config_builder():
    cfg = IndexConfig.builder().indexName("Name").indexBy("C1", "c2", "C3").include("C4", "c5", "C6").create()
    assert isinstance(cfg, CoveringIndexConfig)
    assert cfg.indexName == "Name" and cfg.indexedColumns == ["C1", "c2", "C3"] and cfg.includedColumns == ["C4", "c5", "C6"]
    # 'Test exception on multiple indexBy, include and index name on IndexConfig builder.'
    with pytest.raises(NotImplementedError):
        IndexConfig.builder().indexName("name1").indexName("name2")
    with pytest.raises(NotImplementedError):
        IndexConfig.builder().indexName("name").indexBy("c1").indexBy("c2")
    with pytest.raises(NotImplementedError):
        IndexConfig.builder().indexName("name").indexBy("c1").include("c3").include("c4")


# ---------------------------------------------------------------------------------------------------------------------
# T/index/FileIdTrackerTest.scala
# ---------------------------------------------------------------------------------------------------------------------

def test_file_id_tracker_new_instance():
    t = FileIdTracker()
    assert t.max_file_id == -1 and t.id_to_file() == {}
    assert t.get_file_id("abc", 123, 555) is None
    t.add_file_info([])
    assert t.max_file_id == -1 and t.id_to_file() == {}


def test_file_id_tracker_add_file_info():
    t = FileIdTracker()
    with pytest.raises(HyperspaceException, match="Cannot add file info with unknown id"):
        t.add_file_info([FileInfo("abc", 123, 555, LE.UNKNOWN_FILE_ID)])
    # a conflict raises, but what was added before the conflict stays
    t = FileIdTracker()
    t.add_file_info([FileInfo("def", 123, 555, 10)])
    with pytest.raises(HyperspaceException, match="Adding file info with a conflicting id"):
        t.add_file_info(sorted([FileInfo("abc", 100, 555, 15), FileInfo("def", 123, 555, 11)], key=lambda f: f.name))
    assert t.get_file_id("abc", 100, 555) == 15
    # success: records added, max id raised
    t = FileIdTracker()
    t.add_file_info([FileInfo("abc", 123, 555, 10), FileInfo("def", 234, 777, 5)])
    assert t.get_file_id("abc", 123, 555) == 10 and t.get_file_id("def", 234, 777) == 5 and t.max_file_id == 10


def test_file_id_tracker_add_file():
    t = FileIdTracker()
    t.add_file_info([FileInfo("abc", 123, 555, 10)])
    assert t.add_file("abc", 123, 555) == 10 and t.max_file_id == 10  # existing id, max unchanged
    t = FileIdTracker()
    assert t.add_file("abc", 123, 555) == 0
    assert t.add_file("def", 123, 555) == 1
    assert t.add_file("xyz", 124, 777) == 2
    assert t.max_file_id == 2
    assert (t.get_file_id("abc", 123, 555), t.get_file_id("def", 123, 555), t.get_file_id("xyz", 124, 777)) == (0, 1, 2)


# ---------------------------------------------------------------------------------------------------------------------
# T/index/IndexCollectionManagerTest.scala (the manager's role is played by the Hyperspace facade + PathResolver)
# ---------------------------------------------------------------------------------------------------------------------

def _active_index(system_path, name):
    e = _entry("ACTIVE")
    e.name = name
    lm = LE.IndexLogManager(os.path.join(system_path, name))
    assert lm.write_log(0, e) and lm.create_latest_stable_log(0)


def test_get_indexes_returns_all_indexes(tmp_path):
    """'getIndexes() returns seq of Indexes': every index directory under the system path is listed with its entry."""
    from hyperspace_b200.hyperspace import Hyperspace
    from hyperspace_b200.session import HyperspaceSession

    system_path = str(tmp_path / "indexes")
    for n in ("idx1", "idx2", "idx3"):
        _active_index(system_path, n)
    hs = Hyperspace(HyperspaceSession({"spark.hyperspace.system.path": system_path}))
    got = hs.indexes()
    assert [i["name"] for i in got] == ["idx1", "idx2", "idx3"]
    assert all(i["state"] == "ACTIVE" and i["indexedColumns"] == ["col1"] and i["numBuckets"] == 200 for i in got)


@pytest.mark.parametrize("call", [
    lambda hs: hs.deleteIndex("idx4"),                      # 'delete() throws exception if index is not found'
    lambda hs: hs.vacuumIndex("idx4"),                      # 'vacuum() ...'
    lambda hs: hs.restoreIndex("idx4"),                     # 'restore() ...'
    lambda hs: hs.refreshIndex("idx4", "full"),             # "refresh() with mode = 'full' ..."
    lambda hs: hs.refreshIndex("idx4", "incremental"),      # "refresh() with mode = 'incremental' ..."
    lambda hs: hs.optimizeIndex("idx4"),
    lambda hs: hs.cancel("idx4"),
    lambda hs: hs.index("idx4"),
])
def test_operations_on_a_missing_index_raise(tmp_path, call):
    from hyperspace_b200.hyperspace import Hyperspace
    from hyperspace_b200.session import HyperspaceSession

    system_path = str(tmp_path / "indexes")
    _active_index(system_path, "idx1")
    with pytest.raises(HyperspaceException, match="could not be found"):
        call(Hyperspace(HyperspaceSession({"spark.hyperspace.system.path": system_path})))


# ---------------------------------------------------------------------------------------------------------------------
# T/index/IndexLogEntryTest.scala -- Content / Directory cases
# ---------------------------------------------------------------------------------------------------------------------

def _tree(d: LE.Directory):
    """Order-insensitive shape of a Directory (the reference's directoryEquals compares files and subDirs as sets)."""
    return (d.name, frozenset((f.name, f.size, f.modifiedTime, f.id) for f in d.files), frozenset(_tree(s) for s in d.subDirs))


def _under_root(path: str, leaf: LE.Directory) -> LE.Directory:
    """createDirectory(path, leaf) of the reference test: wrap `leaf` (the directory at `path`) up to the file-system root."""
    parts = [p for p in os.path.dirname(os.path.abspath(path)).split("/") if p]
    cur = leaf
    for name in reversed(parts):
        cur = LE.Directory(name, subDirs=[cur])
    return LE.Directory("file:/", subDirs=[cur])


def _touch(p, text="x"):
    os.makedirs(os.path.dirname(p), exist_ok=True)
    with open(p, "w") as f:
        f.write(text)
    return LE.file_status(p)


def test_content_files_lists_all_files():
    u = LE.UNKNOWN_FILE_ID
    content = LE.Content(LE.Directory("file:/", subDirs=[LE.Directory(
        "a", files=[FileInfo("f1", 0, 0, u), FileInfo("f2", 0, 0, u)],
        subDirs=[LE.Directory("b", files=[FileInfo("f3", 0, 0, u), FileInfo("f4", 0, 0, u)])])]))
    assert set(content.files) == {"file:/a/f1", "file:/a/f2", "file:/a/b/f3", "file:/a/b/f4"}


def test_directory_from_leaf_files_builds_exactly_the_given_tree(tmp_path):
    test_dir = str(tmp_path / "testDir")
    f1, f2 = _touch(f"{test_dir}/f1"), _touch(f"{test_dir}/f2")
    f3, f4 = _touch(f"{test_dir}/nested/f3"), _touch(f"{test_dir}/nested/f4")
    t = FileIdTracker()

    def info(st):
        return FileInfo(os.path.basename(st[0]), st[1], st[2], t.add_file(*st))

    expected = _under_root(test_dir, LE.Directory("testDir", [info(f1), info(f2)], [LE.Directory("nested", [info(f3), info(f4)])]))
    assert _tree(LE.Directory.from_leaf_files([f1, f2, f3, f4], t)) == _tree(expected)
    assert _tree(LE.Directory.from_directory(test_dir, t)) == _tree(expected)
    # 'fromLeafFiles api does not include other files in the directory.'
    expected = _under_root(test_dir, LE.Directory("testDir", [info(f1)], [LE.Directory("nested", [info(f4)])]))
    assert _tree(LE.Directory.from_leaf_files([f1, f4], t)) == _tree(expected)
    # 'Content.fromDirectory api creates the correct Content object.'
    expected = _under_root(f"{test_dir}/nested", LE.Directory("nested", [info(f3), info(f4)]))
    assert _tree(LE.Content.from_directory(f"{test_dir}/nested", t).root) == _tree(expected)
    assert _tree(LE.Content.from_leaf_files([f3, f4], t).root) == _tree(expected)


def test_directory_from_directory_empty_or_nonexistent(tmp_path):
    empty = str(tmp_path / "testDir" / "empty")
    expected = _under_root(empty, LE.Directory("empty"))
    t = FileIdTracker()
    assert _tree(LE.Directory.from_directory(empty, t)) == _tree(expected)   # nonexistent
    os.makedirs(empty)
    assert _tree(LE.Directory.from_directory(empty, t)) == _tree(expected)   # empty


def test_directory_with_a_gap_and_with_multiple_subdirectories(tmp_path):
    t = FileIdTracker()

    def info(st):
        return FileInfo(os.path.basename(st[0]), st[1], st[2], t.add_file(*st))

    # testDir/temp/a/f1, testDir/temp/b/c/f2
    temp = str(tmp_path / "testDir" / "temp")
    f1, f2 = _touch(f"{temp}/a/f1"), _touch(f"{temp}/b/c/f2")
    expected = _under_root(temp, LE.Directory("temp", subDirs=[
        LE.Directory("a", [info(f1)]), LE.Directory("b", subDirs=[LE.Directory("c", [info(f2)])])]))
    assert _tree(LE.Directory.from_leaf_files([f1, f2], t)) == _tree(expected)
    assert _tree(LE.Directory.from_directory(temp, t)) == _tree(expected)
    # testDir/temp2/a/f1, a/b/f2, a/c/f3
    temp2 = str(tmp_path / "testDir" / "temp2")
    g1, g2, g3 = _touch(f"{temp2}/a/f1"), _touch(f"{temp2}/a/b/f2"), _touch(f"{temp2}/a/c/f3")
    expected = _under_root(temp2, LE.Directory("temp2", subDirs=[
        LE.Directory("a", [info(g1)], [LE.Directory("b", [info(g2)]), LE.Directory("c", [info(g3)])])]))
    assert _tree(LE.Directory.from_leaf_files([g1, g2, g3], t)) == _tree(expected)
    assert _tree(LE.Directory.from_directory(f"{temp2}/a", t)) == _tree(expected)


def test_directory_path_filter_adds_only_valid_files(tmp_path):
    """'Directory Test: pathfilter adds only valid files': names starting with '_' or '.' are not data files."""
    d = str(tmp_path / "testDir")
    keep = _touch(f"{d}/f1")
    for hidden in ("_SUCCESS", ".f2.crc", "_committed_1", ".hidden"):
        _touch(f"{d}/{hidden}")
    t = FileIdTracker()
    got = LE.Directory.from_directory(d, t)
    assert [os.path.basename(f) for f in LE.Content(got).files] == ["f1"]
    assert t.get_file_id(*keep) == 0 and t.max_file_id == 0


def test_directory_merge_cases():
    F = lambda n, i: FileInfo(n, 100, 100, i)  # noqa: E731
    d1 = LE.Directory("a", [F("f1", 1), F("f2", 2)])
    d2 = LE.Directory("a", subDirs=[LE.Directory("b", [F("f3", 3), F("f4", 4)])])
    expected = LE.Directory("a", [F("f1", 1), F("f2", 2)], [LE.Directory("b", [F("f3", 3), F("f4", 4)])])
    assert _tree(d1.merge(d2)) == _tree(expected) and _tree(d2.merge(d1)) == _tree(expected)
    # overlapping directories
    d1 = LE.Directory("a", [F("f1", 1), F("f2", 2)], [LE.Directory("b", [F("f3", 3)])])
    d2 = LE.Directory("a", [F("f4", 4)], [LE.Directory("b", [F("f5", 5), F("f6", 6)], [LE.Directory("c", [F("f7", 7)])])])
    expected = LE.Directory("a", [F("f1", 1), F("f2", 2), F("f4", 4)],
                            [LE.Directory("b", [F("f3", 3), F("f5", 5), F("f6", 6)], [LE.Directory("c", [F("f7", 7)])])])
    assert _tree(d1.merge(d2)) == _tree(expected) and _tree(d2.merge(d1)) == _tree(expected)
    # different names
    a, b = LE.Directory("a", [F("f1", 1)]), LE.Directory("b", [F("f3", 3)])
    with pytest.raises(HyperspaceException, match="Merging directories with names a and b failed."):
        a.merge(b)
    with pytest.raises(HyperspaceException, match="Merging directories with names b and a failed."):
        b.merge(a)


# ---------------------------------------------------------------------------------------------------------------------
# T/index/covering/FilterIndexRankerTest.scala, JoinIndexRankerTest.scala
# ---------------------------------------------------------------------------------------------------------------------

def _cand(name, num_buckets=200, index_file_sizes=(10,), common_bytes=0):
    from hyperspace_b200 import rules as R

    e = _entry("ACTIVE")
    e.name, e.numBuckets = name, num_buckets
    files = [(f"file:/indexes/{name}/v__=0/f{i}.parquet", s, 1) for i, s in enumerate(index_file_sizes)]
    e.content = LE.Content.from_leaf_files(files, FileIdTracker())
    return R.Candidate(e, [], [], common_bytes)


def _session(hybrid=False):
    from hyperspace_b200.session import HyperspaceSession

    s = HyperspaceSession()
    s.conf.set("spark.hyperspace.index.hybridscan.enabled", "true" if hybrid else "false")
    assert s.conf.hybrid_scan_enabled == hybrid
    return s


def test_filter_ranker_prefers_the_smallest_index_by_default():
    """'rank() should return the index with smallest size by default.' (ind1: 2 files, ind2: 1 file, ind3: 3 files)"""
    from hyperspace_b200 import rules as R

    ind1, ind2, ind3 = _cand("ind1", index_file_sizes=(10, 10)), _cand("ind2", index_file_sizes=(10,)), _cand("ind3", index_file_sizes=(10, 10, 10))
    assert R.rank_filter_candidates(_session(), [ind1, ind2, ind3]) is ind2
    assert R.rank_filter_candidates(_session(), []) is None


def test_filter_ranker_prefers_largest_common_bytes_under_hybrid_scan():
    """'rank() should return the index with the largest common bytes of source files if HybridScan is enabled.'"""
    from hyperspace_b200 import rules as R

    ind1, ind2, ind3 = _cand("ind1", common_bytes=2), _cand("ind2", common_bytes=4), _cand("ind3", common_bytes=2)
    assert R.rank_filter_candidates(_session(hybrid=True), [ind1, ind2, ind3]) is ind2
    assert R.rank_filter_candidates(_session(hybrid=False), [ind1, ind2, ind3]) is ind1  # equal sizes: the first one


def test_join_ranker_prefers_equal_bucket_pairs_then_more_buckets():
    from hyperspace_b200 import rules as R

    l10, l20, r10, r20 = _cand("l1", 10), _cand("l2", 20), _cand("r1", 10), _cand("r2", 20)
    # 'rank() should prefer equal-bucket index pairs over unequal-bucket.'
    assert R.rank_join_pairs(_session(), [(l10, r20), (l20, r20)]) == [(l20, r20), (l10, r20)]
    # 'rank() should prefer higher number of buckets if multiple equal-bucket index pairs found.'
    assert R.rank_join_pairs(_session(), [(l10, r10), (l10, r20), (l20, r20)]) == [(l20, r20), (l10, r10), (l10, r20)]


def test_join_ranker_prefers_largest_common_bytes_under_hybrid_scan():
    """'rank() should prefer the largest common bytes if HybridScan is enabled.' (fileList1 = 3 bytes, fileList2 = 2 bytes)"""
    from hyperspace_b200 import rules as R

    l10, l20 = _cand("l1", 10, common_bytes=3), _cand("l2", 20, common_bytes=2)
    r10, r20 = _cand("r1", 10, common_bytes=3), _cand("r2", 20, common_bytes=2)
    pairs = [(l10, r10), (l10, r20), (l20, r20)]
    assert R.rank_join_pairs(_session(hybrid=False), pairs) == [(l20, r20), (l10, r10), (l10, r20)]
    assert R.rank_join_pairs(_session(hybrid=True), pairs) == [(l10, r10), (l20, r20), (l10, r20)]
    # 'If both indexes have the same amount of common bytes, follow the original algorithm.'
    l10, l20, r10, r20 = (_cand(n, b, common_bytes=3) for n, b in (("l1", 10), ("l2", 20), ("r1", 10), ("r2", 20)))
    pairs = [(l10, r10), (l10, r20), (l20, r20)]
    assert R.rank_join_pairs(_session(hybrid=True), pairs) == [(l20, r20), (l10, r10), (l10, r20)]


# ---------------------------------------------------------------------------------------------------------------------
# T/index/FileBasedSignatureProviderTest.scala, IndexSignatureProviderTest.scala, T/util/HashingUtilsTest.scala
# ---------------------------------------------------------------------------------------------------------------------

def _relation(files):
    from hyperspace_b200.session import RelationNode

    return RelationNode(["file:/data"], [(f"file:{p}", size, mtime) for size, mtime, p in files], [("c", "long")])


def test_md5_hashing_is_a_function_of_its_input():
    a, b = str(uuid.uuid4()), str(uuid.uuid4())
    assert LE.md5_hex(a) == LE.md5_hex(a) and LE.md5_hex(a) != LE.md5_hex(b)
    assert LE.md5_hex("") == "d41d8cd98f00b204e9800998ecf8427e"  # RFC 1321 test vector: commons-codec md5Hex agrees


def test_file_based_signature():
    length, mtime, path, new_path = 100, 10_000, "/data/f1", "/data/f2"
    sig = lambda files: _relation(files).signature  # noqa: E731
    assert sig([(length, mtime, path)]) == sig([(length, mtime, path)])                   # same file
    assert sig([(length, mtime, path)]) != sig([(length + 10, mtime, path)])              # different length
    assert sig([(length, mtime, path)]) != sig([(length, mtime + 3600, path)])            # different modification time
    assert sig([(length, mtime, path)]) != sig([(length, mtime, new_path)])               # different path
    two = [(length, mtime, path), (length + 10, mtime + 3600, new_path)]
    assert sig(two) == sig(list(two)) == sig(list(reversed(two)))                         # same files (sorted by path first)
    assert sig([(length, mtime, path), (length + 10, mtime, new_path)]) != sig([(length, mtime, path), (length, mtime + 3600, new_path)])


def test_index_signature_combines_file_and_plan_signatures():
    """IndexSignatureProvider.scala:33-51: md5(fileBasedSignature + planSignature); equal for equal relations only."""
    from hyperspace_b200 import rules as R

    r1, r1b, r2 = _relation([(100, 1, "/data/f1")]), _relation([(100, 1, "/data/f1")]), _relation([(101, 1, "/data/f1")])
    assert R.index_signature(r1) == R.index_signature(r1b) != R.index_signature(r2)
    assert R.index_signature(r1) == LE.md5_hex(LE.md5_hex(r1.signature) + LE.md5_hex("LogicalRelation"))


def test_json_round_trip_of_an_index_log_entry():
    """T/util/JsonUtilsTest.scala 'Test for JsonUtils.': fromJson(toJson(entry)) == entry, for an entry with no relations."""
    schema = {"type": "struct", "fields": [{"name": n, "type": t, "nullable": True, "metadata": {}}
                                           for n, t in (("id", "integer"), ("name", "string"), ("school", "string"))]}
    index = LE.IndexLogEntry(name="myIndex", indexedColumns=["id"], includedColumns=["name", "school"], schema=schema,
                             numBuckets=10, derived_properties={}, content=LE.Content(LE.Directory("path")), relations=[],
                             signatures=[LE.Signature("signatureProvider", "dfSignature")], state=States.ACTIVE)
    back = LE.IndexLogEntry.from_json(index.to_json())
    assert back == index and back.to_json() == index.to_json()
    assert back.numBuckets == 10 and back.schema == schema and back.state == "ACTIVE"


# ---------------------------------------------------------------------------------------------------------------------
# T/index/covering/FilterIndexRuleTest.scala (same index set-up: index1 on (c3, c2) incl. c1; index2, index3 on (c4, c2)
# incl. c1, c3; the transformed plan is recognised by the index name in the scan node)
# ---------------------------------------------------------------------------------------------------------------------

def _filter_rule_fixture(tmp_path):
    from hyperspace_b200 import rules as R
    from hyperspace_b200.session import DataFrame, HyperspaceSession, RelationNode

    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path / "indexes")}).enableHyperspace()
    rel = RelationNode([f"file:{tmp_path}/baseTableLocation"], [(f"file:{tmp_path}/baseTableLocation/f1", 100, 1)],
                       [("c1", "long"), ("c2", "long"), ("c3", "long"), ("c4", "integer")])  # literals are ints here: the GPU scan takes int keys

    def make(name, indexed, included):
        e = _entry("ACTIVE")
        e.name, e.indexedColumns, e.includedColumns, e.id = name, indexed, included, 1
        e.content = LE.Content.from_leaf_files([(f"file:{tmp_path}/indexes/{name}/v__=0/part-00000-x_00000.c000.parquet", 10, 1)],
                                               FileIdTracker())
        e.relations = [LE.Relation(rel.root_paths, LE.Content.from_leaf_files(rel.files, FileIdTracker()),
                                   {"type": "struct", "fields": []}, "parquet")]
        e.signatures = [LE.Signature(LE.INDEX_SIGNATURE_PROVIDER, R.index_signature(rel))]
        lm = LE.IndexLogManager(os.path.join(str(tmp_path / "indexes"), name))
        assert lm.write_log(1, e) and lm.create_latest_stable_log(1)

    make("index1", ["c3", "c2"], ["c1"])
    make("index2", ["c4", "c2"], ["c1", "c3"])
    make("index3", ["c4", "c2"], ["c1", "c3"])
    return s, DataFrame(s, rel)


def test_filter_index_rule_cases(tmp_path):
    from hyperspace_b200.session import col

    s, df = _filter_rule_fixture(tmp_path)
    # 'Verify FilterIndex rule is applied correctly.'
    assert "Name: index1" in df.filter(col("c3") == 7).select("c2", "c3").explain()
    # '... for case insensitive query.'
    assert "Name: index1" in df.filter(col("C3") == 7).select("C2", "C3").explain()
    # '... does not apply if all columns are not covered.' (c4 is not covered by index1; index2/3 do not start with c3)
    plan = df.filter(col("c3") == 7).select("c2", "c3", "c4").explain()
    assert "Name: index" not in plan and "GpuSourceScan" in plan
    # '... does not apply if filter does not contain first indexed column.' (c2 is not a first indexed column)
    plan = df.filter(col("c2") == 9).select("c2", "c3").explain()
    assert "Name: index" not in plan
    # '... is applied when all columns are selected.' (index2 and index3 tie; the first one is taken)
    assert "Name: index2" in df.filter(col("c4") == 10).explain()
    # disabled session: untouched plan
    s.disableHyperspace()
    assert "Name: index" not in df.filter(col("c4") == 10).explain()


# ---------------------------------------------------------------------------------------------------------------------
# T/index/covering/JoinIndexRuleTest.scala (cases expressible as single-column inner equi-joins, which is what the GPU
# merge join takes; same five indexes as the reference's fixture)
# ---------------------------------------------------------------------------------------------------------------------

def _join_rule_fixture(tmp_path):
    from hyperspace_b200 import rules as R
    from hyperspace_b200.session import DataFrame, HyperspaceSession, RelationNode

    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path / "indexes")}).enableHyperspace()

    def rel(t):
        return RelationNode([f"file:{tmp_path}/{t}"], [(f"file:{tmp_path}/{t}/f1", 100, 1)],
                            [(f"{t}c{i}", "long") for i in (1, 2, 3, 4)])

    def make(name, r, indexed, included):
        e = _entry("ACTIVE")
        e.name, e.indexedColumns, e.includedColumns, e.id = name, indexed, included, 1
        e.content = LE.Content.from_leaf_files([(f"file:{tmp_path}/indexes/{name}/v__=0/part-00000-x_00000.c000.parquet", 10, 1)],
                                               FileIdTracker())
        e.relations = [LE.Relation(r.root_paths, LE.Content.from_leaf_files(r.files, FileIdTracker()),
                                   {"type": "struct", "fields": []}, "parquet")]
        e.signatures = [LE.Signature(LE.INDEX_SIGNATURE_PROVIDER, R.index_signature(r))]
        lm = LE.IndexLogManager(os.path.join(str(tmp_path / "indexes"), name))
        assert lm.write_log(1, e) and lm.create_latest_stable_log(1)

    t1, t2 = rel("t1"), rel("t2")
    make("t1i1", t1, ["t1c1"], ["t1c3"])
    make("t1i2", t1, ["t1c1", "t1c2"], ["t1c3"])
    make("t1i3", t1, ["t1c2"], ["t1c3"])
    make("t2i1", t2, ["t2c1"], ["t2c3"])
    make("t2i2", t2, ["t2c1", "t2c2"], ["t2c3"])
    return s, DataFrame(s, t1), DataFrame(s, t2)


def test_join_index_rule_cases(tmp_path):
    s, t1, t2 = _join_rule_fixture(tmp_path)
    # 'Join rule works if indexes exist and configs are set correctly.': t1i1 and t2i1 (indexed columns == join columns)
    plan = t1.join(t2, on=("t1c1", "t2c1")).select("t1c3", "t2c3").explain()
    assert "Name: t1i1" in plan and "Name: t2i1" in plan and "exchange=none" in plan
    # '... for case insensitive index and query.'
    plan = t1.join(t2, on=("T1C1", "T2C1")).select("T1C3", "T2C3").explain()
    assert "Name: t1i1" in plan and "Name: t2i1" in plan
    # "... does not update plan if index doesn't exist for either table." (t1i3 is keyed on t1c2, nothing is on t2c2)
    plan = t1.join(t2, on=("t1c2", "t2c2")).select("t1c3", "t2c3").explain()
    assert "Name:" not in plan and "GpuShuffle" in plan
    # a column outside the indexes' coverage on one side keeps that side (and so the join) off the indexes
    plan = t1.join(t2, on=("t1c1", "t2c1")).select("t1c4", "t2c3").explain()
    assert "Name:" not in plan
    # only inner joins reach the rule at all
    with pytest.raises(HyperspaceException):
        t1.join(t2, on=("t1c1", "t2c1"), how="left")
    s.disableHyperspace()
    assert "Name:" not in t1.join(t2, on=("t1c1", "t2c1")).select("t1c3", "t2c3").explain()


# ---------------------------------------------------------------------------------------------------------------------
# T/index/rules/CandidateIndexCollectorTest.scala 'Verify CandidateIndexCollector for hybrid scan.'
# (indexes fabricated over a 4-file relation, one with the lineage column and one without; thresholds as in the reference)
# ---------------------------------------------------------------------------------------------------------------------

def test_candidate_index_collector_for_hybrid_scan(tmp_path):
    from hyperspace_b200 import rules as R
    from hyperspace_b200.session import HyperspaceSession, RelationNode

    base = [(f"file:{tmp_path}/data/f{i}", 100, 10 + i) for i in range(4)]

    def rel(files):
        return RelationNode([f"file:{tmp_path}/data"], list(files), [("id", "long")])

    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path / "indexes")}).enableHyperspace()
    for name, lineage in (("index1", True), ("index2", False)):
        e = _entry("ACTIVE")
        e.name, e.indexedColumns, e.includedColumns, e.id = name, ["id"], [], 1
        e.derived_properties = {LE.LINEAGE_PROPERTY: "true" if lineage else "false"}
        e.content = LE.Content.from_leaf_files([(f"file:{tmp_path}/indexes/{name}/v__=0/part-00000-x_00000.c000.parquet", 10, 1)],
                                               FileIdTracker())
        e.relations = [LE.Relation([f"file:{tmp_path}/data"], LE.Content.from_leaf_files(base, FileIdTracker()),
                                   {"type": "struct", "fields": []}, "parquet")]
        e.signatures = [LE.Signature(LE.INDEX_SIGNATURE_PROVIDER, R.index_signature(rel(base)))]
        lm = LE.IndexLogManager(os.path.join(str(tmp_path / "indexes"), name))
        assert lm.write_log(1, e) and lm.create_latest_stable_log(1)

    def verify(files, hybrid, delete_enabled, expected_names, expected_hybrid_required=None, expected_common=None):
        s.conf.set("spark.hyperspace.index.hybridscan.enabled", "true" if hybrid else "false")
        s.conf.set("spark.hyperspace.index.hybridscan.maxAppendedRatio", "0.99")
        s.conf.set("spark.hyperspace.index.hybridscan.maxDeletedRatio", "0.99" if delete_enabled else "0")
        cands = R.candidates_for(s, rel(files))
        assert sorted(c.entry.name for c in cands) == sorted(expected_names)
        for c in cands:
            if expected_hybrid_required is not None:
                assert bool(c.appended or c.deleted_ids) == expected_hybrid_required
            if expected_common is not None:
                assert c.common_bytes == expected_common

    # unmodified source: candidates whether Hybrid Scan is enabled or not
    verify(base, False, False, ["index1", "index2"])
    verify(base, True, False, ["index1", "index2"], expected_hybrid_required=False, expected_common=400)
    # Scenario #1: append new files
    appended = base + [(f"file:{tmp_path}/data/g{i}", 100, 50 + i) for i in range(4)]
    verify(appended, False, False, [])
    verify(appended, True, False, ["index1", "index2"], expected_hybrid_required=True, expected_common=400)
    # Scenario #2: delete one file (needs the lineage column and a non-zero delete threshold)
    deleted = appended[1:]
    verify(deleted, False, False, [])
    verify(deleted, True, False, [])
    verify(deleted, True, True, ["index1"], expected_hybrid_required=True, expected_common=300)
    # Scenario #3: replace all files
    replaced = [(f"file:{tmp_path}/data/h{i}", 100, 90 + i) for i in range(4)]
    verify(replaced, False, False, [])
    verify(replaced, True, True, [])


# ---------------------------------------------------------------------------------------------------------------------
# T/HyperspaceConfTest.scala 'Test configs that support legacy configs'
# ---------------------------------------------------------------------------------------------------------------------

def test_num_buckets_conf_supports_the_legacy_key():
    from hyperspace_b200 import session as S

    conf = S.HyperspaceSession().conf
    legacy, new = S.INDEX_NUM_BUCKETS_LEGACY, S.INDEX_NUM_BUCKETS

    def clear():
        conf.unset(legacy)
        conf.unset(new)

    clear()
    assert conf.num_buckets == S.INDEX_NUM_BUCKETS_DEFAULT == 200   # default if no key is set
    conf.set(legacy, 10)
    assert conf.num_buckets == 10                                   # only the legacy key
    clear()
    conf.set(new, 5)
    assert conf.num_buckets == 5                                    # only the new key
    clear()
    conf.set(legacy, 10)
    conf.set(new, 5)
    assert conf.num_buckets == 5                                    # both: the new key wins
    assert (legacy, new) == ("spark.hyperspace.index.num.buckets", "spark.hyperspace.index.numBuckets")


# ---------------------------------------------------------------------------------------------------------------------
# T/actions/RefreshActionTest.scala (validate() only: no data path involved)
# ---------------------------------------------------------------------------------------------------------------------

def _refresh_fixture(tmp_path, state):
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq

    from hyperspace_b200.session import HyperspaceSession

    data = tmp_path / "sampleparquet"
    data.mkdir(exist_ok=True)
    pq.write_table(pa.table({"clicks": np.arange(10, dtype=np.int32), "imprs": np.arange(10, dtype=np.int64)}), str(data / "part-0.parquet"))
    files = [LE.file_status(str(data / "part-0.parquet"))]
    e = _entry(state)
    e.name, e.indexedColumns, e.includedColumns, e.numBuckets = "index1", ["clicks"], [], 10
    e.relations = [LE.Relation([LE.to_uri(str(data))], LE.Content.from_leaf_files(files, FileIdTracker()),
                               {"type": "struct", "fields": []}, "parquet")]
    lm = _RecordingLogManager(latest_id=None, log=e)
    dm = _RecordingDataManager([])
    dm.get_path = lambda id_: str(tmp_path / "indexPath")
    return HyperspaceSession(), lm, dm, data


def _append_source_file(data):
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq

    pq.write_table(pa.table({"clicks": np.arange(5, dtype=np.int32), "imprs": np.arange(5, dtype=np.int64)}), str(data / "part-1.parquet"))


def test_refresh_action_validate(tmp_path):
    from hyperspace_b200.hyperspace import NoChangesException, RefreshAction

    # 'validate() passes if old index logs are found with ACTIVE state'
    s, lm, dm, data = _refresh_fixture(tmp_path, "ACTIVE")
    _append_source_file(data)
    RefreshAction(s, lm, dm).validate()
    # 'validate() fails if old index logs found with non-ACTIVE state'
    s, lm, dm, data = _refresh_fixture(tmp_path, "CREATING")
    with pytest.raises(HyperspaceException, match="Refresh is only supported in ACTIVE state"):
        RefreshAction(s, lm, dm).validate()
    # 'validate() fails if there is no source data change.'
    (data / "part-1.parquet").unlink()
    s, lm, dm, data = _refresh_fixture(tmp_path, "ACTIVE")
    with pytest.raises(NoChangesException, match="Refresh full aborted as no source data changed."):
        RefreshAction(s, lm, dm).validate()
    # and run() treats that as a no-op: nothing is written to the log (Action.scala:96-99)
    RefreshAction(s, lm, dm).run()
    assert lm.calls == []


# ---------------------------------------------------------------------------------------------------------------------
# T/actions/CreateActionTest.scala (validate() cases)
# ---------------------------------------------------------------------------------------------------------------------

def test_create_action_validate(tmp_path):
    from hyperspace_b200.hyperspace import CreateAction
    from hyperspace_b200.session import DataFrame

    s, lm, dm, data = _refresh_fixture(tmp_path, "ACTIVE")
    df = s.read.parquet(str(data))
    cfg = IndexConfig("index1", ["clicks"], ["imprs"])
    lm.get_latest_log = lambda: None
    # 'validate passes for valid index config and df' / '... if no earlier index logs are found'
    CreateAction(s, df, cfg, lm, dm).validate()
    # 'validate() fails if df is not logical plan' (here: anything but a bare file-based relation)
    with pytest.raises(HyperspaceException, match="Only creating index over HDFS file based scan nodes is supported."):
        CreateAction(s, df.select("clicks"), IndexConfig("name", ["clicks"]), lm, dm).validate()
    # "validate() fails if index config doesn't contain columns from df"
    with pytest.raises(HyperspaceException, match="Index config is not applicable to dataframe schema"):
        CreateAction(s, df, IndexConfig("name", ["c1"], ["c2"]), lm, dm).validate()
    # 'validate() passes if old index logs are found with DOESNOTEXIST state'
    lm.get_latest_log = lambda: _entry("DOESNOTEXIST")
    CreateAction(s, df, cfg, lm, dm).validate()
    # 'validate() fails if old index logs found with non-DOESNOTEXIST state'
    lm.get_latest_log = lambda: _entry("ACTIVE")
    with pytest.raises(HyperspaceException, match="Another Index with name index1 already exists"):
        CreateAction(s, df, cfg, lm, dm).validate()
    # index config columns resolve case-insensitively and the log entry keeps the source's spelling (ResolverUtils)
    lm.get_latest_log = lambda: None
    a = CreateAction(s, df, IndexConfig("index1", ["CLICKS"], ["Imprs"]), lm, dm)
    a.validate()
    e = a.log_entry()
    assert e.indexedColumns == ["clicks"] and e.includedColumns == ["imprs"] and e.numBuckets == 200
    assert isinstance(df, DataFrame)


# ---------------------------------------------------------------------------------------------------------------------
# T/actions/VacuumOutdatedActionTest.scala
# ---------------------------------------------------------------------------------------------------------------------

def test_vacuum_outdated_action(tmp_path):
    from hyperspace_b200.hyperspace import VacuumOutdatedAction

    dm0 = _RecordingDataManager([])
    # validate(): ACTIVE only, with the reference's message
    VacuumOutdatedAction(_RecordingLogManager(log=_entry("ACTIVE")), dm0).validate()
    with pytest.raises(HyperspaceException, match="VacuumOutdated is only supported in ACTIVE state. Current state is CREATING."):
        VacuumOutdatedAction(_RecordingLogManager(log=_entry("CREATING")), dm0).validate()

    def fixture(all_versions, live_files):
        index_path = str(tmp_path / str(uuid.uuid4()))
        for v in all_versions:
            _touch(os.path.join(index_path, f"v__={v}", f"stale-{v}.parquet"))
        statuses = [_touch(os.path.join(index_path, rel)) for rel in live_files]
        e = _entry("ACTIVE")
        e.content = LE.Content.from_leaf_files(statuses, FileIdTracker())
        return index_path, LE.IndexDataManager(index_path), _RecordingLogManager(log=e)

    # 'op() calls which deletes nothing since every data is up-to-date' (versions 0, 1, 2 all referenced)
    path, dm, lm = fixture([0, 1, 2], ["v__=0/a.parquet", "v__=1/b.parquet", "v__=2/part-00053-.c000.snappy.parquet"])
    VacuumOutdatedAction(lm, dm).op()
    assert dm.get_all_version_ids() == [0, 1, 2]
    # 'op() calls delete for all outdated data': versions 0 and 1 go, 2 and 3 stay -- minus the files the entry does not list
    path, dm, lm = fixture([0, 1, 2, 3], ["v__=2/part-00053-.c000.snappy.parquet", "v__=2/part-00027-.c000.snappy.parquet",
                                           "v__=3/part-00001-.c000.snappy.parquet"])
    VacuumOutdatedAction(lm, dm).op()
    assert dm.get_all_version_ids() == [2, 3]
    assert sorted(os.listdir(os.path.join(path, "v__=2"))) == ["part-00027-.c000.snappy.parquet", "part-00053-.c000.snappy.parquet"]
    assert os.listdir(os.path.join(path, "v__=3")) == ["part-00001-.c000.snappy.parquet"]


def test_index_version_directories_of_a_log_entry():
    """'versionInfos gets correct version info.': the v__=N directories the entry's content refers to."""
    u = LE.UNKNOWN_FILE_ID
    version_dirs = [LE.Directory(f"v__={v}", files=[FileInfo(f"index_{v}", 0, 0, u)]) for v in (4, 5)]
    e = _entry("ACTIVE")
    e.content = LE.Content(LE.Directory("file:/", subDirs=[LE.Directory(
        "a", files=[FileInfo("f1", 0, 0, u), FileInfo("f2", 0, 0, u)],
        subDirs=[LE.Directory("b", files=[FileInfo("f3", 0, 0, u), FileInfo("f4", 0, 0, u)], subDirs=version_dirs)])]))
    assert e.index_version_dirs() == [4, 5]

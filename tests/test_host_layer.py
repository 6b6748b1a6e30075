"""CPU-only tests of the host layer: on-disk log layout, action state machine, index configs, rule conditions.

Modelled on the reference's unit tests that need no data: T/index/IndexLogEntryTest.scala (spec JSON, Content/Directory,
FileIdTracker), T/index/IndexLogManagerImplTest.scala, T/actions/ActionTest.scala, T/index/IndexConfigTest.scala,
T/index/covering/FilterIndexRuleTest.scala / JoinIndexRuleTest.scala (rules over fabricated ACTIVE index entries, as
T/index/HyperspaceRuleSuite.scala:35-85 does).
"""
import json
import os

import pytest

from hyperspace_b200 import log_entry as LE
from hyperspace_b200 import rules as R
from hyperspace_b200.hyperspace import Hyperspace, _StateFlip
from hyperspace_b200.index_config import CoveringIndexConfig, IndexConfig
from hyperspace_b200.session import HyperspaceSession, RelationNode, col, DataFrame

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _spec() -> str:
    return open(os.path.join(GOLDEN, "index_log_entry_spec.json")).read()


def test_golden_hash_vectors_match_oracle():
    import numpy as np

    from oracle import oracle as O

    v = json.load(open(os.path.join(GOLDEN, "spark_hash_vectors.json")))
    bu = v["bucket_union_test"]
    assert O.bucket_ids([np.array(bu["keys_int32"], dtype=np.int32)], bu["num_partitions"]).tolist() == bu["partitions"]
    for k, h in v["hash_long_seed42"].items():
        assert O.lib().hso_hash_long(int(k), 42) == h
        assert int(O.np_hash_long(np.array([int(k)]))[0]) == h


def test_index_log_entry_spec_example():
    """IndexLogEntryTest.scala:74-224: the spec JSON parses to the expected entry and sourceFilesSizeInBytes == 200."""
    e = LE.IndexLogEntry.from_json(_spec())
    assert e.name == "indexName" and e.indexedColumns == ["col1"] and e.includedColumns == ["col2", "col3"]
    assert e.numBuckets == 200 and e.state == "ACTIVE" and e.timestamp == 1578818514080 and e.enabled and e.id == 0
    assert e.schema["fields"][0]["name"] == "RGUID"
    assert e.relations[0].rootPaths == ["rootpath"] and e.relations[0].fileFormat == "type"
    assert [(f.name, f.size, f.modifiedTime, f.id) for f in e.source_file_infos] == [("test/f1", 100, 100, 0), ("test/f2", 100, 200, 1)]
    assert e.source_files_size_in_bytes == 200
    assert [(f.name, f.id) for f in e.deleted_files] == [("/f1", 2)] or [(f.name, f.id) for f in e.deleted_files] == [("f1", 2)]
    assert e.appended_files == []
    assert e.signatures == [LE.Signature("provider", "signatureValue")]
    assert e.properties == {"hyperspaceVersion": "0.5.0-SNAPSHOT"}
    # serialisation keeps every field of the spec (same keys, same nesting, same values)
    assert json.loads(e.to_json()) == json.loads(_spec())
    assert LE.IndexLogEntry.from_json(e.to_json()).to_json() == e.to_json()


def test_content_files_and_directory_tree(tmp_path):
    """IndexLogEntryTest 'Content.files api lists all files' + Directory.fromDirectory / fromLeafFiles."""
    for rel in ("a/f1", "a/f2", "a/b/f3", "a/b/_SUCCESS", "a/.hidden"):
        p = tmp_path / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text("x")
    t = LE.FileIdTracker()
    c = LE.Content.from_directory(str(tmp_path / "a"), t)
    names = sorted(os.path.relpath(LE.from_uri(f), str(tmp_path)) for f in c.files)
    assert names == ["a/b/f3", "a/f1", "a/f2"]  # DataPathFilter drops _SUCCESS and dot files
    assert c.root.name == "file:/"
    assert sorted(f.id for f in c.file_infos) == [0, 1, 2] and t.max_file_id == 2
    empty = LE.Content.from_directory(str(tmp_path / "missing"), t)
    assert empty.files == []


def test_directory_merge():
    a = LE.Directory("file:/", subDirs=[LE.Directory("a", subDirs=[LE.Directory("b", files=[LE.FileInfo("f1", 1, 1, 0), LE.FileInfo("f2", 1, 1, 1)])])])
    b = LE.Directory("file:/", subDirs=[LE.Directory("a", files=[LE.FileInfo("f3", 1, 1, 2), LE.FileInfo("f4", 1, 1, 3)])])
    m = a.merge(b)
    assert sorted(LE.Content(m).files) == ["file:/a/b/f1", "file:/a/b/f2", "file:/a/f3", "file:/a/f4"]
    with pytest.raises(LE.HyperspaceException):
        LE.Directory("x").merge(LE.Directory("y"))


def test_file_id_tracker():
    t = LE.FileIdTracker()
    assert t.max_file_id == -1
    assert t.add_file("file:/a", 10, 1) == 0 and t.add_file("file:/b", 10, 1) == 1 and t.add_file("file:/a", 10, 1) == 0
    assert t.add_file("file:/a", 11, 1) == 2  # a new version of the same path is a new file
    t.add_file_info([LE.FileInfo("file:/c", 5, 5, 7)])
    assert t.max_file_id == 7 and t.get_file_id("file:/c", 5, 5) == 7
    with pytest.raises(LE.HyperspaceException):
        t.add_file_info([LE.FileInfo("file:/c", 5, 5, 8)])
    with pytest.raises(LE.HyperspaceException):
        t.add_file_info([LE.FileInfo("file:/d", 5, 5, LE.UNKNOWN_FILE_ID)])


def _entry(state="ACTIVE", name="idx") -> LE.IndexLogEntry:
    e = LE.IndexLogEntry.from_json(_spec())
    e.name, e.state = name, state
    return e


def test_log_manager_optimistic_concurrency_and_latest_stable(tmp_path):
    lm = LE.IndexLogManager(str(tmp_path / "idx"))
    assert lm.get_latest_id() is None and lm.get_latest_stable_log() is None
    assert lm.write_log(0, _entry("CREATING"))
    assert not lm.write_log(0, _entry("CREATING"))  # the loser of the race gets False
    assert lm.write_log(1, _entry("ACTIVE"))
    assert lm.get_latest_id() == 1
    assert lm.get_latest_stable_log().state == "ACTIVE"  # found by scanning back when latestStable is absent
    assert lm.create_latest_stable_log(1) and os.path.exists(tmp_path / "idx" / "_hyperspace_log" / "latestStable")
    assert not lm.create_latest_stable_log(0)  # CREATING is not a stable state
    assert lm.write_log(2, _entry("REFRESHING"))
    assert lm.get_latest_stable_log().state == "ACTIVE"
    assert lm.delete_latest_stable_log() and lm.get_latest_stable_log().state == "ACTIVE"
    assert lm.get_index_versions(["ACTIVE"]) == [1]


def test_action_protocol_delete_restore_cancel(tmp_path):
    """ActionTest.scala:55-63: begin writes base+1 (transient), end deletes latestStable, writes base+2 (final), recreates it."""
    lm = LE.IndexLogManager(str(tmp_path / "idx"))
    lm.write_log(0, _entry("CREATING"))
    lm.write_log(1, _entry("ACTIVE"))
    lm.create_latest_stable_log(1)
    _StateFlip(lm, "ACTIVE", "DELETING", "DELETED", "Delete").run()
    assert lm.get_log(2).state == "DELETING" and lm.get_log(3).state == "DELETED" and lm.get_latest_stable_log().id == 3
    with pytest.raises(LE.HyperspaceException):  # delete is only valid from ACTIVE
        _StateFlip(lm, "ACTIVE", "DELETING", "DELETED", "Delete").run()
    _StateFlip(lm, "DELETED", "RESTORING", "ACTIVE", "Restore").run()
    assert lm.get_log(5).state == "ACTIVE"
    # a crashed action leaves a transient entry; cancel rolls back to the last stable state
    lm.write_log(6, _entry("REFRESHING"))
    from hyperspace_b200.hyperspace import CancelAction

    CancelAction(lm).run()
    assert lm.get_log(7).state == "CANCELLING" and lm.get_log(8).state == "ACTIVE"
    with pytest.raises(LE.HyperspaceException):
        CancelAction(lm).run()


def test_index_config_validation():
    """IndexConfigTest.scala: empty names / columns and duplicates are rejected; equality is case-insensitive."""
    for bad in (lambda: IndexConfig("", ["a"]), lambda: IndexConfig("i", []), lambda: IndexConfig("i", ["a", "A"]),
                lambda: IndexConfig("i", ["a"], ["b", "B"]), lambda: IndexConfig("i", ["a"], ["A"])):
        with pytest.raises(ValueError):
            bad()
    assert IndexConfig("Idx", ["A"], ["b", "c"]) == CoveringIndexConfig("idx", ["a"], ["C", "B"])
    assert IndexConfig is CoveringIndexConfig


def test_conf_defaults_and_legacy_key():
    s = HyperspaceSession()
    assert s.conf.num_buckets == 200  # T/index/IndexManagerTest.scala:91
    s.conf.set("spark.hyperspace.index.num.buckets", 10)
    assert s.conf.num_buckets == 10
    s.conf.set("spark.hyperspace.index.numBuckets", 20)
    assert s.conf.num_buckets == 20
    assert not s.conf.lineage_enabled and not s.conf.hybrid_scan_enabled
    assert s.conf.hybrid_scan_appended_ratio == 0.3 and s.conf.hybrid_scan_deleted_ratio == 0.2
    assert not s.isHyperspaceEnabled() and s.enableHyperspace().isHyperspaceEnabled()


def test_path_resolver_is_case_insensitive(tmp_path):
    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path)})
    (tmp_path / "MyIndex").mkdir()
    assert LE.PathResolver(s.conf).get_index_path("myindex") == str(tmp_path / "MyIndex")
    assert LE.PathResolver(s.conf).get_index_path("other") == str(tmp_path / "other")


# ---------------------------------------------------------------------------------------------------------------------
# rules over fabricated indexes (no index data, no GPU) -- like HyperspaceRuleSuite.createIndexLogEntry
# ---------------------------------------------------------------------------------------------------------------------

def _fabricate(tmp_path, session, name, rel: RelationNode, indexed, included, num_buckets=200, index_bytes=10):
    tracker = LE.FileIdTracker()
    idx_files = [(f"file:{tmp_path}/indexes/{name}/v__=0/part-00000-x_{b:05d}.c000.parquet", index_bytes, 1) for b in range(2)]
    e = LE.IndexLogEntry(
        name=name, indexedColumns=indexed, includedColumns=included, schema={"type": "struct", "fields": []},
        numBuckets=num_buckets, derived_properties={"lineage": "false"}, content=LE.Content.from_leaf_files(idx_files, LE.FileIdTracker()),
        relations=[LE.Relation(rel.root_paths, LE.Content.from_leaf_files(rel.files, tracker), {"type": "struct", "fields": []}, "parquet")],
        signatures=[LE.Signature(LE.INDEX_SIGNATURE_PROVIDER, R.index_signature(rel))], state="ACTIVE", id=1)
    lm = LE.IndexLogManager(os.path.join(LE.PathResolver(session.conf).system_path, name))
    lm.write_log(1, e)
    lm.create_latest_stable_log(1)
    return e


def _rel(path, ncols=("k", "v1", "v2"), files=(("f1", 100, 1), ("f2", 100, 2))):
    return RelationNode([f"file:{path}"], [(f"file:{path}/{n}", s, m) for n, s, m in files], [(c, "long") for c in ncols])


def test_filter_index_rule_conditions(tmp_path):
    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path / "indexes")}).enableHyperspace()
    rel = _rel(tmp_path / "t")
    _fabricate(tmp_path, s, "big", rel, ["k"], ["v1", "v2"], index_bytes=1000)
    _fabricate(tmp_path, s, "small", rel, ["k"], ["v1"], index_bytes=10)
    _fabricate(tmp_path, s, "other", rel, ["v1"], ["k"])
    df = DataFrame(s, rel)
    # first indexed column must be in the filter and the index must cover all referenced columns
    assert "Name: small" in df.filter(col("k") >= 1).select("k", "v1").explain()            # smallest covering index wins
    assert "Name: big" in df.filter(col("k") >= 1).select("k", "v2").explain()              # only 'big' covers v2
    assert "Name: other" in df.filter(col("v1") == 3).select("k").explain()
    assert "GpuSourceScan" in df.filter(col("v2") >= 1).select("k").explain()               # no index starts with v2
    assert "GpuSourceScan" in df.select("k").explain()                                      # no filter -> FilterIndexRule does not apply
    s.disableHyperspace()
    assert "GpuSourceScan" in df.filter(col("k") >= 1).select("k", "v1").explain()
    # a changed source (different signature) disqualifies the index unless Hybrid Scan is on
    s.enableHyperspace()
    rel2 = _rel(tmp_path / "t", files=(("f1", 100, 1), ("f2", 100, 2), ("f3", 10, 3)))
    df2 = DataFrame(s, rel2)
    assert "GpuSourceScan" in df2.filter(col("k") >= 1).select("k", "v1").explain()
    s.conf.set("spark.hyperspace.index.hybridscan.enabled", True)
    assert "hybridScan(appended=1" in df2.filter(col("k") >= 1).select("k", "v1").explain()
    rel3 = _rel(tmp_path / "t", files=(("f1", 100, 1), ("f2", 100, 2), ("f3", 1000, 3)))    # appended ratio 1000/1200 > 0.3
    assert "GpuSourceScan" in DataFrame(s, rel3).filter(col("k") >= 1).select("k", "v1").explain()
    rel4 = _rel(tmp_path / "t", files=(("f1", 100, 1),))                                    # deleted file, no lineage
    assert "GpuSourceScan" in DataFrame(s, rel4).filter(col("k") >= 1).select("k", "v1").explain()


def test_join_index_rule_conditions(tmp_path):
    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path / "indexes")}).enableHyperspace()
    lrel, rrel = _rel(tmp_path / "l", ("k", "a")), _rel(tmp_path / "r", ("k", "b"))
    _fabricate(tmp_path, s, "l200", lrel, ["k"], ["a"], 200)
    _fabricate(tmp_path, s, "l50", lrel, ["k"], ["a"], 50)
    _fabricate(tmp_path, s, "r200", rrel, ["k"], ["b"], 200)
    _fabricate(tmp_path, s, "r_wrong", rrel, ["b"], ["k"], 200)
    plan = DataFrame(s, lrel).join(DataFrame(s, rrel), on="k").select("a", "b").explain()
    assert "Name: l200" in plan and "Name: r200" in plan and "exchange=none" in plan  # equal bucket counts pair up
    # join key must equal the indexed columns on both sides
    plan2 = DataFrame(s, lrel).join(DataFrame(s, rrel), on=("k", "b")).select("a", "k").explain()
    assert "Name: l200" in plan2 and "Name: r_wrong" in plan2   # l.k = r.b: the right index on b is the match
    plan3 = DataFrame(s, lrel).join(DataFrame(s, rrel), on=("a", "k")).select("a", "b").explain()
    assert "GpuShuffle" in plan3 and "Name:" not in plan3       # no left index is keyed on a
    s.disableHyperspace()
    assert "Name:" not in DataFrame(s, lrel).join(DataFrame(s, rrel), on="k").select("a", "b").explain()


def test_hyperspace_indexes_listing_and_errors(tmp_path):
    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path / "indexes")})
    hs = Hyperspace(s)
    assert hs.indexes() == []
    rel = _rel(tmp_path / "t")
    _fabricate(tmp_path, s, "idx1", rel, ["k"], ["v1"])
    got = hs.indexes()
    assert [(i["name"], i["state"], i["numBuckets"]) for i in got] == [("idx1", "ACTIVE", 200)]
    assert hs.index("IDX1")["indexedColumns"] == ["k"]
    with pytest.raises(LE.HyperspaceException):
        hs.deleteIndex("nope")
    hs.deleteIndex("idx1")
    assert hs.indexes()[0]["state"] == "DELETED"
    with pytest.raises(LE.HyperspaceException):
        hs.deleteIndex("idx1")
    hs.restoreIndex("idx1")
    assert hs.indexes()[0]["state"] == "ACTIVE"
    hs.deleteIndex("idx1")
    hs.vacuumIndex("idx1")
    assert hs.indexes() == []
    with pytest.raises(LE.HyperspaceException):
        hs.refreshIndex("idx1", "bogus")

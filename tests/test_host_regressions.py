"""Regression tests of host-layer defects found in review (CPU only, no index data is read).

* VacuumOutdated must compare canonical paths (a non-normalised spark.hyperspace.system.path once made it delete live files).
* The lineage column is recorded in the log entry's schema, never in includedColumns (CoveringIndex.createIndexData,
  src/main/scala/com/microsoft/hyperspace/index/covering/CoveringIndex.scala:152-186).
* refreshIndex(mode="quick") leaves an index usable without the hybridscan conf (CoveringIndexRuleUtils.scala:68-84).
* A join never picks an index whose source lost files (it would need the lineage NOT-IN filter).
* Comparison operators round non-integral literals towards the predicate's meaning.
"""
import os

import pytest

from hyperspace_b200 import log_entry as LE
from hyperspace_b200 import rules as R
from hyperspace_b200.hyperspace import VacuumOutdatedAction, _DataAction
from hyperspace_b200.session import DataFrame, HyperspaceSession, RelationNode, col


def _rel(path, ncols=("k", "v1", "v2"), files=(("f1", 100, 1), ("f2", 100, 2))):
    return RelationNode([f"file:{path}"], [(f"file:{path}/{n}", s, m) for n, s, m in files], [(c, "long") for c in ncols])


def _entry_over(index_root: str, name: str, files, lineage="false") -> LE.IndexLogEntry:
    rel = _rel("/src")
    idx_files = [(LE.to_uri(p), os.path.getsize(p), 1) for p in files]
    return LE.IndexLogEntry(
        name=name, indexedColumns=["k"], includedColumns=["v1"], schema={"type": "struct", "fields": []}, numBuckets=2,
        derived_properties={"lineage": lineage}, content=LE.Content.from_leaf_files(idx_files, LE.FileIdTracker()),
        relations=[LE.Relation(rel.root_paths, LE.Content.from_leaf_files(rel.files, LE.FileIdTracker()),
                               {"type": "struct", "fields": []}, "parquet")],
        signatures=[LE.Signature(LE.INDEX_SIGNATURE_PROVIDER, R.index_signature(rel))], state="ACTIVE", id=1)


@pytest.mark.parametrize("spelling", ["double_slash", "trailing_slash", "dotdot", "relative", "symlink"])
def test_vacuum_outdated_keeps_live_files_under_any_spelling_of_the_system_path(tmp_path, spelling, monkeypatch):
    root = tmp_path / "vt" / "indexes"
    vdir = root / "idx" / "v__=0"
    vdir.mkdir(parents=True)
    live = vdir / "part-00000-u_00000.c000.parquet"
    dead = vdir / "part-00000-old_00001.c000.parquet"
    live.write_bytes(b"live")
    dead.write_bytes(b"dead")
    (vdir / "_SUCCESS").write_bytes(b"")
    if spelling == "double_slash":
        sys_path = str(tmp_path) + "/vt//indexes/"
    elif spelling == "trailing_slash":
        sys_path = str(root) + "/"
    elif spelling == "dotdot":
        sys_path = str(tmp_path / "vt" / "x" / ".." / "indexes")
        (tmp_path / "vt" / "x").mkdir()
    elif spelling == "relative":
        monkeypatch.chdir(tmp_path)
        sys_path = "vt/indexes"
    else:
        os.symlink(str(root), str(tmp_path / "lnk"))
        sys_path = str(tmp_path / "lnk")
    s = HyperspaceSession({"spark.hyperspace.system.path": sys_path})
    index_path = LE.PathResolver(s.conf).get_index_path("idx")
    lm = LE.IndexLogManager(index_path)
    e = _entry_over(str(root), "idx", [str(live)])
    assert lm.write_log(1, e)
    lm.create_latest_stable_log(1)
    VacuumOutdatedAction(lm, LE.IndexDataManager(index_path)).run()
    assert live.exists(), "a file the latest log entry references was deleted"
    assert not dead.exists(), "a file no log entry references survived"
    assert (vdir / "_SUCCESS").exists()
    assert lm.get_latest_stable_log().state == "ACTIVE"


def test_lineage_column_lives_in_schema_not_in_included_columns(tmp_path):
    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path / "indexes")})
    index_path = LE.PathResolver(s.conf).get_index_path("idx")
    act = _DataAction(s, LE.IndexLogManager(index_path), LE.IndexDataManager(index_path))
    rel = _rel(tmp_path / "t")
    content = LE.Content.from_leaf_files([(f"file:{tmp_path}/indexes/idx/v__=0/part-00000-x_00000.c000.parquet", 10, 1)], LE.FileIdTracker())
    e = act._build_entry("idx", ["k"], ["v1"], 4, True, rel, content)
    assert e.includedColumns == ["v1"]
    assert [f["name"] for f in e.schema["fields"]] == ["k", "v1", LE.DATA_FILE_NAME_ID]
    assert e.has_lineage_column
    e2 = act._build_entry("idx", ["k"], ["v1", LE.DATA_FILE_NAME_ID], 4, True, rel, content)  # an old-style caller
    assert e2.includedColumns == ["v1"]
    e3 = act._build_entry("idx", ["k"], ["v1"], 4, False, rel, content)
    assert [f["name"] for f in e3.schema["fields"]] == ["k", "v1"] and not e3.has_lineage_column
    # round trip through JSON keeps the split
    back = LE.IndexLogEntry.from_json(e.to_json())
    assert back.includedColumns == ["v1"] and back.has_lineage_column


def _fabricate(tmp_path, session, name, rel, indexed, included, num_buckets=200, lineage="false", update=None, sig_rel=None):
    tracker = LE.FileIdTracker()
    idx_files = [(f"file:{tmp_path}/indexes/{name}/v__=0/part-00000-x_{b:05d}.c000.parquet", 10, 1) for b in range(2)]
    src = LE.Content.from_leaf_files(rel.files, tracker)
    upd = None
    if update:
        app, dele = update
        upd = LE.Update(LE.Content.from_leaf_files(app, tracker) if app else None,
                        LE.Content.from_leaf_files(dele, tracker) if dele else None)
    e = LE.IndexLogEntry(
        name=name, indexedColumns=indexed, includedColumns=included, schema={"type": "struct", "fields": []},
        numBuckets=num_buckets, derived_properties={"lineage": lineage},
        content=LE.Content.from_leaf_files(idx_files, LE.FileIdTracker()),
        relations=[LE.Relation(rel.root_paths, src, {"type": "struct", "fields": []}, "parquet", {}, upd)],
        signatures=[LE.Signature(LE.INDEX_SIGNATURE_PROVIDER, R.index_signature(sig_rel or rel))], state="ACTIVE", id=1)
    lm = LE.IndexLogManager(os.path.join(LE.PathResolver(session.conf).system_path, name))
    lm.write_log(1, e)
    lm.create_latest_stable_log(1)
    return e


def test_quick_refreshed_index_is_used_without_the_hybridscan_conf(tmp_path):
    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path / "indexes")}).enableHyperspace()
    old = _rel(tmp_path / "t")
    now = _rel(tmp_path / "t", files=(("f1", 100, 1), ("f2", 100, 2), ("f3", 10, 3)))
    # what RefreshQuickAction leaves behind: source content = old files, Update.appended = f3, signature of the new listing
    _fabricate(tmp_path, s, "q", old, ["k"], ["v1"], update=([(f"file:{tmp_path}/t/f3", 10, 3)], None), sig_rel=now)
    assert not s.conf.hybrid_scan_enabled
    plan = DataFrame(s, now).filter(col("k") >= 1).select("k", "v1").explain()
    assert "Name: q" in plan and "hybridScan(appended=1" in plan
    # recorded deletes need lineage; without it the index is not applicable
    gone = _rel(tmp_path / "t2", files=(("f1", 100, 1),))
    full = _rel(tmp_path / "t2")
    _fabricate(tmp_path, s, "d_nolineage", full, ["v2"], ["k"], update=(None, [(f"file:{tmp_path}/t2/f2", 100, 2)]), sig_rel=gone)
    assert "GpuSourceScan" in DataFrame(s, gone).filter(col("v2") >= 1).select("k").explain()
    _fabricate(tmp_path, s, "d_lineage", full, ["v1"], ["k"], lineage="true",
               update=(None, [(f"file:{tmp_path}/t2/f2", 100, 2)]), sig_rel=gone)
    plan = DataFrame(s, gone).filter(col("v1") >= 1).select("k").explain()
    assert "Name: d_lineage" in plan and "deletedIds=[1]" in plan


def test_join_skips_an_index_whose_source_lost_files(tmp_path):
    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path / "indexes")}).enableHyperspace()
    s.conf.set("spark.hyperspace.index.hybridscan.enabled", True)
    s.conf.set("spark.hyperspace.index.hybridscan.maxDeletedRatio", 0.9)
    lfull = _rel(tmp_path / "l", ("k", "a"))
    lnow = _rel(tmp_path / "l", ("k", "a"), files=(("f1", 100, 1),))  # f2 deleted
    rrel = _rel(tmp_path / "r", ("k", "b"))
    _fabricate(tmp_path, s, "lidx", lfull, ["k"], ["a"], lineage="true")
    _fabricate(tmp_path, s, "ridx", rrel, ["k"], ["b"])
    # the filter rule may use the index (lineage NOT-IN filter) ...
    assert "Name: lidx" in DataFrame(s, lnow).filter(col("k") >= 0).select("k", "a").explain()
    # ... the join must not, and planning must not raise
    plan = DataFrame(s, lnow).join(DataFrame(s, rrel), on="k").select("a", "b").explain()
    assert "Name: lidx" not in plan


def test_comparison_operators_round_fractional_literals_correctly():
    b = lambda p: p.bounds["k"]  # noqa: E731
    assert b(col("k") < 1.5) == (None, 1)
    assert b(col("k") < 2) == (None, 1)
    assert b(col("k") <= 1.5) == (None, 1)
    assert b(col("k") > 1.5) == (2, None)
    assert b(col("k") > 1) == (2, None)
    assert b(col("k") >= 1.5) == (2, None)
    assert b(col("k") >= -1.5) == (-1, None)
    assert b(col("k") < -1.5) == (None, -2)
    assert b(col("k") > -1.5) == (-1, None)
    assert b(col("k") <= -1.5) == (None, -2)
    assert b(col("k").between(0.5, 2.5)) == (1, 2)
    lo, hi = b(col("k") == 1.5)
    assert lo > hi  # empty
    assert b(col("k") == 3) == (3, 3)


def test_string_literals_give_byte_bounds_in_utf8_order():
    """Predicates on string columns (the reference's filter-rule tests filter on `c3 == "facebook"`): inclusive byte bounds in
    UTF8String order, which is what hs_filter_scan's lo_bytes / hi_bytes take."""
    b = lambda p: p.bounds["q"]  # noqa: E731
    assert b(col("q") == "facebook") == (b"facebook", b"facebook")
    assert b(col("q") >= "é") == ("é".encode("utf-8"), None)
    assert b(col("q") <= b"\xff\x00") == (None, b"\xff\x00")
    assert b(col("q") > "abc") == (b"abc\x00", None)  # the smallest value above "abc"
    assert b(col("q").between("a", "b")) == (b"a", b"b")
    assert b((col("q") >= "b") & (col("q") <= "y") & (col("q") >= "c")) == (b"c", b"y")  # conjunction keeps the tighter bound
    with pytest.raises(ValueError):
        col("q") < "x"  # noqa: B015  (no inclusive form; the message says to use <= or between)


def test_join_rule_needs_equal_key_types(tmp_path):
    """Spark casts one side of `int = long` (or `string = long`), and a condition over a Cast is not what JoinIndexRule accepts
    (index/covering/JoinIndexRule.scala:143-163); physically hashInt / hashLong / hashUnsafeBytes bucket equal values
    differently.  Same types: both indexes are used."""
    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path / "indexes")})
    s.enableHyperspace()
    lrel = _rel(tmp_path / "l", ("k", "a"))
    rrel = _rel(tmp_path / "r", ("k", "b"))
    _fabricate(tmp_path, s, "lidx", lrel, ["k"], ["a"])
    _fabricate(tmp_path, s, "ridx", rrel, ["k"], ["b"])
    plan = DataFrame(s, lrel).join(DataFrame(s, rrel), on="k").select("a", "b").explain()
    assert "Name: lidx" in plan and "Name: ridx" in plan
    for other in ("integer", "string"):
        rrel.schema[0] = ("k", other)
        plan = DataFrame(s, lrel).join(DataFrame(s, rrel), on="k").select("a", "b").explain()
        assert "Name: lidx" not in plan and "Name: ridx" not in plan, other

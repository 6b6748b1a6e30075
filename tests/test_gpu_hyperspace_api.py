"""GPU end-to-end tests through the Hyperspace API surface (createIndex / refreshIndex / optimizeIndex / queries).

Modelled on the reference's hot-path suites: T/index/CreateIndexTest.scala, T/index/IndexManagerTest.scala (versions,
file names, refresh full / incremental, optimize -> one file per bucket), T/index/RefreshIndexTest.scala (append + delete
with lineage), T/index/E2EHyperspaceRulesTest.scala `verifyIndexUsage` (answers identical with Hyperspace on and off, plan
uses the index files) and T/index/HybridScanSuite.scala (appended / deleted source files without refresh).
"""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _write(dirpath, name, cols):
    os.makedirs(dirpath, exist_ok=True)
    pq.write_table(pa.table(cols), os.path.join(dirpath, name), compression="snappy")  # Spark's default codec


def _table(first, n):
    c = O.synthetic_table(first, n, 3)
    c["k"] = (c["k"] % 5000).astype(np.int64)  # plenty of duplicate keys and a bounded range for predicates
    return c


def _rows(res, cols):
    return np.sort(np.rec.fromarrays([np.asarray(res[c]).view(np.int64) if np.asarray(res[c]).dtype.itemsize == 8
                                      else np.asarray(res[c]) for c in cols]))


@pytest.fixture()
def env(tmp_path):
    from hyperspace_b200.hyperspace import Hyperspace
    from hyperspace_b200.session import HyperspaceSession

    s = HyperspaceSession({"spark.hyperspace.system.path": str(tmp_path / "indexes"), "spark.hyperspace.index.numBuckets": "8"})
    yield s, Hyperspace(s), tmp_path
    s.stop()


def _index_dir(tmp_path, name, v):
    return tmp_path / "indexes" / name / f"v__={v}"


def test_create_index_layout_and_log(env):
    from hyperspace_b200 import log_entry as LE
    from hyperspace_b200.index_config import IndexConfig

    s, hs, tmp = env
    _write(tmp / "t", "part-0.parquet", _table(0, 20_000))
    _write(tmp / "t", "part-1.parquet", _table(20_000, 20_000))
    df = s.read.parquet(str(tmp / "t"))
    hs.createIndex(df, IndexConfig("idx", ["k"], ["v1", "v2"]))
    files = sorted(os.listdir(_index_dir(tmp, "idx", 0)))
    assert len(files) == 8 and all(f.startswith("part-0") and f.endswith(".parquet") for f in files)
    lm = LE.IndexLogManager(str(tmp / "indexes" / "idx"))
    assert lm.get_log(0).state == "CREATING" and lm.get_log(1).state == "ACTIVE" and lm.get_latest_stable_log().id == 1
    e = lm.get_log(1)
    assert e.numBuckets == 8 and e.indexedColumns == ["k"] and e.includedColumns == ["v1", "v2"]
    assert sorted(os.path.basename(f) for f in e.index_files) == files
    assert len(e.source_file_infos) == 2 and [f.id for f in e.source_file_infos] == [0, 1]
    assert [f["type"] for f in e.schema["fields"]] == ["long", "long", "double"]
    # every file: bucket id from the name == Spark hash of every key; sorted (DataFrameWriterExtensionsTest.scala:93-158)
    total = 0
    for f in files:
        b = int(f.rsplit("_", 1)[1].split(".")[0])
        k = pq.ParquetFile(str(_index_dir(tmp, "idx", 0) / f)).read().column("k").to_numpy()
        assert np.all(O.np_bucket_ids([k], 8) == b) and np.all(k[:-1] <= k[1:])
        total += len(k)
    assert total == 40_000
    assert [i["name"] for i in hs.indexes()] == ["idx"]
    with pytest.raises(LE.HyperspaceException):  # CreateIndexTest: same name twice
        hs.createIndex(df, IndexConfig("idx", ["k"], ["v1"]))
    with pytest.raises(LE.HyperspaceException):  # unknown column
        hs.createIndex(df, IndexConfig("idx2", ["nope"], ["v1"]))


def test_filter_and_join_answers_match_with_and_without_index(env):
    from hyperspace_b200.index_config import IndexConfig
    from hyperspace_b200.session import col

    s, hs, tmp = env
    L, R = _table(0, 30_000), _table(100_000, 25_000)
    R = {"k": R["k"], "w": R["v1"]}
    _write(tmp / "l", "a.parquet", L)
    _write(tmp / "r", "a.parquet", R)
    dl, dr = s.read.parquet(str(tmp / "l")), s.read.parquet(str(tmp / "r"))
    hs.createIndex(dl, IndexConfig("lidx", ["k"], ["v1", "v2"]))
    hs.createIndex(dr, IndexConfig("ridx", ["k"], ["w"]))

    q = dl.filter(col("k").between(100, 300)).select("k", "v2")
    s.disableHyperspace()
    assert "GpuSourceScan" in q.explain()
    base = q.collect()
    s.enableHyperspace()
    assert "Name: lidx" in q.explain()
    got = q.collect()
    assert len(got["k"]) == int(((L["k"] >= 100) & (L["k"] <= 300)).sum())
    assert np.array_equal(_rows(got, ["k", "v2"]), _rows(base, ["k", "v2"]))  # verifyIndexUsage: sorted rows identical

    j = dl.join(dr, on="k").select("v1", "w")
    s.disableHyperspace()
    assert "GpuShuffle" in j.explain()
    jb = j.collect()
    s.enableHyperspace()
    plan = j.explain()
    assert "Name: lidx" in plan and "Name: ridx" in plan and "exchange=none" in plan
    jg = j.collect()
    # oracle answer
    order = np.argsort(R["k"], kind="stable")
    lo, hi = np.searchsorted(R["k"][order], L["k"], "left"), np.searchsorted(R["k"][order], L["k"], "right")
    want = sum(int(h - l) for l, h in zip(lo, hi))
    assert len(jg["v1"]) == want == len(jb["v1"])
    assert np.array_equal(_rows(jg, ["v1", "w"]), _rows(jb, ["v1", "w"]))


def test_refresh_full_incremental_quick_and_optimize(env):
    from hyperspace_b200 import log_entry as LE
    from hyperspace_b200.index_config import IndexConfig
    from hyperspace_b200.session import col

    s, hs, tmp = env
    s.conf.set("spark.hyperspace.index.lineage.enabled", True)
    _write(tmp / "t", "f0.parquet", _table(0, 10_000))
    _write(tmp / "t", "f1.parquet", _table(10_000, 10_000))
    hs.createIndex(s.read.parquet(str(tmp / "t")), IndexConfig("idx", ["k"], ["v1"]))
    lm = LE.IndexLogManager(str(tmp / "indexes" / "idx"))
    assert lm.get_latest_stable_log().has_lineage_column
    t0 = pq.ParquetFile(str(_index_dir(tmp, "idx", 0) / sorted(os.listdir(_index_dir(tmp, "idx", 0)))[0])).read()
    assert t0.column_names == ["k", "v1", "_data_file_id"]  # CreateIndexTest.scala:157-243 lineage column present
    hs.refreshIndex("idx", "full")  # no source change -> recorded no-op (RefreshAction.scala:53-59)
    assert lm.get_latest_id() == 1

    # ---- incremental: one appended file, one deleted file ---------------------------------------------------
    _write(tmp / "t", "f2.parquet", _table(20_000, 5_000))
    os.remove(tmp / "t" / "f0.parquet")
    hs.refreshIndex("idx", "incremental")
    e = lm.get_latest_stable_log()
    assert e.state == "ACTIVE" and e.id == 3 and e.index_version_dirs() == [1]
    live = np.concatenate([_table(10_000, 10_000)["k"], _table(20_000, 5_000)["k"]])
    got = np.concatenate([pq.ParquetFile(LE.from_uri(f)).read().column("k").to_numpy() for f in e.index_files])
    assert np.array_equal(np.sort(got), np.sort(live))  # RefreshIndexTest.scala:95-106: deleted rows gone, appended rows in
    assert sorted(f.name.rsplit("/", 1)[1] for f in e.source_file_infos) == ["f1.parquet", "f2.parquet"]
    # after append-only incremental refreshes a bucket holds several files ...
    _write(tmp / "t", "f3.parquet", _table(25_000, 5_000))
    hs.refreshIndex("idx", "incremental")
    e = lm.get_latest_stable_log()
    assert e.index_version_dirs() == [1, 2]  # Merge mode: old files U new files (RefreshIncrementalAction.scala:115-128)
    per_bucket = {}
    for f in e.index_files:
        per_bucket.setdefault(int(f.rsplit("_", 1)[1].split(".")[0]), []).append(f)
    assert max(len(v) for v in per_bucket.values()) > 1
    # ... queries still answer correctly over multi-file buckets
    s.enableHyperspace()
    df = s.read.parquet(str(tmp / "t"))
    q = df.filter(col("k") <= 50).select("k", "v1")
    assert "Name: idx" in q.explain()
    cur = np.concatenate([_table(a, b)["k"] for a, b in ((10_000, 10_000), (20_000, 5_000), (25_000, 5_000))])
    assert len(q.collect()["k"]) == int((cur <= 50).sum())
    # ---- optimize: one file per bucket again (IndexManagerTest.scala:473-475) ---------------------------------
    hs.optimizeIndex("idx", "full")
    e = lm.get_latest_stable_log()
    buckets = [int(f.rsplit("_", 1)[1].split(".")[0]) for f in e.index_files]
    assert len(buckets) == len(set(buckets))
    got = np.concatenate([pq.ParquetFile(LE.from_uri(f)).read().column("k").to_numpy() for f in e.index_files])
    assert np.array_equal(np.sort(got), np.sort(cur))
    assert len(q.collect()["k"]) == int((cur <= 50).sum())
    # ---- quick refresh + Hybrid Scan: metadata only, appended file scanned at query time ------------------------
    _write(tmp / "t", "f4.parquet", _table(30_000, 2_000))
    n_before = lm.get_latest_id()
    hs.refreshIndex("idx", "quick")
    e = lm.get_latest_stable_log()
    assert lm.get_latest_id() == n_before + 2 and [f.name.rsplit("/", 1)[1] for f in e.appended_files] == ["f4.parquet"]
    s.conf.set("spark.hyperspace.index.hybridscan.enabled", True)
    df = s.read.parquet(str(tmp / "t"))
    q = df.filter(col("k") <= 50).select("k", "v1")
    assert "hybridScan(appended=1" in q.explain()
    cur2 = np.concatenate([cur, _table(30_000, 2_000)["k"]])
    assert len(q.collect()["k"]) == int((cur2 <= 50).sum())
    # ---- delete / restore / vacuum ------------------------------------------------------------------------
    hs.deleteIndex("idx")
    assert "GpuSourceScan" in q.explain()
    hs.restoreIndex("idx")
    hs.vacuumIndex("idx")  # ACTIVE -> VacuumOutdated: unreferenced versions go away
    e = lm.get_latest_stable_log()
    assert sorted(int(d.split("=")[1]) for d in os.listdir(tmp / "indexes" / "idx") if d.startswith("v__=")) == e.index_version_dirs()
    hs.deleteIndex("idx")
    hs.vacuumIndex("idx")
    assert hs.indexes() == [] and not any(d.startswith("v__=") for d in os.listdir(tmp / "indexes" / "idx"))


def test_hybrid_scan_with_deleted_source_file(env):
    from hyperspace_b200.index_config import IndexConfig
    from hyperspace_b200.session import col

    s, hs, tmp = env
    s.conf.set("spark.hyperspace.index.lineage.enabled", True)
    for i in range(6):
        _write(tmp / "t", f"f{i}.parquet", _table(i * 5_000, 5_000))
    hs.createIndex(s.read.parquet(str(tmp / "t")), IndexConfig("idx", ["k"], ["v1"]))
    os.remove(tmp / "t" / "f5.parquet")  # 1/6 of the bytes < maxDeletedRatio 0.2
    s.enableHyperspace()
    s.conf.set("spark.hyperspace.index.hybridscan.enabled", True)
    q = s.read.parquet(str(tmp / "t")).filter(col("k") <= 100).select("k", "v1")
    assert "deletedIds=[5]" in q.explain()
    cur = np.concatenate([_table(i * 5_000, 5_000)["k"] for i in range(5)])
    assert len(q.collect()["k"]) == int((cur <= 100).sum())  # HybridScanSuite.scala:378-441 checkAnswer


def test_hybrid_scan_join_with_appended_files(env):
    """JoinIndexRule under Hybrid Scan: appended source files are bucketed on the fly and merged per bucket with the index
    (BucketUnion, S/index/covering/CoveringIndexRuleUtils.scala:256-284; T/index/HybridScanSuite.scala)."""
    from hyperspace_b200.index_config import IndexConfig

    s, hs, tmp = env
    for i in range(5):
        _write(tmp / "l", f"f{i}.parquet", _table(i * 4_000, 4_000))
    R = _table(50_000, 15_000)
    _write(tmp / "r", "a.parquet", {"k": R["k"], "w": R["v1"]})
    hs.createIndex(s.read.parquet(str(tmp / "l")), IndexConfig("lidx", ["k"], ["v1"]))
    hs.createIndex(s.read.parquet(str(tmp / "r")), IndexConfig("ridx", ["k"], ["w"]))
    _write(tmp / "l", "f5.parquet", _table(20_000, 3_000))  # appended after the index was built: 3/23 of the bytes < 0.3
    s.enableHyperspace()
    dl, dr = s.read.parquet(str(tmp / "l")), s.read.parquet(str(tmp / "r"))
    j = dl.join(dr, on="k").select("v1", "w")
    assert "Name: lidx" not in j.explain()  # stale signature: the left index is not used without Hybrid Scan
    s.conf.set("spark.hyperspace.index.hybridscan.enabled", True)
    plan = j.explain()
    assert "Name: lidx" in plan and "Name: ridx" in plan
    got = j.collect()
    s.disableHyperspace()
    base = j.collect()
    assert len(got["v1"]) == len(base["v1"]) > 0
    assert np.array_equal(_rows(got, ["v1", "w"]), _rows(base, ["v1", "w"]))


def test_string_index_through_the_api_like_the_reference_examples(env):
    """The reference's canonical flow (its README and T/index/E2EHyperspaceRulesTest.scala): index a string column, filter on
    it, the plan uses the index and the answer does not change.  Data: T/SampleData.scala:25-35."""
    from hyperspace_b200.index_config import IndexConfig
    from hyperspace_b200.session import col

    s, hs, tmp = env
    sample = [("2017-09-03", "810a20a2baa24ff3ad493bfbf064569a", "donde", 2, 1000),
              ("2017-09-03", "fd093f8a05604515957083e70cb3dceb", "facebook", 1, 3000),
              ("2017-09-03", "af3ed6a197a8447cba8bc8ea21fad208", "facebook", 1, 3000),
              ("2017-09-03", "975134eca06c4711a0406d0464cbe7d6", "facebook", 1, 4000),
              ("2018-09-03", "e90a6028e15b4f4593eef557daf5166d", "ibraco", 2, 3000),
              ("2018-09-03", "576ed96b0d5340aa98a47de15c9f87ce", "facebook", 2, 3000),
              ("2018-09-03", "50d690516ca641438166049a6303650c", "ibraco", 2, 1000),
              ("2019-10-03", "380786e6495d4cd8a5dd4cc8d3d12917", "facebook", 2, 3000),
              ("2019-10-03", "ff60e4838b92421eafc3e6ee59a9e9f1", "miperro", 2, 2000),
              ("2019-10-03", "187696fe0a6a40cc9516bc6e47c70bc1", "facebook", 4, 3000)]
    c = list(zip(*sample))
    cols = {"Date": pa.array(c[0]), "RGUID": pa.array(c[1]), "Query": pa.array(c[2]), "imprs": pa.array(c[3], pa.int32()),
            "clicks": pa.array(c[4], pa.int32())}
    os.makedirs(tmp / "sample", exist_ok=True)
    pq.write_table(pa.table(cols), str(tmp / "sample" / "part-0.parquet"), compression="snappy")
    df = s.read.parquet(str(tmp / "sample"))
    hs.createIndex(df, IndexConfig("qidx", ["Query"], ["RGUID", "clicks"]))
    q = df.filter(col("Query") == "facebook").select("RGUID", "clicks", "Query")
    s.disableHyperspace()
    assert "GpuSourceScan" in q.explain()
    base = q.collect()
    s.enableHyperspace()
    assert "Name: qidx" in q.explain()
    got = q.collect()
    want = sorted((r[1], r[4], r[2]) for r in sample if r[2] == "facebook")
    assert sorted(zip(got["RGUID"], (int(x) for x in got["clicks"]), got["Query"])) == want
    assert sorted(zip(base["RGUID"], (int(x) for x in base["clicks"]), base["Query"])) == want
    rng_q = df.filter(col("Query").between("e", "j")).select("Query", "clicks")
    got = rng_q.collect()
    assert sorted(zip(got["Query"], (int(x) for x in got["clicks"]))) == sorted((r[2], r[4]) for r in sample if "e" <= r[2] <= "j")
    # the reference's E2E join tests join the sample data with itself on the string column (leftDf("c3") === rightDf("c3"))
    os.makedirs(tmp / "sample2", exist_ok=True)
    pq.write_table(pa.table({"Query": cols["Query"], "shown": cols["imprs"]}), str(tmp / "sample2" / "part-0.parquet"),
                   compression="snappy")
    dr = s.read.parquet(str(tmp / "sample2"))
    hs.createIndex(dr, IndexConfig("qidx2", ["Query"], ["shown"]))
    j = df.join(dr, on="Query").select("RGUID", "shown")
    s.disableHyperspace()
    jb = j.collect()
    s.enableHyperspace()
    plan = j.explain()
    assert "Name: qidx" in plan and "Name: qidx2" in plan and "exchange=none" in plan
    jg = j.collect()
    want = sorted((a[1], b[3]) for a in sample for b in sample if a[2] == b[2])
    assert sorted(zip(jg["RGUID"], (int(x) for x in jg["shown"]))) == want
    assert sorted(zip(jb["RGUID"], (int(x) for x in jb["shown"]))) == want

"""String (BYTE_ARRAY) keys and included columns on the write path, through the C ABI, bit-exact against the CPU oracle.

Restates the reference's own write-path scenarios, all of which bucket on a string column
(T/index/DataFrameWriterExtensionsTest.scala:160-180 over T/SampleData.scala:25-35): one bucket column `Query`; two bucket
columns `clicks, Query`; Append into the same directory -- each checked the way testInternal does it (:93-158): the
bucket id of every row == Spark's HashPartitioning (pmod(murmur3, n), strings hashed with hashUnsafeBytes), every file
sorted on the bucket columns (UTF8String byte order), the rows of all files == the DataFrame's rows.
"""
import os

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

SAMPLE = [
    ("2017-09-03", "810a20a2baa24ff3ad493bfbf064569a", "donde", 2, 1000),
    ("2017-09-03", "fd093f8a05604515957083e70cb3dceb", "facebook", 1, 3000),
    ("2017-09-03", "af3ed6a197a8447cba8bc8ea21fad208", "facebook", 1, 3000),
    ("2017-09-03", "975134eca06c4711a0406d0464cbe7d6", "facebook", 1, 4000),
    ("2018-09-03", "e90a6028e15b4f4593eef557daf5166d", "ibraco", 2, 3000),
    ("2018-09-03", "576ed96b0d5340aa98a47de15c9f87ce", "facebook", 2, 3000),
    ("2018-09-03", "50d690516ca641438166049a6303650c", "ibraco", 2, 1000),
    ("2019-10-03", "380786e6495d4cd8a5dd4cc8d3d12917", "facebook", 2, 3000),
    ("2019-10-03", "ff60e4838b92421eafc3e6ee59a9e9f1", "miperro", 2, 2000),
    ("2019-10-03", "187696fe0a6a40cc9516bc6e47c70bc1", "facebook", 4, 3000),
]
COLUMNS = ["Date", "RGUID", "Query", "imprs", "clicks"]


@pytest.fixture(scope="module")
def ctx():
    from hyperspace_b200 import _native

    c = _native.Context(0)
    yield c
    c.close()


def _sample_table():
    cols = list(zip(*SAMPLE))
    return pa.table({"Date": pa.array(cols[0]), "RGUID": pa.array(cols[1]), "Query": pa.array(cols[2]),
                     "imprs": pa.array(cols[3], pa.int32()), "clicks": pa.array(cols[4], pa.int32())})


def _np_cols(table):
    out, valid = {}, {}
    for name in table.column_names:
        col = table.column(name).combine_chunks()
        if pa.types.is_string(col.type) or pa.types.is_binary(col.type) or pa.types.is_large_string(col.type):
            out[name] = np.array([("" if v is None else v) for v in col.to_pylist()], dtype=object)
        else:
            out[name] = col.fill_null(0).to_numpy(zero_copy_only=False)
        if col.null_count:
            valid[name] = np.asarray(col.is_valid())
    return out, valid


def _write(table, path, **kw):
    pq.write_table(table, path, compression=kw.pop("compression", "NONE"), **kw)
    return path


def _check_index(res, read_file, table, indexed, included, nb, times=1):
    """testInternal of the reference: bucket ids, per-file order and row multiset; plus exact row order against the oracle."""
    cols, valid = _np_cols(table)
    perm, offs, order = O.index_rows(cols, indexed, included, nb, valids=valid or None)
    seen_rows = 0
    by_bucket = {}
    for i, f in enumerate(res.files):
        by_bucket.setdefault(f.bucket, []).append(read_file(i, f))
    for b in range(nb):
        lo, hi = int(offs[b]), int(offs[b + 1])
        if hi == lo:
            assert b not in by_bucket
            continue
        assert len(by_bucket[b]) == times
        for t in by_bucket[b]:
            assert t.column_names == order
            assert t.num_rows == hi - lo
            for name in order:
                got = t.column(name).combine_chunks()
                idx = perm[lo:hi]
                if cols[name].dtype == object:
                    want = [None if (name in valid and not valid[name][j]) else cols[name][j] for j in idx]
                    assert got.to_pylist() == want, (b, name)
                else:
                    want_valid = valid[name][idx] if name in valid else np.ones(hi - lo, bool)
                    assert np.array_equal(np.asarray(got.is_valid()), want_valid), (b, name)
                    g = got.fill_null(0).to_numpy(zero_copy_only=False)
                    assert np.array_equal(g[want_valid], cols[name][idx][want_valid]), (b, name)
            seen_rows += t.num_rows
    assert seen_rows == table.num_rows * times


@pytest.mark.parametrize("bucket_by", [["Query"], ["clicks", "Query"]])
@pytest.mark.parametrize("dictionary", [True, False])
def test_save_with_buckets_on_sample_data(ctx, tmp_path, bucket_by, dictionary):
    from hyperspace_b200 import _native as N

    table = _sample_table()
    src = _write(table, str(tmp_path / "src.parquet"), use_dictionary=dictionary)
    included = [c for c in COLUMNS if c not in bucket_by]
    out_dir = str(tmp_path / "v__=0")
    res, st = ctx.create_index([N.FileImage(path=src)], bucket_by, included, 3, out_dir=out_dir, output=N.HS_OUT_FILES, job_uuid="s")
    assert st["rows_out"] == 10
    read = lambda i, f: pq.ParquetFile(os.path.join(out_dir, f.name)).read()  # noqa: E731  (no hive column from 'v__=0')
    _check_index(res, read, table, bucket_by, included, 3)
    res.free()
    # Append mode: a second write into the same directory doubles every bucket's files (DataFrameWriterExtensionsTest.scala:173-180)
    res2, _ = ctx.create_index([N.FileImage(path=src)], bucket_by, included, 3, out_dir=out_dir, output=N.HS_OUT_FILES,
                               job_uuid="t", save_mode=N.HS_SAVE_APPEND)
    names = sorted(n for n in os.listdir(out_dir) if not n.startswith((".", "_")))
    assert len(names) == 2 * len(res2.files)
    for n in names:
        t = pq.ParquetFile(os.path.join(out_dir, n)).read()
        keys = list(zip(*[t.column(c).to_pylist() for c in bucket_by]))
        enc = [tuple(x.encode() if isinstance(x, str) else x for x in k) for k in keys]
        assert enc == sorted(enc)
    res2.free()


def _random_strings(rng, n, max_len=24, alphabet=b"abcdefghijklmnopqrstuvwxyz0123456789-_/ \xc3\xa9"):
    lens = rng.integers(0, max_len + 1, size=n)
    raw = rng.integers(0, len(alphabet), size=int(lens.sum()))
    blob = bytes(alphabet[i] for i in raw)
    out, p = [], 0
    for ln in lens:
        out.append(blob[p:p + ln])
        p += ln
    return out


@pytest.mark.parametrize("n,nb", [(1_000_000, 200), (50_000, 7)])
def test_random_string_keys_match_the_oracle(ctx, n, nb):
    """Binary keys of 0..24 bytes (shared prefixes, empty strings, bytes >= 0x80 -- Spark mixes tail bytes as SIGNED ints),
    nullable string and numeric included columns, several files, small pages: bucket per row, order incl. ties, payload."""
    from hyperspace_b200 import _native as N

    rng = np.random.default_rng(n + nb)
    keys = _random_strings(rng, n)
    for i in range(0, n, 97):  # runs of equal and of prefix-related keys
        keys[i] = keys[(i * 7) % n][:5]
    words = [b"alpha", b"beta", b"gamma", b"", b"delta-delta-delta"]
    s = [words[i] for i in rng.integers(0, len(words), size=n)]
    smask = rng.random(n) < 0.1
    v = rng.integers(-10**9, 10**9, size=n, dtype=np.int64)
    table = pa.table({"k": pa.array(keys, pa.binary()), "s": pa.array(s, pa.binary(), mask=smask), "v": pa.array(v)})
    images = []
    per = n // 3
    for i in range(3):
        sink = pa.BufferOutputStream()
        part = table.slice(i * per, per if i < 2 else n - 2 * per)
        pq.write_table(part, sink, compression="SNAPPY" if i == 1 else "NONE", use_dictionary=["s"], data_page_size=64 << 10,
                       data_page_version="2.0" if i == 2 else "1.0")
        images.append(N.FileImage(data=sink.getvalue().to_pybytes()))
    res, st = ctx.create_index(images, ["k"], ["s", "v"], nb, output=N.HS_OUT_HOST, job_uuid="r")
    assert st["rows_out"] == n
    read = lambda i, f: pq.ParquetFile(pa.BufferReader(res.host_bytes(i))).read()  # noqa: E731
    _check_index(res, read, table, ["k"], ["s", "v"], nb)
    # whole-index verification agrees with the same checks done on the source (one "bucket": row multiset only)
    rep = ctx.verify_index(res.as_sources(), [f.bucket for f in res.files], ["k"], ["s", "v"], nb)
    src = ctx.verify_index(images, [0, 0, 0], ["k"], ["s", "v"], 1)
    assert rep["rows"] == n and rep["bucket_mismatches"] == 0 and rep["order_violations"] == 0
    assert rep["row_checksum"] == src["row_checksum"] and rep["column_checksum"] == src["column_checksum"]
    res.free()


def test_nullable_string_key_and_long_values(ctx):
    from hyperspace_b200 import _native as N

    rng = np.random.default_rng(4)
    n = 20_000
    keys = [("key-%05d-" % int(x)) * int(1 + x % 9) for x in rng.integers(0, 3000, size=n)]  # up to ~100 bytes, many ties
    kmask = rng.random(n) < 0.05
    v = np.arange(n, dtype=np.int32)
    table = pa.table({"k": pa.array(keys, pa.string(), mask=kmask), "v": pa.array(v)})
    sink = pa.BufferOutputStream()
    pq.write_table(table, sink, compression="NONE")
    img = [N.FileImage(data=sink.getvalue().to_pybytes())]
    res, _ = ctx.create_index(img, ["k"], ["v"], 11, output=N.HS_OUT_HOST)
    read = lambda i, f: pq.ParquetFile(pa.BufferReader(res.host_bytes(i))).read()  # noqa: E731
    _check_index(res, read, table, ["k"], ["v"], 11)
    res.free()
    # a value longer than 65535 bytes is refused, not truncated
    big = pa.table({"k": pa.array(["x" * 70_000, "y"]), "v": pa.array([1, 2], pa.int32())})
    sink = pa.BufferOutputStream()
    pq.write_table(big, sink, compression="NONE", use_dictionary=False)
    with pytest.raises(N.HyperspaceGpuError) as e:
        ctx.create_index([N.FileImage(data=sink.getvalue().to_pybytes())], ["k"], ["v"], 3, output=N.HS_OUT_HOST)
    assert e.value.code == N.HS_EUNSUPPORTED


# ---------------------------------------------------------------------------------------------------------------------
# read side: FilterIndexRule's scan over an index on a string column (index/covering/FilterIndexRule.scala:135-149)
# ---------------------------------------------------------------------------------------------------------------------

def _scan_rows(batch, names):
    cols = {n: d for n, d, _ in batch.columns}
    valid = {n: v for n, _, v in batch.columns}
    out = []
    for r in range(batch.num_rows):
        out.append(tuple(None if (valid[n] is not None and not valid[n][r]) else
                         (bytes(cols[n][r]) if cols[n].dtype == object else cols[n][r].item()) for n in names))
    return sorted(out, key=repr)


def test_filter_scan_on_a_string_key_like_the_filter_rule_tests(ctx, tmp_path):
    """`SELECT ... WHERE Query = 'facebook'` over an index on Query (the predicate of the reference's filter-rule tests,
    T/index/E2EHyperspaceRulesTest.scala; data T/SampleData.scala:25-35): two binary searches per sorted index file, string
    and integer columns projected; the same predicate as a full scan over the source (Hybrid Scan's appended files)."""
    from hyperspace_b200 import _native as N

    table = _sample_table()
    src = _write(table, str(tmp_path / "src.parquet"))
    res, _ = ctx.create_index([N.FileImage(path=src)], ["Query"], ["RGUID", "imprs", "clicks"], 3, output=N.HS_OUT_HOST)
    idx = res.as_sources()
    rows = [(q.encode(), g.encode(), i, c) for (_, g, q, i, c) in SAMPLE]
    names = ["Query", "RGUID", "imprs", "clicks"]
    for lo, hi in (("facebook", "facebook"), ("donde", "facebook"), ("g", None), (None, "e"), ("zzz", "zzzz"), ("", "\xff")):
        want = sorted([r for r in rows if (lo is None or r[0] >= lo.encode()) and (hi is None or r[0] <= hi.encode())], key=repr)
        b, st = ctx.filter_scan(idx, "Query", names, lo=lo, hi=hi, sorted_on_key=True)
        assert _scan_rows(b, names) == want, (lo, hi)
        b.free()
        b, _ = ctx.filter_scan([N.FileImage(path=src)], "Query", names, lo=lo, hi=hi, sorted_on_key=False)
        assert _scan_rows(b, names) == want, (lo, hi)
        b.free()
    res.free()


def test_filter_scan_on_random_binary_keys(ctx):
    """300 K binary keys (shared prefixes, empty values, bytes >= 0x80), nullable string payload: ranges and equalities
    against numpy on the source rows."""
    from hyperspace_b200 import _native as N

    rng = np.random.default_rng(12)
    n, nb = 300_000, 16
    keys = _random_strings(rng, n, max_len=12)
    words = [b"alpha", b"beta", b"", b"delta-delta"]
    s = [words[i] for i in rng.integers(0, len(words), size=n)]
    smask = rng.random(n) < 0.2
    v = rng.integers(0, 10**6, size=n, dtype=np.int64)
    table = pa.table({"k": pa.array(keys, pa.binary()), "s": pa.array(s, pa.binary(), mask=smask), "v": pa.array(v)})
    sink = pa.BufferOutputStream()
    pq.write_table(table, sink, compression="SNAPPY", data_page_size=32 << 10)
    res, _ = ctx.create_index([N.FileImage(data=sink.getvalue().to_pybytes())], ["k"], ["s", "v"], nb, output=N.HS_OUT_HOST)
    idx = res.as_sources()
    karr = np.array(keys, dtype=object)
    probes = [(keys[5], keys[5]), (b"a", b"b"), (b"", b""), (b"m", None), (None, b"0"), (keys[77][:2], keys[77][:2] + b"\xff\xff")]
    for lo, hi in probes:
        sel = np.ones(n, bool)
        if lo is not None:
            sel &= np.array([k >= lo for k in karr])
        if hi is not None:
            sel &= np.array([k <= hi for k in karr])
        want = sorted([(keys[i], None if smask[i] else s[i], int(v[i])) for i in np.flatnonzero(sel)], key=repr)
        b, _ = ctx.filter_scan(idx, "k", ["k", "s", "v"], lo=lo, hi=hi, sorted_on_key=True)
        assert b.num_rows == len(want)
        assert _scan_rows(b, ["k", "s", "v"]) == want, (lo, hi)
        b.free()
    res.free()


def test_bucket_join_on_string_keys(ctx):
    """JoinIndexRule's join over two indexes on a string column (the reference's E2E join tests join SampleData on c3 = Query,
    T/index/E2EHyperspaceRulesTest.scala): per-bucket merge join on byte order, string and integer columns projected; one
    side with two files per bucket (after an incremental refresh) is re-sorted first."""
    from hyperspace_b200 import _native as N

    rng = np.random.default_rng(21)
    nl, nr, nb = 40_000, 30_000, 8
    vocab = _random_strings(rng, 3000, max_len=10)
    lk = [vocab[i] for i in rng.integers(0, 2500, size=nl)]          # keys 2500..2999 never appear on the left
    rk = [vocab[i] for i in rng.integers(500, 3000, size=nr)]        # keys 0..499 never appear on the right
    lv = rng.integers(0, 10**6, size=nl, dtype=np.int64)
    rs = [b"r-%d" % i for i in range(nr)]
    L = pa.table({"k": pa.array(lk, pa.binary()), "lv": pa.array(lv)})
    R = pa.table({"k": pa.array(rk, pa.binary()), "rs": pa.array(rs, pa.binary())})

    def image(t):
        sink = pa.BufferOutputStream()
        pq.write_table(t, sink, compression="NONE")
        return N.FileImage(data=sink.getvalue().to_pybytes())

    li, _ = ctx.create_index([image(L)], ["k"], ["lv"], nb, output=N.HS_OUT_HOST)
    r1, _ = ctx.create_index([image(R.slice(0, nr // 2))], ["k"], ["rs"], nb, output=N.HS_OUT_HOST)
    r2, _ = ctx.create_index([image(R.slice(nr // 2))], ["k"], ["rs"], nb, output=N.HS_OUT_HOST)
    right = r1.as_sources() + r2.as_sources()
    right_b = [f.bucket for f in r1.files] + [f.bucket for f in r2.files]
    b, st = ctx.bucket_join(li.as_sources(), [f.bucket for f in li.files], right, right_b, nb, "k", "k", ["k", "lv"], ["rs"])
    by_key = {}
    for k, s in zip(rk, rs):
        by_key.setdefault(k, []).append(s)
    want = sorted((k, int(v), s) for k, v in zip(lk, lv) for s in by_key.get(k, ()))
    cols = {n: d for n, d, _ in b.columns}
    got = sorted((bytes(k), int(v), bytes(s)) for k, v, s in zip(cols["k"], cols["lv"], cols["rs"]))
    assert b.num_rows == len(want) and got == want
    b.free()
    for r in (li, r1, r2):
        r.free()

"""Pins the CPU oracle against every golden vector the reference holds for the hot path."""
import numpy as np
import pytest

from oracle import oracle as O


def test_bucket_union_golden_vector():
    # T/index/BucketUnionTest.scala:101-123: int keys {2,3} hash-partitioned into 10 partitions give the
    # per-partition key sums Seq(0, 6, 0, 0, 4, 0, 0, 0, 0, 0)  (each key appears twice: df1 union df2)
    keys = np.array([2, 3, 2, 3], dtype=np.int32)
    b = O.bucket_ids([keys], 10)
    assert b.tolist() == [4, 1, 4, 1]
    sums = [int(keys[b == p].sum()) for p in range(10)]
    assert sums == [0, 6, 0, 0, 4, 0, 0, 0, 0, 0]
    assert O.np_bucket_ids([keys], 10).tolist() == [4, 1, 4, 1]


def test_hash_long_known_answers():
    # Spark's documented hash(1L) and the vectors listed in SURVEY.md section 8c
    expect = {0: -1670924195, 1: -1712319331, 2: -797927272, 3: 519220707, -1: -939490007}
    for k, h in expect.items():
        assert O.lib().hso_hash_long(k, 42) == h
        assert int(O.np_hash_long(np.array([k]))[0]) == h
    ks = np.array(list(expect.keys()), dtype=np.int64)
    assert O.bucket_ids([ks], 200).tolist() == [5, 69, 128, 107, 193]
    assert O.np_bucket_ids([ks], 200).tolist() == [5, 69, 128, 107, 193]


def test_c_and_numpy_agree_on_random_keys():
    rng = np.random.default_rng(7)
    k64 = rng.integers(-2**63, 2**63 - 1, size=100_000, dtype=np.int64)
    k32 = rng.integers(-2**31, 2**31 - 1, size=100_000, dtype=np.int32)
    f64 = rng.standard_normal(100_000)
    f64[:4] = [0.0, -0.0, np.nan, np.inf]
    f32 = f64.astype(np.float32)
    for cols in ([k64], [k32], [k32, k64], [f64], [f32, k64]):
        for n in (1, 7, 200, 1000):
            assert np.array_equal(O.bucket_ids(cols, n), O.np_bucket_ids(cols, n))
    # -0.0 and 0.0 share a bucket; NaN is canonicalised
    assert O.bucket_ids([f64[:2]], 200)[0] == O.bucket_ids([f64[:2]], 200)[1]


def test_null_key_leaves_hash_unchanged():
    k = np.array([5, 5, 9], dtype=np.int64)
    valid = np.array([1, 0, 1], dtype=np.uint8)
    b = O.bucket_ids([k], 200, [valid])
    assert b[1] == 42 % 200  # hash stays at the seed
    assert np.array_equal(b, O.np_bucket_ids([k], 200, [valid]))


def test_hash_bytes_matches_python_restatement():
    for s in (b"", b"a", b"ab", b"abc", b"abcd", b"hello world", bytes(range(250, 256)) + b"xyz"):
        assert O.lib().hso_hash_bytes(s, len(s), 42) == O.py_hash_bytes(s, 42)


def test_sort_perm_orders_by_bucket_then_key():
    rng = np.random.default_rng(11)
    k = rng.integers(-50, 50, size=5000, dtype=np.int64)  # many duplicates
    b = O.bucket_ids([k], 13)
    perm, offs = O.sort_perm([k], 13, b)
    assert sorted(perm.tolist()) == list(range(5000))
    assert offs[0] == 0 and offs[-1] == 5000
    for p in range(13):
        seg = perm[offs[p]:offs[p + 1]]
        assert np.all(b[seg] == p)
        ks = k[seg]
        assert np.all(ks[:-1] <= ks[1:])
        # deterministic tie order: original row index
        same = ks[:-1] == ks[1:]
        assert np.all(seg[:-1][same] < seg[1:][same])
    # numpy lexsort cross-check
    ref = np.lexsort((np.arange(5000), k, b))
    assert np.array_equal(ref, perm)


def test_sort_perm_nulls_first_and_multi_key():
    k1 = np.array([3, 1, 2, 1, 0], dtype=np.int32)
    k2 = np.array([1.5, 2.5, np.nan, -1.0, 0.0])
    v1 = np.array([1, 1, 1, 0, 1], dtype=np.uint8)
    b = np.zeros(5, dtype=np.int32)
    perm, _ = O.sort_perm([k1, k2], 1, b, [v1, None])
    assert perm.tolist() == [3, 4, 1, 2, 0]


def test_range_select_and_merge_join():
    k = np.array([-5, -5, 0, 1, 1, 1, 7, 9], dtype=np.int64)
    assert O.range_select(k, 1, 7) == (3, 7)
    assert O.range_select(k, 2, 6) == (6, 6)
    assert O.range_select(k, -100, 100) == (0, 8)
    l = np.array([1, 1, 2, 5, 9], dtype=np.int64)
    r = np.array([0, 1, 1, 5, 5, 5, 10], dtype=np.int64)
    li, ri = O.merge_join(l, r)
    assert list(zip(li.tolist(), ri.tolist())) == [(0, 1), (0, 2), (1, 1), (1, 2), (3, 3), (3, 4), (3, 5)]


def test_splitmix_c_matches_numpy():
    idx = np.arange(1000, dtype=np.uint64)
    a = O.splitmix64(42, idx)
    for i in (0, 1, 999):
        assert int(a[i]) == O.lib().hso_splitmix64(42, i)


def test_create_index_cpu_path(tmp_path):
    import pyarrow as pa
    import pyarrow.parquet as pq

    cols = O.synthetic_table(0, 10_000, 3)
    src = tmp_path / "src.parquet"
    pq.write_table(pa.table(cols), src)
    files = O.create_index([str(src)], ["k"], ["v1", "v2"], 200, str(tmp_path / "v__=0"), job_uuid="u")
    seen = 0
    for f in files:
        bucket = int(f.rsplit("_", 1)[1].split(".")[0])
        t = pq.ParquetFile(f).read()  # (read_table would infer a hive partition column from "v__=0")
        assert t.column_names == ["k", "v1", "v2"]
        k = np.asarray(t.column("k"))
        assert np.all(O.np_bucket_ids([k], 200) == bucket)
        assert np.all(k[:-1] <= k[1:])
        seen += len(k)
    assert seen == 10_000


def test_division_free_pmod_formula_is_exact():
    """The GPU computes Spark's pmod(hash, n) without a division (hash_partition.cu: make_mod_const / fast_pmod):
    shift the signed hash into unsigned range, Lemire fastmod with M = 2^64 // n + 1, take the shift out again modulo n.
    The restatement below follows that code line by line and must agree with the oracle's pmod for every bucket count."""
    rng = np.random.default_rng(5)
    hashes = np.concatenate([rng.integers(-2**31, 2**31, size=2000, dtype=np.int64),
                             np.array([0, 1, -1, 2**31 - 1, -2**31, 42, -42], dtype=np.int64)])
    mask64 = (1 << 64) - 1
    for n in list(range(1, 300)) + [1000, 1023, 1024, 4095, 4096, 65535, 2**31 - 1]:
        M = (mask64 // n + 1) & mask64          # wraps to 0 for n == 1
        bias = (1 << 31) % n
        for h in hashes.tolist():
            u = (h & 0xffffffff) ^ 0x80000000    # h + 2^31 as an unsigned 32-bit value
            r = (((M * u) & mask64) * n) >> 64   # fastmod: (u mod n) for every 32-bit u
            assert r == u % n
            got = r - bias if r >= bias else r + n - bias
            assert got == h % n                  # Python's % is already the non-negative (pmod) remainder


def test_spark_documented_hash_example():
    """Spark's own documentation of the function the reference buckets with (Murmur3Hash's ExpressionDescription, Spark
    3.1.1 `sql/catalyst/.../expressions/hash.scala`): SELECT hash('Spark', array(123), 2) -> -1321691492.
    It exercises hashUnsafeBytes (5 bytes: one 4-byte block + the per-byte tail), hashInt and the seed fold, seed 42."""
    L = O.lib()
    h = L.hso_hash_bytes(b"Spark", 5, 42)
    h = L.hso_hash_int(123, h)   # an array hashes its elements in order, each seeded with the running hash
    h = L.hso_hash_int(2, h)
    assert h == -1321691492


def test_string_keys_bucket_and_order_like_spark():
    """The oracle's string path, which the GPU string tests are held against: bucket = pmod(hashUnsafeBytes(bytes, 42), n) --
    pinned through Spark's documented hash('Spark', ...) vector above and the pure-Python restatement here -- and order =
    UTF8String.compareTo: unsigned bytes, the shorter value first on a common prefix, nulls first, ties in source order."""
    rng = np.random.default_rng(3)
    vocab = [b"", b"a", b"ab", b"abc", b"b", b"\xc3\xa9", b"\xff", b"\x7f", b"zz", b"facebook", b"donde", b"ibraco", b"miperro"]
    n, nb = 5000, 7
    keys = np.array([vocab[i] for i in rng.integers(0, len(vocab), size=n)], dtype=object)
    valid = rng.random(n) > 0.1
    b = O.bucket_ids([keys], nb, [valid])
    for i in range(0, n, 37):
        h = O.py_hash_bytes(keys[i], 42) if valid[i] else 42  # a null leaves the running hash (the seed) unchanged
        h = h - (1 << 32) if h >= (1 << 31) else h
        assert b[i] == ((h % nb) + nb) % nb, (i, keys[i])
    perm, offs = O.sort_perm([keys], nb, b, [valid])
    assert sorted(perm.tolist()) == list(range(n)) and offs[-1] == n
    for q in range(nb):
        rows = perm[offs[q]:offs[q + 1]]
        assert np.all(b[rows] == q)
        want = sorted(rows.tolist(), key=lambda r: (bool(valid[r]), keys[r] if valid[r] else b"", r))
        assert rows.tolist() == want  # Python's bytes order is the unsigned byte order with shorter-first

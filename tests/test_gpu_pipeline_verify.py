"""GPU tests of the staging / asynchronous createIndex API, of hs_verify_index, and parity at the sizes where the
1 B-row benchmark runs (3 sorted key bytes + tie-run fix-up, many sort tiles per bucket), through the C ABI.

Large cases use the two size-independent instruments together: hs_verify_index over every row (bucket id per row, per-file
sortedness, row multiset via order-independent checksums -- the three properties of
T/index/DataFrameWriterExtensionsTest.scala:93-158) and the CPU oracle for whole buckets (oracle.synthetic_bucket)."""
import os
import subprocess
import sys

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INDEXED, INCLUDED = ["k"], ["v1", "v2", "v3", "v4"]


@pytest.fixture(scope="module")
def ctx():
    from hyperspace_b200 import _native

    c = _native.Context(0)
    yield c
    c.close()


def _files_bytes(res):
    return {f.name: res.host_bytes(i) for i, f in enumerate(res.files)}


def test_staged_async_path_is_byte_identical_to_the_synchronous_call(ctx):
    from hyperspace_b200 import _native as N

    n, nb = 300_000, 50
    hsrc = ctx.synth_table(0, n, 5, n_files=6, row_groups_per_file=2, output=N.HS_OUT_HOST)
    sync, st0 = ctx.create_index(hsrc.as_sources(), INDEXED, INCLUDED, nb, output=N.HS_OUT_HOST, job_uuid="u")
    want = _files_bytes(sync)
    assert st0["ms_d2h"] > 0 and st0["ms_total"] >= st0["ms_d2h"]
    # three calls in flight, as bench.py drives them
    results = []
    nxt, prev = ctx.stage_sources(hsrc.as_sources()), None
    for i in range(3):
        cur, nxt = nxt, (ctx.stage_sources(hsrc.as_sources()) if i < 2 else None)
        assert all(f.on_device for f in cur.as_sources())
        pend = ctx.create_index_async(cur.as_sources(), INDEXED, INCLUDED, nb, output=N.HS_OUT_HOST, job_uuid="u")
        cur.free()
        if prev is not None:
            results.append(prev.wait())
        prev = pend
    results.append(prev.wait())
    for res, st in results:
        assert _files_bytes(res) == want
        assert st["rows_out"] == n and st["gpu_launches"] > 0
        res.free()
    sync.free()
    # a pending build that is dropped without being waited for must not leak or hang
    p = ctx.create_index_async(hsrc.as_sources(), INDEXED, INCLUDED, nb, output=N.HS_OUT_HOST, job_uuid="u")
    p.free()
    hsrc.free()


def test_staged_sources_from_the_file_system_and_files_output(ctx, tmp_path):
    from hyperspace_b200 import _native as N

    n, nb = 40_000, 8
    cols = O.synthetic_table(0, n, 3)
    paths = []
    for i in range(2):
        p = str(tmp_path / f"part-{i:05d}.parquet")
        pq.write_table(pa.table({k: v[i * n // 2:(i + 1) * n // 2] for k, v in cols.items()}), p, compression="NONE")
        paths.append(p)
    staged = ctx.stage_sources([N.FileImage(path=p) for p in paths])
    out_dir = str(tmp_path / "v__=0")
    pend = ctx.create_index_async(staged.as_sources(), ["k"], ["v1", "v2"], nb, out_dir=out_dir, output=N.HS_OUT_FILES, job_uuid="fs")
    assert not os.path.isdir(out_dir) or not os.listdir(out_dir)  # files appear at wait()
    res, st = pend.wait()
    staged.free()
    perm, offs, order = O.index_rows(cols, ["k"], ["v1", "v2"], nb)
    names = sorted(os.listdir(out_dir))
    assert names == sorted(f.name for f in res.files)
    for f in res.files:
        t = pq.read_table(os.path.join(out_dir, f.name))
        lo, hi = int(offs[f.bucket]), int(offs[f.bucket + 1])
        for c in order:
            assert t.column(c).to_numpy().tobytes() == cols[c][perm[lo:hi]].tobytes()
    res.free()


def test_verify_index_accepts_a_good_index_and_pinpoints_a_bad_one(ctx):
    from hyperspace_b200 import _native as N

    n, nb = 500_000, 40
    src = ctx.synth_table(7, n, 5, n_files=4, row_groups_per_file=2, output=N.HS_OUT_DEVICE)
    res, _ = ctx.create_index(src.as_sources(), INDEXED, INCLUDED, nb, output=N.HS_OUT_HOST, job_uuid="v", dictionary=False)
    buckets = [f.bucket for f in res.files]
    good = ctx.verify_index(res.as_sources(), buckets, INDEXED, INCLUDED, nb)
    gen = ctx.synth_checksum(7, n, 5)
    assert good["rows"] == n and good["bucket_mismatches"] == 0 and good["order_violations"] == 0
    assert good["row_checksum"] == gen["row_checksum"] and good["column_checksum"] == gen["column_checksum"]
    # the checksums are those of the oracle's table, too (independent of both the generator kernel and the codecs)
    cols = O.synthetic_table(7, n, 5)
    tab = pa.table(cols)
    buf = pa.BufferOutputStream()
    pq.write_table(tab, buf, compression="NONE", use_dictionary=False)
    raw = ctx.verify_index([N.FileImage(data=buf.getvalue().to_pybytes())], [0], INDEXED, INCLUDED, 1)
    assert raw["row_checksum"] == gen["row_checksum"] and raw["column_checksum"] == gen["column_checksum"]
    assert raw["bucket_mismatches"] == 0 and raw["order_violations"] > 0  # one bucket holds everything; source order is not sorted
    # files attributed to the wrong bucket: every row of the two swapped files is flagged
    swapped = list(buckets)
    swapped[0], swapped[1] = swapped[1], swapped[0]
    bad = ctx.verify_index(res.as_sources(), swapped, INDEXED, INCLUDED, nb)
    assert bad["bucket_mismatches"] == res.files[0].rows + res.files[1].rows
    # one value of one included column changed in place (PLAIN pages: the bytes are in the image): row + that column only
    i = 3
    img = bytearray(res.host_bytes(i))
    t = pq.ParquetFile(pa.BufferReader(bytes(img)))
    off = t.metadata.row_group(0).column(2).data_page_offset  # column v2
    img[off + 64] ^= 0x01
    files = [N.FileImage(data=(bytes(img) if j == i else res.host_bytes(j))) for j in range(len(res.files))]
    bad = ctx.verify_index(files, buckets, INDEXED, INCLUDED, nb)
    assert bad["row_checksum"] != gen["row_checksum"]
    diff = [a != b for a, b in zip(bad["column_checksum"], gen["column_checksum"])]
    assert diff == [False, False, True, False, False]
    assert bad["bucket_mismatches"] == 0 and bad["order_violations"] == 0
    # two adjacent keys swapped inside a file: sortedness breaks, the multiset does not
    img = bytearray(res.host_bytes(i))
    off = t.metadata.row_group(0).column(0).data_page_offset
    body = bytes(img).index(cols_key_bytes(t, 0), off)
    img[body:body + 8], img[body + 8:body + 16] = img[body + 8:body + 16], img[body:body + 8]
    files = [N.FileImage(data=(bytes(img) if j == i else res.host_bytes(j))) for j in range(len(res.files))]
    bad = ctx.verify_index(files, buckets, INDEXED, INCLUDED, nb)
    assert bad["order_violations"] >= 1 and bad["column_checksum"][0] == gen["column_checksum"][0]
    assert bad["row_checksum"] != gen["row_checksum"]  # the two rows exchanged their keys
    res.free()
    src.free()


def cols_key_bytes(parquet_file, row):
    """little-endian bytes of the first two keys of the file (to locate the PLAIN page body)"""
    k = parquet_file.read(columns=["k"]).column("k").to_numpy()
    return k[row:row + 2].tobytes()


@pytest.mark.parametrize("rows,nb,files", [(64_000_000, 200, 64), (64_000_000, 13, 16)])
def test_create_index_parity_in_the_benchmark_regime(ctx, rows, nb, files):
    """64 M rows: thousands of sort tiles per bucket, three LSD passes on the key's top bytes + the tie-run fix-up (with 13
    buckets the runs are as dense as in the 1 B-row / 200-bucket build: ~4.9 M rows per bucket)."""
    from hyperspace_b200 import _native as N

    src = ctx.synth_table(0, rows, 5, n_files=files, row_groups_per_file=2, output=N.HS_OUT_DEVICE)
    res, st = ctx.create_index(src.as_sources(), INDEXED, INCLUDED, nb, output=N.HS_OUT_HOST, job_uuid="big")
    src.free()
    assert st["rows_out"] == rows and len(res.files) == nb
    rep = ctx.verify_index(res.as_sources(), [f.bucket for f in res.files], INDEXED, INCLUDED, nb)
    gen = ctx.synth_checksum(0, rows, 5)
    assert rep["rows"] == rows and rep["bucket_mismatches"] == 0 and rep["order_violations"] == 0
    assert rep["row_checksum"] == gen["row_checksum"] and rep["column_checksum"] == gen["column_checksum"]
    threads = os.cpu_count() or 1
    for i in sorted({0, len(res.files) // 2, len(res.files) - 1}):
        f = res.files[i]
        want = O.synthetic_bucket(0, rows, nb, f.bucket, 5, nthreads=threads)
        got = pq.ParquetFile(pa.py_buffer(res.host_view(i))).read()
        assert got.num_rows == len(want["k"]) == f.rows
        for c in INDEXED + INCLUDED:
            assert np.array_equal(got.column(c).to_numpy().view(np.uint8), want[c].view(np.uint8)), (f.bucket, c)
    res.free()
    ctx.trim()


def test_multi_gpu_parity_under_torchrun():
    """tests/multi_gpu_check.py (every rank's bucket files == the oracle's single-process answer) on all visible GPUs."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "multi-gpu parity ok" in r.stdout


def test_zero_copy_plain_columns_give_byte_identical_files(ctx):
    """PLAIN, null-free, value-aligned columns are read in place by the hash / partition kernels (no decode pass).  Files
    whose row counts are no multiple of the partition tile make tiles straddle two pages (two files)."""
    from hyperspace_b200 import _native as N

    n, nb = 777_777, 64  # 7 files x 111 111 rows: every file boundary falls inside a partition tile
    for dictionary in (True, False):
        src = ctx.synth_table(3, n, 5, n_files=7, row_groups_per_file=1, output=N.HS_OUT_DEVICE, dictionary=dictionary)
        ctx.profile_enable(True)
        zc, _ = ctx.create_index(src.as_sources(), INDEXED, INCLUDED, nb, output=N.HS_OUT_HOST, job_uuid="z")
        kernels = ctx.profile_report()
        ctx.profile_enable(False)
        assert "k_fill_zc_tiles" in kernels, sorted(kernels)
        os.environ["HS_NO_ZEROCOPY"] = "1"
        try:
            ctx.profile_enable(True)
            ref, _ = ctx.create_index(src.as_sources(), INDEXED, INCLUDED, nb, output=N.HS_OUT_HOST, job_uuid="z")
            assert "k_fill_zc_tiles" not in ctx.profile_report()
            ctx.profile_enable(False)
        finally:
            del os.environ["HS_NO_ZEROCOPY"]
        assert _files_bytes(zc) == _files_bytes(ref)
        rep = ctx.verify_index(zc.as_sources(), [f.bucket for f in zc.files], INDEXED, INCLUDED, nb)
        gen = ctx.synth_checksum(3, n, 5)
        assert rep["bucket_mismatches"] == 0 and rep["order_violations"] == 0 and rep["row_checksum"] == gen["row_checksum"]
        zc.free()
        ref.free()
        src.free()
    # a pyarrow file with small pages (fewer rows than a tile) or nulls is not eligible and must still come out right
    cols = O.synthetic_table(0, 50_000, 3)
    buf = pa.BufferOutputStream()
    pq.write_table(pa.table(cols), buf, compression="NONE", use_dictionary=False, data_page_size=8192)
    res, _ = ctx.create_index([N.FileImage(data=buf.getvalue().to_pybytes())], ["k"], ["v1", "v2"], 5, output=N.HS_OUT_HOST)
    perm, offs, order = O.index_rows(cols, ["k"], ["v1", "v2"], 5)
    for i, f in enumerate(res.files):
        t = pq.ParquetFile(pa.BufferReader(res.host_bytes(i))).read()
        for c in order:
            assert t.column(c).to_numpy().tobytes() == cols[c][perm[int(offs[f.bucket]):int(offs[f.bucket + 1])]].tobytes()
    res.free()

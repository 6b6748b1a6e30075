"""CPU-only test of the native format layer (hyperspace_b200/csrc/thrift_compact.h + parquet_meta.h).

tests/native/format_roundtrip.cu builds an index-shaped Parquet file on the host with the engine's own writers (page
headers, definition-level splits, dictionary pages, footer with key statistics), parses it back with the engine's footer
reader, and pyarrow -- an independent implementation -- must read the same rows.  nvcc compiles host code here; the
program makes no CUDA call, so it runs without a GPU."""
import os
import shutil
import subprocess

import numpy as np
import pyarrow.parquet as pq
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def roundtrip(tmp_path_factory):
    if shutil.which("nvcc") is None:
        pytest.skip("nvcc not on PATH")
    d = tmp_path_factory.mktemp("native")
    exe, out = str(d / "format_roundtrip"), str(d / "host_written.parquet")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "native", "format_roundtrip.cu")])
    report = subprocess.check_output([exe, out], text=True)
    return out, report


def test_footer_reader_rejects_mutated_footers_without_crashing(roundtrip, tmp_path):
    """20 000 deterministic mutations of a valid footer (byte flips, false lengths, truncations, 0xff / 0x00 runs): the
    engine's reader must return or raise its own error every time -- no crash, no hang, no allocation sized by an untrusted
    count (a list count larger than the bytes that are left is corrupt by definition)."""
    out, _ = roundtrip
    exe = str(tmp_path / "footer_fuzz")
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "native", "footer_fuzz.cu")])
    report = subprocess.run([exe, out, "20000"], text=True, capture_output=True, timeout=120)
    assert report.returncode == 0, report.stdout + report.stderr
    fields = dict(kv.split("=") for kv in report.stdout.split())
    assert int(fields["rounds"]) == 20000 and int(fields["rejected"]) + int(fields["accepted"]) == 20000
    assert int(fields["rejected"]) > 5000  # most mutations are detected as corrupt


def test_engine_footer_reader_parses_what_the_engine_writer_wrote(roundtrip):
    _, report = roundtrip
    lines = report.splitlines()
    assert "aligned_prefixes=ok" in lines  # page bodies of >= 4096-row pages start 8-byte aligned at every file offset
    assert "rows=1000 row_groups=3 columns=3 stat_slots=3" in lines
    assert [l for l in lines if l.startswith("column ")] == ["column k type=2 optional=1", "column d type=1 optional=1",
                                                             "column x type=5 optional=1"]
    chunks = [l.split() for l in lines if l.startswith("rg ")]
    assert [(c[1], c[3], c[4]) for c in chunks] == [(str(g), str(c), f"values={n}") for g, n in ((0, 400), (1, 400), (2, 200))
                                                    for c in range(3)]
    # chunks are laid out back to back: start + bytes of one == start of the next
    spans = [(int(c[5].split("=")[1]), int(c[6].split("=")[1])) for c in chunks]
    assert spans[0][0] == 4 and all(a[0] + a[1] == b[0] for a, b in zip(spans, spans[1:]))
    # the recorded statistics slots point at the min / max bytes inside the serialised footer
    assert [l for l in lines if l.startswith("stat ")] == ["stat rg 0 col 0 min=-300 max=2493", "stat rg 1 col 0 min=2500 max=5293",
                                                           "stat rg 2 col 0 min=5300 max=6693"]


def test_pyarrow_reads_the_host_written_file(roundtrip):
    out, _ = roundtrip
    t = pq.read_table(out)
    i = np.arange(1000)
    assert t.column_names == ["k", "d", "x"] and t.num_rows == 1000
    assert np.array_equal(t.column("k").to_numpy(), i * 7 - 300)
    assert np.array_equal(t.column("d").to_numpy(), np.array([10, 20, 30, 40, 50], dtype=np.int32)[(i * 3) % 5])
    x = t.column("x").to_pylist()
    assert all((v is None) == (j % 7 == 3) for j, v in enumerate(x))
    assert all(v == j * 0.25 for j, v in enumerate(x) if v is not None)
    md = pq.ParquetFile(out).metadata
    assert md.num_row_groups == 3 and [md.row_group(g).num_rows for g in range(3)] == [400, 400, 200]
    for g, (lo, hi) in enumerate(((0, 400), (400, 800), (800, 1000))):
        st = md.row_group(g).column(0).statistics
        assert st.has_min_max and st.min == lo * 7 - 300 and st.max == (hi - 1) * 7 - 300 and st.null_count == 0
        assert md.row_group(g).column(1).has_dictionary_page and not md.row_group(g).column(0).has_dictionary_page
        assert md.row_group(g).column(2).statistics.null_count == sum(1 for j in range(lo, hi) if j % 7 == 3)
    # Spark reads its schema from the footer's key-value metadata
    meta = md.metadata[b"org.apache.spark.sql.parquet.row.metadata"].decode()
    assert '"name":"k"' in meta.replace(" ", "") and '"type":"long"' in meta.replace(" ", "")
    # row-group pruning on the key works through the statistics
    hit = pq.read_table(out, filters=[("k", "==", 2500)])
    assert hit.num_rows == 1 and hit.column("d")[0].as_py() == [10, 20, 30, 40, 50][(400 * 3) % 5]

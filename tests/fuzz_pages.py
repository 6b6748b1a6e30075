"""Device-side fuzzing of the Parquet page readers (run as a script on a GPU box; tests/test_gpu_fuzz.py spawns it).

    python tests/fuzz_pages.py [--iterations N] [--seed S]

pyarrow writes small files in every shape the decoder handles (PLAIN / dictionary, v1 / v2 pages, optional / required,
snappy / uncompressed, booleans); each iteration corrupts a few bytes inside page HEADERS, level blocks, run headers or
index streams and hands the image to hs_create_index.  Whatever the mutation, the call must either succeed (the mutation hit
something harmless) or fail with HS_EFORMAT / HS_EUNSUPPORTED / HS_EINVAL -- never HS_ECUDA (an out-of-bounds access), and
the context must stay usable.  Run it under `compute-sanitizer --tool memcheck` to also catch out-of-bounds accesses that
happen not to fault (profiles/r02_sanitizer_*.log).  Runs in its own process because a CUDA fault is sticky.
"""
import argparse
import io
import json
import os
import sys

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_images(rng):
    n = 6000
    k = rng.integers(-2**62, 2**62, size=n, dtype=np.int64)
    v1 = rng.integers(0, 37, size=n, dtype=np.int64)          # dictionary-friendly
    v2 = rng.standard_normal(n)
    v3 = rng.integers(0, 5, size=n).astype(np.int32)
    vb = rng.integers(0, 2, size=n).astype(bool)
    mask = rng.random(n) < 0.1
    images = []
    for opts in (dict(use_dictionary=False, data_page_version="1.0", compression="NONE"),
                 dict(use_dictionary=True, data_page_version="1.0", compression="NONE"),
                 dict(use_dictionary=True, data_page_version="2.0", compression="NONE"),
                 dict(use_dictionary=True, data_page_version="1.0", compression="SNAPPY"),
                 dict(use_dictionary=["v1", "v3"], data_page_version="2.0", compression="SNAPPY")):
        for nullable in (False, True):
            cols = {"k": pa.array(k), "v1": pa.array(v1, mask=mask if nullable else None),
                    "v2": pa.array(v2, mask=mask if nullable else None), "v3": pa.array(v3), "vb": pa.array(vb)}
            t = pa.table(cols)
            if not nullable:
                t = t.cast(pa.schema([pa.field(f.name, f.type, nullable=False) for f in t.schema]))
            buf = io.BytesIO()
            pq.write_table(t, buf, data_page_size=4096, row_group_size=4000, **opts)
            images.append(buf.getvalue())
    return images


def page_regions(image):
    """(offset, length) of every column chunk: page headers, level blocks and bodies all live there."""
    md = pq.ParquetFile(pa.BufferReader(image)).metadata
    out = []
    for rg in range(md.num_row_groups):
        for c in range(md.num_columns):
            col = md.row_group(rg).column(c)
            start = col.dictionary_page_offset if col.has_dictionary_page and col.dictionary_page_offset else col.data_page_offset
            out.append((start, col.total_compressed_size))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    from hyperspace_b200 import _native as N

    rng = np.random.default_rng(args.seed)
    images = make_images(rng)
    regions = [page_regions(im) for im in images]
    ctx = N.Context(0)
    counts = {"ok": 0, "rejected": 0, "cuda_error": 0, "other": 0}
    codes = {}
    for it in range(args.iterations):
        i = int(rng.integers(0, len(images)))
        img = bytearray(images[i])
        for _ in range(int(rng.integers(1, 4))):
            start, length = regions[i][int(rng.integers(0, len(regions[i])))]
            # two thirds of the mutations land in the first 40 bytes of a chunk (page header + level prefix + run headers)
            span = min(length, 40) if rng.random() < 0.66 else length
            pos = start + int(rng.integers(0, max(1, span)))
            if rng.random() < 0.5:
                img[pos] = int(rng.integers(0, 256))
            else:
                img[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            if it % 3 == 2:  # the read side decodes through the same kernels and also takes the boolean column
                b, _ = ctx.filter_scan([N.FileImage(data=bytes(img))], "k", ["v1", "vb"], lo=-2**61, hi=2**61, sorted_on_key=False)
                b.free()
            else:
                res, _ = ctx.create_index([N.FileImage(data=bytes(img))], ["k"], ["v1", "v2", "v3"], 8, output=N.HS_OUT_DEVICE)
                res.free()
            counts["ok"] += 1
        except N.HyperspaceGpuError as e:
            codes[e.code] = codes.get(e.code, 0) + 1
            if e.code == N.HS_ECUDA:
                counts["cuda_error"] += 1
                print(f"iteration {it}: CUDA error: {e.message}", file=sys.stderr)
                break
            elif e.code in (N.HS_EFORMAT, N.HS_EUNSUPPORTED, N.HS_EINVAL):
                counts["rejected"] += 1
            else:
                counts["other"] += 1
    # the context is still healthy: an intact file builds and verifies
    healthy = False
    if counts["cuda_error"] == 0:
        res, st = ctx.create_index([N.FileImage(data=images[1])], ["k"], ["v1", "v2", "v3"], 8, output=N.HS_OUT_DEVICE)
        healthy = st["rows_out"] == 6000
        res.free()
    ctx.close()
    print(json.dumps({"iterations": args.iterations, "counts": counts, "codes": {str(k): v for k, v in codes.items()}, "healthy": healthy}))
    sys.exit(0 if counts["cuda_error"] == 0 and counts["other"] == 0 and healthy else 1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench_queries.py -- read-side measurements (BASELINE.json configs[2] and [3]) on ONE GPU; not the driver's contract
(that is bench.py).  Index file images stay resident in HBM; results are copied back to the host inside the timed region.

  C3  range filter `k BETWEEN lo AND hi` covering 1 % of the int64 key space over the 1 B-row / 200-bucket index,
      projecting k, v1, v2 (~10 M rows out); 20 distinct ranges; reports queries/s and rows/s.
  C4' bucket-aligned merge join of two 500 M-row indexes (rows [0, 500 M) and [250 M, 750 M) of T: 250 M matching keys),
      `SELECT L.v1, R.v2`; reports joins/s and output rows/s.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--join-rows", type=int, default=500_000_000)
    ap.add_argument("--queries", type=int, default=20)
    args = ap.parse_args()
    import torch

    from hyperspace_b200 import _native as N

    stream = torch.cuda.current_stream()
    ctx = N.Context(0, stream.cuda_stream)
    nb, files = 200, 256

    def build(first, rows, included):
        src = ctx.synth_table(first, rows, 5, n_files=files, row_groups_per_file=4, output=N.HS_OUT_DEVICE)
        idx, st = ctx.create_index(src.as_sources(), ["k"], included, nb, output=N.HS_OUT_DEVICE, job_uuid="q")
        src.free()
        ctx.trim()
        return idx

    # ---- C3 ---------------------------------------------------------------------------------------------
    idx = build(0, args.rows, ["v1", "v2"])
    width = int(0.01 * 2**64)
    ranges = [(-(width // 2) + i * (width // 40), (width // 2) + i * (width // 40)) for i in range(args.queries)]
    ctx.filter_scan(idx.as_sources(), "k", ["k", "v1", "v2"], lo=ranges[0][0], hi=ranges[0][1])[0].free()  # warm
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    out_rows = 0
    t0 = time.perf_counter()
    for lo, hi in ranges:
        b, st = ctx.filter_scan(idx.as_sources(), "k", ["k", "v1", "v2"], lo=lo, hi=hi)
        out_rows += b.num_rows
        b.free()
    e1.record(stream)
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3
    print("last filter stats", {k: round(v, 2) for k, v in st.items() if v}, file=sys.stderr)
    print(json.dumps({"metric": "filter queries/sec", "config": {"workload": "C3: k BETWEEN lo AND hi (1% of key space) over the "
                      f"{args.rows}-row 200-bucket index, project k,v1,v2", "index": "resident in HBM", "n_gpus": 1},
                      "value": args.queries / sec, "unit": "queries/s", "ms_per_query": sec * 1e3 / args.queries,
                      "rows_out_per_query": out_rows / args.queries, "rows_out_per_s": out_rows / sec,
                      "wall_s": time.perf_counter() - t0}))
    idx.free()
    ctx.trim()
    # ---- C4' ---------------------------------------------------------------------------------------------
    L = build(0, args.join_rows, ["v1"])
    R = build(args.join_rows // 2, args.join_rows, ["v2"])
    lb, rb = [f.bucket for f in L.files], [f.bucket for f in R.files]
    ctx.bucket_join(L.as_sources(), lb, R.as_sources(), rb, nb, "k", "k", ["v1"], ["v2"])[0].free()  # warm
    torch.cuda.synchronize()
    e0.record(stream)
    reps = 3
    for _ in range(reps):
        b, st = ctx.bucket_join(L.as_sources(), lb, R.as_sources(), rb, nb, "k", "k", ["v1"], ["v2"])
        nout = b.num_rows
        b.free()
    e1.record(stream)
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / reps
    print(json.dumps({"metric": "join queries/sec", "config": {"workload": f"C4': {args.join_rows} x {args.join_rows} rows, 200 buckets, "
                      "bucket-aligned merge join, SELECT L.v1, R.v2", "index": "resident in HBM", "n_gpus": 1},
                      "value": 1 / sec, "unit": "joins/s", "ms_per_join": sec * 1e3, "rows_out": nout,
                      "rows_out_per_s": nout / sec, "algorithmic_GB": (16 * 2 * args.join_rows + 16 * nout) / 1e9}))
    print("last join stats", {k: round(v, 2) for k, v in st.items() if v}, file=sys.stderr)
    ctx.close()


if __name__ == "__main__":
    main()

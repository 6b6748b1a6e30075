#!/usr/bin/env python
"""bench.py -- createIndex rows/s on the synthetic table T of SURVEY.md section 8d (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W [--rows R] [--impl reference]

A "step" is one createIndex over the whole table: scan (Parquet decode) -> project -> hash-repartition into 200 buckets
-> sort within bucket -> Parquet encode, through the C ABI (hs_create_index).

* ``value``  rows/s with the source Parquet file images already resident in HBM and the index file images left in HBM.
* ``e2e``    the same call with HOST file images in and HOST file images out (pinned memory); H2D and D2H inside the
             timed region.
* ``roofline``  achieved HBM GB/s of the dominant kernel (k_sort_scatter: 24 algorithmic bytes per row per launch, 20 for a step's first launch),
             from CUDA events recorded by the library around every launch on its stream.
* ``cpu_baseline``  the CPU oracle port (pyarrow decode/encode + pthreads C bucket/sort) timed on this host's cores on a
             bounded sample of the same table (rank 0, N=1 only).
* ``--impl reference``  the reference arm.  The reference itself (Scala on Spark) cannot run here (no JVM in this
             image), so this arm times the oracle port with all host threads, as the task statement prescribes.

Multi-GPU (torchrun, one rank per GPU): the 256 source files are split across ranks, rows move to the owner of their
bucket with one NCCL all-to-all, the table size is fixed (strong scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

INDEXED = ["k"]
INCLUDED = ["v1", "v2", "v3", "v4"]
NUM_BUCKETS = 200
ROW_BYTES = 32  # decoded bytes per row of T
ALGO_BYTES_PER_ROW = 64  # 32 read + 32 written
SORT_SCATTER_BYTES_PER_ROW = 24  # k_sort_scatter: (8 B key + 4 B row index) read + written once per launch ...
SORT_SCATTER_FIRST_PASS_BYTES_PER_ROW = 20  # ... except a step's first launch: reads the raw 8 B key column only


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--files", type=int, default=256)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-rows", type=int, default=16_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--plain", action="store_true", help="PLAIN-only source and index files (no dictionary encoding)")
    return ap.parse_args()


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(smax), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------------------
# CPU oracle arm
# ---------------------------------------------------------------------------------------------------------------------

def cpu_create_index(sample_rows: int, nthreads: int, workdir: str):
    """Times the oracle port (CPU restatement of the reference path) on `sample_rows` rows of T.  Returns seconds."""
    import numpy as np
    import pyarrow as pa
    import pyarrow.parquet as pq
    from concurrent.futures import ThreadPoolExecutor

    from oracle import oracle as O

    pa.set_cpu_count(nthreads)
    src_dir = os.path.join(workdir, "src")
    os.makedirs(src_dir, exist_ok=True)
    n_files = 8
    per = sample_rows // n_files
    paths = []
    for f in range(n_files):
        p = os.path.join(src_dir, f"part-{f:05d}.parquet")
        if not os.path.exists(p):
            pq.write_table(pa.table(O.synthetic_table(f * per, per, 5)), p, compression="NONE", use_dictionary=["v1", "v3", "v4"])
        paths.append(p)
    out_dir = os.path.join(workdir, "idx")
    t0 = time.perf_counter()
    order = INDEXED + INCLUDED
    with ThreadPoolExecutor(max_workers=min(nthreads, n_files)) as ex:
        tables = list(ex.map(lambda p: pq.read_table(p, columns=order, use_threads=True), paths))
    t = pa.concat_tables(tables).combine_chunks()
    cols = {name: t.column(name).chunk(0).to_numpy() for name in order}
    perm, offs, _ = O.index_rows(cols, INDEXED, INCLUDED, NUM_BUCKETS, nthreads=nthreads)
    os.makedirs(out_dir, exist_ok=True)

    def write_bucket(b):
        lo, hi = int(offs[b]), int(offs[b + 1])
        if hi == lo:
            return
        idx = perm[lo:hi]
        part = pa.table({name: cols[name][idx] for name in order})
        pq.write_table(part, os.path.join(out_dir, O.bucket_file_name(b, "cpu")), compression="NONE",
                       use_dictionary=["v1", "v3", "v4"])

    with ThreadPoolExecutor(max_workers=nthreads) as ex:
        list(ex.map(write_bucket, range(NUM_BUCKETS)))
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    O.build()
    cores = os.cpu_count() or 1
    rows = args.cpu_sample_rows
    with tempfile.TemporaryDirectory() as wd:
        for _ in range(max(1, min(args.warmup, 1))):
            cpu_create_index(rows, cores, wd)
        times = [cpu_create_index(rows, cores, wd) for _ in range(args.steps)]
    sec = sum(times) / len(times)
    value = rows / sec
    line = {
        "impl": "reference", "metric": "createIndex rows/sec", "value": value, "unit": "rows/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args, sample_rows=rows),
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": cores, "kind": "port",
                         "sample": f"{rows} rows of T per step (pyarrow decode/encode + pthreads C bucket/sort; the reference "
                                   "is Scala on Spark and cannot run without a JVM)"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, sample_rows=None):
    cfg = {"workload": "createIndex: 1B rows x (k:int64 indexed; v1:int64, v2:float64, v3:int32, v4:float32 included), "
                       "200 buckets, 256 source Parquet files" if args.rows == 1_000_000_000 else
                       f"createIndex: {args.rows} rows x 5 columns of T, 200 buckets, {args.files} source Parquet files",
           "rows": args.rows, "source_files": args.files, "num_buckets": NUM_BUCKETS, "source_encoding": "PLAIN, UNCOMPRESSED" if args.plain else
           "PLAIN_DICTIONARY (v1, v3, v4) + PLAIN (k, v2), UNCOMPRESSED -- what parquet-mr / pyarrow write by default",
           "index_encoding": "PLAIN, UNCOMPRESSED" if args.plain else "PLAIN_DICTIONARY (v1, v3, v4) + PLAIN (k, v2), UNCOMPRESSED", "l2": "inputs (>= 32 B/row x rows) far exceed the 126 MB L2; no flush needed"}
    if sample_rows:
        cfg["cpu_sample_rows"] = sample_rows
    return cfg


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------

def run_ours(args):
    import numpy as np
    import torch

    from hyperspace_b200 import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    stream = torch.cuda.current_stream()
    ctx = N.Context(local_rank, stream.cuda_stream)
    if world > 1:
        ids = [N.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(rank, world, ids[0])

    # this rank's share of the table: files [f0, f1) of args.files, rows split evenly over files
    n_files = args.files
    rows_per_file = args.rows // n_files
    f0, f1 = rank * n_files // world, (rank + 1) * n_files // world
    my_files = f1 - f0
    my_rows = my_files * rows_per_file
    total_rows = rows_per_file * n_files
    first_row = f0 * rows_per_file

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if not dist:
            return ms
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs resident in HBM -----------------------------------------------------------------------------------
    src = ctx.synth_table(first_row, my_rows, 5, n_files=my_files, row_groups_per_file=4, output=N.HS_OUT_DEVICE,
                          dictionary=not args.plain)
    sources = src.as_sources()
    src_bytes = sum(f.size for f in src.files)

    def step_device():
        res, st = ctx.create_index(sources, INDEXED, INCLUDED, NUM_BUCKETS, output=N.HS_OUT_DEVICE, job_uuid="bench",
                                   dictionary=not args.plain)
        res.free()
        return st

    for _ in range(args.warmup):
        step_device()
    barrier()
    ctx.profile_enable(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    launches = 0
    stage_ms = {}
    for _ in range(args.steps):
        st = step_device()
        launches += int(st["gpu_launches"])
        for k, v in st.items():
            if k.startswith("ms_"):
                stage_ms[k] = stage_ms.get(k, 0.0) + v / args.steps
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    kernels = ctx.profile_report()
    ctx.profile_enable(False)
    ms_dev = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    value = total_rows / (ms_dev / 1e3)

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------------
    top_name, top = max(kernels.items(), key=lambda kv: kv[1]["ms"]) if kernels else (None, None)
    roofline = None
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback"
    if os.path.exists(peaks_path):
        try:
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    if top_name:
        ss = kernels.get("k_sort_scatter", top)
        avg_ms = ss["ms"] / max(1, ss["launches"])
        rows_after_exchange = total_rows / world  # rows each rank sorts (uniform hash)
        per_step = max(1.0, ss["launches"] / args.steps)  # radix passes per step; the first one reads no row indices
        bytes_per_row = (SORT_SCATTER_FIRST_PASS_BYTES_PER_ROW + SORT_SCATTER_BYTES_PER_ROW * (per_step - 1)) / per_step
        achieved = bytes_per_row * rows_after_exchange / (avg_ms / 1e3) / 1e9
        # DRAM traffic of the kernel from the committed ncu capture (bytes per row are size-independent for this kernel:
        # every pair is read once and written once), scaled to the rows of one launch here
        traffic, traffic_note = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))["k_sort_scatter"]
            per_row = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) / tr["rows_per_launch"]
            traffic = per_row * rows_after_exchange
            traffic_note = (f"dram__bytes_read+write.sum = {per_row:.2f} B/row measured by ncu --set full at "
                            f"{tr['rows_per_launch']} rows/launch (profiles/r01_ncu_traffic.json), scaled to this launch size")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": "k_sort_scatter", "achieved": achieved, "peak": peak, "peak_source": peak_src,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                    "algorithmic_bytes_per_launch": bytes_per_row * rows_after_exchange,
                    "algorithmic_bytes_note": f"{per_step:.0f} launches per step: the first moves 20 B/row (raw key in, "
                                              "encoded key + row index out), the others 24 B/row",
                    "avg_launch_ms": avg_ms, "launches_timed": ss["launches"],
                    "whole_path": {"achieved": ALGO_BYTES_PER_ROW * value / world / 1e9, "unit": "GB/s",
                                   "frac": ALGO_BYTES_PER_ROW * value / world / 1e9 / peak,
                                   "note": "64 algorithmic B/row x rows/s per GPU vs HBM peak (SURVEY.md 8d yardstick)"},
                    "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])}}

    # ---- e2e: host images in, host images out -----------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        # "the caller's Parquet files in host memory": the same synthetic table, generated again straight into pinned
        # host memory (outside the timed region)
        src.free()
        ctx.trim()
        hsrc = ctx.synth_table(first_row, my_rows, 5, n_files=my_files, row_groups_per_file=4, output=N.HS_OUT_HOST,
                               dictionary=not args.plain)
        host_in = hsrc.as_sources()

        def step_host():
            res, st = ctx.create_index(host_in, INDEXED, INCLUDED, NUM_BUCKETS, output=N.HS_OUT_HOST, job_uuid="bench",
                                       dictionary=not args.plain)
            out_bytes = sum(f.size for f in res.files)
            # read the result on the host: first and last byte of every file image (the images are complete Parquet files)
            chk = 0
            for i in range(len(res.files)):
                v = res.host_view(i)
                chk += int(v[0]) + int(v[-1])
            res.free()
            return st, out_bytes, chk

        for _ in range(args.warmup):
            step_host()
        barrier()
        e0.record(stream)
        out_bytes = 0
        for _ in range(args.steps):
            st, out_bytes, _ = step_host()
        e1.record(stream)
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        e2e = {"value": total_rows / (ms_e2e / 1e3), "unit": "rows/s", "h2d_bytes_per_step": int(src_bytes),
               "d2h_bytes_per_step": int(out_bytes), "ms_per_step": ms_e2e,
               "note": "hs_create_index with pinned HOST Parquet images in and HOST index images out; per-rank bytes"}
        hsrc.free()

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle as O

            O.build()
            cores = os.cpu_count() or 1
            with tempfile.TemporaryDirectory() as wd:
                cpu_create_index(args.cpu_sample_rows, cores, wd)  # warm (also writes the sample source files)
                sec = cpu_create_index(args.cpu_sample_rows, cores, wd)
            cpu = {"value": args.cpu_sample_rows / sec, "unit": "rows/s", "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_sample_rows} rows of T, one createIndex ({sec:.2f} s): pyarrow decode/encode + "
                             "pthreads C Murmur3 bucket + per-bucket radix sort (oracle port, not Spark)"}
        except Exception as ex:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": "rows/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {
            "metric": "createIndex rows/sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "config": workload_config(args),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "stage_ms_per_step": stage_ms,
        }
        print(json.dumps(line))
    ctx.close()
    if dist:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/bin/bash
# Reproduces the rows of profiles/r02_nvlink_runs.md that the code still has switches for (2 x B200, run under
# `gpurun --gpus 2 -- bash tools/n2_experiments.sh`): the fused partition + exchange kernel with plain / bulk stores, with
# all destinations forced local (the kernel without the link), with the small tile shape, and with fewer buckets (longer
# runs per (tile, bucket)).  Prints one line per variant: G rows/s, ms/step, k_partition_rows ms, stage times.
# (The persistent-CTA and 16 K-row-tile variants of that table were removed from the code after they lost.)
run() {
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus 2 --steps 6 --warmup 3 --no-extra --no-e2e > gpurun_out/r02_n2_$name.json 2> gpurun_out/r02_n2_$name.err
  python -c "
import json
d = json.loads(open('gpurun_out/r02_n2_$name.json').read().strip().splitlines()[-1])
ks = d['roofline']['kernel_ms_per_step']
print('$name', round(d['value'] / 1e9, 2), 'G rows/s', round(d['ms_per_step'], 2), 'ms; verified', (d['verified'] or {}).get('ok'),
      '; k_partition_rows', round(ks.get('k_partition_rows', 0), 2), 'ms;', {k: round(v, 2) for k, v in d['stage_ms_per_step'].items() if v})
" || tail -5 gpurun_out/r02_n2_$name.err
}
run default HS_NOOP=1
run plain_stores HS_PART_BULK=0
run local_plain HS_DEBUG_LOCAL_PEERS=1 HS_PART_BULK=0     # wrong results by design: isolates the link
run local_bulk HS_DEBUG_LOCAL_PEERS=1
run small_tiles HS_PEER_TILE=small
run small_tiles_plain HS_PEER_TILE=small HS_PART_BULK=0
run buckets50 HS_BENCH_BUCKETS=50
run buckets50_local HS_BENCH_BUCKETS=50 HS_DEBUG_LOCAL_PEERS=1
run buckets16 HS_BENCH_BUCKETS=16
run buckets16_local HS_BENCH_BUCKETS=16 HS_DEBUG_LOCAL_PEERS=1

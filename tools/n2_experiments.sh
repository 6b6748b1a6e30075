run() { name=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra --no-e2e $EXTRA > gpurun_out/r02_n2_$name.json 2> gpurun_out/r02_n2_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_n2_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value']/1e9,2), round(d['ms_per_step'],2), (d['verified'] or {}).get('ok'), 'part', round(d['roofline']['kernel_ms_per_step']['k_partition_rows'],2), 'exch', round(d['stage_ms_per_step']['ms_exchange'],2))"; }
EXTRA=--no-verify run localpeers HS_DEBUG_LOCAL_PEERS=1
EXTRA=--no-verify run localpeers_plain HS_DEBUG_LOCAL_PEERS=1 HS_PART_BULK=0
EXTRA= run small_plain HS_PEER_TILE=small HS_PART_BULK=0
EXTRA= run small_bulk HS_PEER_TILE=small HS_PART_BULK=1

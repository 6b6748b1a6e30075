run() { name=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 6 --warmup 3 --no-extra --no-e2e $EXTRA > gpurun_out/r02_n2_$name.json 2> gpurun_out/r02_n2_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/r02_n2_$name.json').read().strip().splitlines()[-1]); ks=d['roofline']['kernel_ms_per_step']; print('$name', round(d['value']/1e9,2), round(d['ms_per_step'],2), (d['verified'] or {}).get('ok'), 'part', round(ks.get('k_partition_rows',0),2), 'kernels', round(sum(ks.values()),2), {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v})" || tail -5 gpurun_out/r02_n2_$name.err; }
python -m pytest tests/test_gpu_pipeline_verify.py -m gpu -q -k multi_gpu 2>&1 | tail -3
EXTRA= run stream16k HS_PART_STREAM=1
EXTRA= run nostream16k HS_PART_STREAM=0
EXTRA= run nostream16k_plain HS_PART_STREAM=0 HS_PART_BULK=0

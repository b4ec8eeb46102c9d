#!/bin/bash
# EIGHT GPUs: the driver's scaling command at N=8 (device-timed rate, e2e with host buffers, train_e2e)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02_c3_n8.json 2> gpurun_out/bench_r02_c3_n8.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_c3_n8.json').read().strip().splitlines()[-1]); print('N=8', round(d['value']/1e6,1), 'M', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']/1e6,1), 'upd-keys', round(d['e2e']['update_keys']['value']/1e6,1), 'train', d.get('train_e2e',{}).get('value'), d.get('gather'))"
tail -4 gpurun_out/bench_r02_c3_n8.err | cut -c1-300
echo done

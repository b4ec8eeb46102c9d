#!/bin/bash
# TWO GPUs: lock-step replicas (parameters broadcast from rank 0), gloo-free; bench at N=2
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multirank.py -m gpu -q --tb=short -s > gpurun_out/r2_pytest_2gpu.log 2>&1
grep "MULTIRANK\|passed\|failed\|AssertionError: (" gpurun_out/r2_pytest_2gpu.log | cut -c1-250 | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_c3_n2.json 2> gpurun_out/bench_r02_c3_n2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_c3_n2.json').read().strip().splitlines()[-1]); print('N=2', round(d['value']/1e6,1), 'M', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']/1e6,1), 'upd-keys', round(d['e2e']['update_keys']['value']/1e6,1), 'train', round(d['train_e2e']['value']/1e6,2) if d.get('train_e2e') and 'value' in d['train_e2e'] else d.get('train_e2e'))"
tail -3 gpurun_out/bench_r02_c3_n2.err | cut -c1-300
echo done

#!/bin/bash
# cart-pole with carried sin/cos now that the episode-start preparation is off the critical path (vC1) vs main; hidden-128 tensor-core collector with
# the two-warp preparation; bench lines of c3 / c2 / c4 / hidden 128
mkdir -p gpurun_out; rm -f gpurun_out/r2_ab27.log
for rep in 1 2 3; do
for c in HEAD vC1; do
  lib=libscg_b200.so; [ $c = vC1 ] && lib=libscg_b200_vC1.so
  ( echo "== $c" && SCG_LIB_NAME=$lib SCG_MLP_ENGINE=ffma2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=25)
q.run('cartpole_stab', 4096, 200, 64, 'f64', 32, iters=9)
" 2>&1 | tail -3 ) >> gpurun_out/r2_ab27.log
done
done
awk '/^==/{c=$2} /N=/{print c, $1, $7, $(NF-5), $(NF-4)}' gpurun_out/r2_ab27.log | sort -s -k1,3
timeout 400 python bench.py > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; tail -c 200 gpurun_out/bench_r02_c3.json
timeout 300 python bench.py --config c2 --no-train > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err
timeout 300 python bench.py --config c4 --no-train --steps 10 > gpurun_out/bench_r02_c4.json 2> gpurun_out/bench_r02_c4.err
timeout 300 python bench.py --hidden 128 --no-train --no-cpu-baseline --steps 10 > gpurun_out/bench_r02_c3_h128.json 2> gpurun_out/bench_r02_c3_h128.err
for f in c3 c2 c4 c3_h128; do python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,1), 'M', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']/1e6,1))"; done
timeout 600 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_api.py -m gpu -q --tb=short > gpurun_out/r2_pytest_some.log 2>&1
tail -4 gpurun_out/r2_pytest_some.log
echo done

#!/bin/bash
# decoupled critic (CA schedule): whole GPU suite on the new library, A/B against the previous build on the same box, per-warp phase timing, bench line
mkdir -p gpurun_out; rm -f gpurun_out/r2_ab33.log gpurun_out/r2_prof_ticks33.log
SCG_LIB_NAME=libscg_b200_vca.so timeout 200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_pytest_gpu_ca.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_ca.log | cut -c1-300
ab() {
for c in libscg_b200.so libscg_b200_vca.so; do
  ( echo "== $c" && SCG_LIB_NAME=$c SCG_MLP_ENGINE=ffma2 timeout 100 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=15)
q.run('quadrotor_3D_track', 8192, 200, 64, 'f64', 32, iters=5)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=15)
" 2>&1 | tail -3 ) >> gpurun_out/r2_ab33.log
done
awk '/^==/{c=$2} /N=/{print c, $1, $7, $(NF-5), $(NF-4)}' gpurun_out/r2_ab33.log | sort -s -k2,2 -k1,1
}
ab
( SCG_LIB_NAME=libscg_b200_prof_ca.so SCG_MLP_ENGINE=ffma2 timeout 60 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=1)
" 2>&1 | grep -v "^$" | tail -12 ) >> gpurun_out/r2_prof_ticks33.log
grep "env-phase\|CTA0\|ms " gpurun_out/r2_prof_ticks33.log | cut -c1-330
SCG_LIB_NAME=libscg_b200_vca.so timeout 150 python bench.py > gpurun_out/bench_r02_c3_ca.json 2> gpurun_out/bench_r02_c3_ca.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_c3_ca.json').read().strip().splitlines()[-1]); print('CA bench', round(d['value']/1e6,1), 'M', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']/1e6,1))"
echo "== rep 2"
ab
echo done

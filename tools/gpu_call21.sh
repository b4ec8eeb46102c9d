#!/bin/bash
# per-warp env-phase timing (SCG_PROF), A/B of the main lib (MLP of the horizon loop inlined for cartpole / 1-D / 2-D) against round 1, bench lines
mkdir -p gpurun_out; rm -f gpurun_out/r2_ab21.log gpurun_out/r2_prof_ticks21.log; sed -i "s#x#x#" /dev/null

for rep in 1 2 3; do
for c in e7bb432 HEAD; do
  d=.; [ $c = e7bb432 ] && d=_bisect/e7bb432
  ( cd $d && echo "== $c" && SCG_MLP_ENGINE=ffma2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=25)
q.run('quadrotor_3D_track', 8192, 200, 64, 'f64', 32, iters=9)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=25)
q.run('quadrotor_2D_track', 4096, 200, 64, 'f32', 32, iters=25)
" 2>&1 | tail -5 ) >> gpurun_out/r2_ab21.log
done
done
cat gpurun_out/r2_ab21.log
timeout 400 python bench.py > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; tail -c 300 gpurun_out/bench_r02_c3.json
timeout 900 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q --tb=short > gpurun_out/r2_pytest_some.log 2>&1
tail -5 gpurun_out/r2_pytest_some.log
echo done

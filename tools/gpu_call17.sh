#!/bin/bash
# A/B of the compiled-in schedule (SCG_FM_FAST) against the previous head and the round-1 head on the same box, then the whole GPU suite
mkdir -p gpurun_out; rm -f gpurun_out/r2_fm_ab.log
for rep in 1 2 3; do
for c in e7bb432 f3b6836 HEAD; do
  d=_bisect/$c; [ $c = HEAD ] && d=.
  ( cd $d && echo "== $c" && SCG_MLP_ENGINE=ffma2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=25)
q.run('quadrotor_3D_track', 8192, 200, 64, 'f64', 32, iters=9)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=25)
" 2>&1 | tail -4 ) >> gpurun_out/r2_fm_ab.log
done
done
cat gpurun_out/r2_fm_ab.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -15 gpurun_out/r2_pytest_gpu_full.log
timeout 300 python bench.py --no-train > gpurun_out/bench_r02_c3_fm.json 2> gpurun_out/bench_r02_c3_fm.err
tail -c 400 gpurun_out/bench_r02_c3_fm.json
echo done

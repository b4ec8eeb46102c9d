#!/bin/bash
# warp-cooperative episode draws: per-warp env-phase timing, A/B against the round-1 head, whole GPU suite
mkdir -p gpurun_out; rm -f gpurun_out/r2_ab29.log gpurun_out/r2_prof_ticks29.log
( SCG_LIB_NAME=libscg_b200_prof.so SCG_MLP_ENGINE=ffma2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=1)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=1)
" 2>&1 | grep -v "^$" | tail -24 ) >> gpurun_out/r2_prof_ticks29.log
grep "env-phase\|CTA0\|ms " gpurun_out/r2_prof_ticks29.log | cut -c1-330
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
" 2>&1 | tail -5 ) >> gpurun_out/r2_ab29.log
done
done
awk '/^==/{c=$2} /N=/{print c, $1, $7, $(NF-5), $(NF-4)}' gpurun_out/r2_ab29.log | sort -s -k1,3
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -12 gpurun_out/r2_pytest_gpu_full.log
echo done

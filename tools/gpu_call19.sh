#!/bin/bash
# A/B: round-1 head, main lib (no clobber in multimem_st, relay always, cold paths out of line), MLP-inline variant; phase ticks; full GPU suite
mkdir -p gpurun_out; rm -f gpurun_out/r2_ab19.log gpurun_out/r2_prof_ticks19.log
for rep in 1 2 3; do
for c in e7bb432 HEAD vI; do
  d=.; lib=libscg_b200.so
  [ $c = e7bb432 ] && d=_bisect/e7bb432
  [ $c = vI ] && lib=libscg_b200_vI.so
  ( cd $d && echo "== $c" && SCG_LIB_NAME=$lib SCG_MLP_ENGINE=ffma2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=25)
q.run('quadrotor_3D_track', 8192, 200, 64, 'f64', 32, iters=9)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=25)
" 2>&1 | tail -4 ) >> gpurun_out/r2_ab19.log
done
done
cat gpurun_out/r2_ab19.log
( SCG_LIB_NAME=libscg_b200_prof.so SCG_MLP_ENGINE=ffma2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=1)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=1)
" 2>&1 | grep -v "^$" | tail -20 ) >> gpurun_out/r2_prof_ticks19.log
cat gpurun_out/r2_prof_ticks19.log | cut -c1-400
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -15 gpurun_out/r2_pytest_gpu_full.log
echo done

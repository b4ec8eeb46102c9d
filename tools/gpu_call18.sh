#!/bin/bash
# A/B of the cold-path outlining + acquire/release flags (main lib) and of the relay-always variant against the round-1 head; phase ticks of
# both (-DSCG_PROF builds); then the tests that failed in the previous call
mkdir -p gpurun_out; rm -f gpurun_out/r2_cold_ab.log gpurun_out/r2_prof_ticks.log
for rep in 1 2 3; do
for c in e7bb432 HEAD vR; do
  d=.; lib=libscg_b200.so
  [ $c = e7bb432 ] && d=_bisect/e7bb432
  [ $c = vR ] && lib=libscg_b200_vR.so
  ( cd $d && echo "== $c" && SCG_LIB_NAME=$lib SCG_MLP_ENGINE=ffma2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=25)
q.run('quadrotor_3D_track', 8192, 200, 64, 'f64', 32, iters=9)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=25)
" 2>&1 | tail -4 ) >> gpurun_out/r2_cold_ab.log
done
done
cat gpurun_out/r2_cold_ab.log
for d in _bisect/e7bb432 .; do
  ( cd $d && echo "== $d" && SCG_LIB_NAME=libscg_b200_prof.so SCG_MLP_ENGINE=ffma2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=3)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=3)
" 2>&1 | grep -v "^$" | tail -40 ) >> gpurun_out/r2_prof_ticks.log
done
cat gpurun_out/r2_prof_ticks.log | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_extended.py tests/test_gpu_safe_explorer.py tests/test_gpu_parity.py tests/test_gpu_bench_parity.py -m gpu -q --tb=short > gpurun_out/r2_pytest_ext.log 2>&1
tail -25 gpurun_out/r2_pytest_ext.log
echo done

#!/bin/bash
# A/B: round-1 head, main lib (one unrolled layer-1 variant per system, peer loops not unrolled, sin/cos theta carried in the 2-D integrator state,
# blocked 2-D sub-step loop), vS = the same without the blocked loop; then the whole GPU suite
mkdir -p gpurun_out; rm -f gpurun_out/r2_ab23.log
for rep in 1 2 3; do
for c in e7bb432 HEAD vS; do
  d=.; lib=libscg_b200.so
  [ $c = e7bb432 ] && d=_bisect/e7bb432
  [ $c = vS ] && lib=libscg_b200_vS.so
  ( cd $d && echo "== $c" && SCG_LIB_NAME=$lib SCG_MLP_ENGINE=ffma2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=25)
q.run('quadrotor_3D_track', 8192, 200, 64, 'f64', 32, iters=9)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=25)
q.run('quadrotor_2D_track', 4096, 200, 64, 'f32', 32, iters=25)
" 2>&1 | tail -5 ) >> gpurun_out/r2_ab23.log
done
done
cat gpurun_out/r2_ab23.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -15 gpurun_out/r2_pytest_gpu_full.log
echo done

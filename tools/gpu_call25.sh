#!/bin/bash
# A/B of single switches against the main lib (running-sum 2-D sub-step loop, carried sin/cos, GAE prefetch): vC = no carried sin/cos in the cart-pole,
# vG = no GAE prefetch, vQ = the 16-operation 2-D sub-step loop; round-1 head as the anchor; then the parity suites on the main lib
mkdir -p gpurun_out; rm -f gpurun_out/r2_ab25.log
for rep in 1 2 3; do
for c in e7bb432 HEAD vC vG vQ; do
  d=.; lib=libscg_b200.so
  [ $c = e7bb432 ] && d=_bisect/e7bb432
  [ $c != e7bb432 ] && [ $c != HEAD ] && lib=libscg_b200_$c.so
  ( cd $d && echo "== $c" && SCG_LIB_NAME=$lib SCG_MLP_ENGINE=ffma2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=25)
q.run('quadrotor_3D_track', 8192, 200, 64, 'f64', 32, iters=9)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=25)
" 2>&1 | tail -4 ) >> gpurun_out/r2_ab25.log
done
done
awk '/^==/{c=$2} /N=/{print c, $1, $7, $(NF-5), $(NF-4)}' gpurun_out/r2_ab25.log | sort -s -k1,3
timeout 900 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_parity.py tests/test_gpu_surface_parity.py tests/test_gpu_extended.py -m gpu -q --tb=short > gpurun_out/r2_pytest_some.log 2>&1
tail -6 gpurun_out/r2_pytest_some.log
echo done

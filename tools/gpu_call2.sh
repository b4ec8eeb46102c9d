#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/mlp_tc2 > gpurun_out/r2_ubench_tc2.log 2>&1
python -m pytest tests/test_gpu_bench_parity.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_pytest_bench_parity.log
python -m pytest tests -m gpu -q --deselect tests/test_gpu_bench_parity.py 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu_rest.log
for lib in libscg_b200.so libscg_b200_dev.so; do
SCG_LIB_NAME=$lib python - >> gpurun_out/r2_relay.log 2>&1 <<'PY'
import sys, os; sys.path.insert(0, 'tools'); sys.path.insert(0, 'tests')
import quick_bench as q
print(os.environ['SCG_LIB_NAME'])
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=9)
q.run('cartpole_c2', 4096, 200, 64, 'f64', 32, iters=9)
q.run('quadrotor_3D_track_c4', 8192, 200, 64, 'f64', 32, iters=5)
PY
done
echo done

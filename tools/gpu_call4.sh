#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/mlp_tc2 > gpurun_out/r2_ubench_tc2.log 2>&1
for eng in ffma2 tcgen05; do
SCG_MLP_ENGINE=$eng timeout 300 python - >> gpurun_out/r2_tc_timing.log 2>&1 <<'PY'
import sys, os; sys.path.insert(0, 'tools'); sys.path.insert(0, 'tests')
import quick_bench as q
print(os.environ['SCG_MLP_ENGINE'])
for prec in ('f64', 'f32'):
    q.run('quadrotor_2D_track', 4096, 200, 64, prec, 0, iters=9)
q.run('cartpole_c2', 4096, 200, 64, 'f64', 0, iters=9)
q.run('quadrotor_3D_track_c4', 8192, 200, 64, 'f64', 0, iters=5)
q.run('quadrotor_2D_track', 32768, 200, 64, 'f64', 0, iters=5)
PY
done
timeout 900 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_surface_parity.py -m gpu -q -s --tb=line -rA 2>&1 | grep -E "Assertion|largest|forks|PASS|FAIL|passed|failed|Error|error" | cut -c1-900 > gpurun_out/r2_pytest_bench_parity.log
echo done

#!/bin/bash
mkdir -p gpurun_out
for eng in ffma2 tcgen05; do
SCG_MLP_ENGINE=$eng timeout 300 python - >> gpurun_out/r2_tc128_timing.log 2>&1 <<'PY'
import sys, os; sys.path.insert(0, 'tools'); sys.path.insert(0, 'tests')
import quick_bench as q
print(os.environ['SCG_MLP_ENGINE'])
q.run('quadrotor_2D_track', 4096, 200, 128, 'f64', 0, iters=7)
q.run('quadrotor_2D_track', 4096, 200, 128, 'f32', 0, iters=7)
q.run('cartpole_c2', 4096, 200, 128, 'f64', 0, iters=7)
q.run('quadrotor_2D_track', 32768, 200, 128, 'f64', 0, iters=3)
PY
done
timeout 1200 python -m pytest tests -m gpu -q -s --tb=line -rA 2>&1 | grep -E "Assertion|largest|forks|PASS|FAIL|passed|failed|Error|error|SKIP" | cut -c1-700 > gpurun_out/r2_pytest_all.log
echo done

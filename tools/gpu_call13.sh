#!/bin/bash
# one GPU: timing regression check of the default collector after the SafeExplorer / normaliser instantiation split, then the whole GPU suite
mkdir -p gpurun_out
SCG_MLP_ENGINE=ffma2 timeout 200 python - > gpurun_out/r2_regress.log 2>&1 <<'PY'
import sys, os; sys.path.insert(0, 'tools'); sys.path.insert(0, 'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 0, iters=15)
q.run('quadrotor_3D_track_c4', 8192, 200, 64, 'f64', 0, iters=7)
q.run('cartpole_c2', 4096, 200, 64, 'f64', 0, iters=15)
PY
timeout 600 python -m pytest tests/test_gpu_safe_explorer.py -m gpu -q -s --tb=short > gpurun_out/r2_pytest_safe.log 2>&1
tail -5 gpurun_out/r2_pytest_safe.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_gpu_safe_explorer.py > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -8 gpurun_out/r2_pytest_gpu_full.log
echo done

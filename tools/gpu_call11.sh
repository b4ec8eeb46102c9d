#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2_regress.log
for v in A C A C; do
SCG_LIB_NAME=libscg_b200_v$v.so SCG_MLP_ENGINE=ffma2 timeout 300 python - >> gpurun_out/r2_regress.log 2>&1 <<'PY'
import sys, os; sys.path.insert(0, 'tools'); sys.path.insert(0, 'tests')
import quick_bench as q
print(os.environ['SCG_LIB_NAME'])
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 0, iters=15)
q.run('quadrotor_3D_track_c4', 8192, 200, 64, 'f64', 0, iters=7)
PY
done
SCG_LIB_NAME=libscg_b200_dev.so timeout 300 python tools/bench_update.py 131072 64 1 5 > gpurun_out/r2_ppo_phases.log 2>&1
timeout 300 python tools/bench_update.py 819200 64 1 20 > gpurun_out/r2_bench_update.log 2>&1
timeout 300 python tools/bench_update.py 819200 256 1 20 >> gpurun_out/r2_bench_update.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ppo_update.py tests/test_normalizers.py tests/test_gpu_parity.py -m gpu -q --tb=line 2>&1 | tail -8 > gpurun_out/r2_pytest_some.log
echo done

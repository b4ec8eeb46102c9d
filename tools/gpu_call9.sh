#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_tc64f_timing.log
for eng in ffma2 tcgen05 tcgen05_tf32; do
SCG_MLP_ENGINE=$eng timeout 300 python - >> gpurun_out/r2_tc64f_timing.log 2>&1 <<'PY'
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
timeout 1500 python -m pytest tests -m gpu -q -s --tb=line -rA 2>&1 | grep -E "^E |Assertion|PASS|FAIL|passed|failed|Error|error|SKIP" | cut -c1-500 > gpurun_out/r2_pytest_all.log

SCG_LIB_NAME=libscg_b200_dev.so timeout 300 python tools/bench_update.py 131072 64 1 5 > gpurun_out/r2_ppo_phases.log 2>&1
SCG_LIB_NAME=libscg_b200_dev.so SCG_MLP_ENGINE=tcgen05 timeout 300 python tools/profile_one.py quadrotor_2D_track 4096 64 f64 0 2 > gpurun_out/r2_tc_phases.log 2>&1
echo done

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ppo_update.py tests/test_gpu_api.py -m gpu -q --tb=short -rA 2>&1 | grep -E "^E |PASS|FAIL|passed|failed|Error" | cut -c1-600 > gpurun_out/r2_pytest_ppo.log
timeout 300 python tools/bench_update.py 819200 64 1 100 > gpurun_out/r2_bench_update.log 2>&1
SCG_PPO_NO_CLUSTER=1 timeout 300 python tools/bench_update.py 819200 64 1 100 >> gpurun_out/r2_bench_update.log 2>&1
timeout 300 python tools/bench_update.py 819200 256 1 50 >> gpurun_out/r2_bench_update.log 2>&1
timeout 300 python tools/bench_update.py 819200 32768 2 2 >> gpurun_out/r2_bench_update.log 2>&1
echo done

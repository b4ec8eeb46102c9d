#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/mlp_tc2 > gpurun_out/r2_ubench_tc2.log 2>&1
python -m pytest tests/test_gpu_bench_parity.py -m gpu -q -s --tb=line -rA 2>&1 | grep -E "Assertion|largest|forks|PASS|FAIL|passed|failed|Error" | cut -c1-700 > gpurun_out/r2_pytest_bench_parity.log
echo done

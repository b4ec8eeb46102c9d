#!/bin/bash
# TWO GPUs: why do the replicas diverge?  the lock-step test with (1) the main lib, (2) unicast peer stores, (3) the memory clobber back on multimem.st
mkdir -p gpurun_out; rm -f gpurun_out/r2_pytest_2gpu_*.log
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -q --tb=short -s > gpurun_out/r2_pytest_2gpu_main.log 2>&1; tail -3 gpurun_out/r2_pytest_2gpu_main.log | cut -c1-200
SCG_NO_MULTICAST=1 timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -q --tb=short -s > gpurun_out/r2_pytest_2gpu_unicast.log 2>&1; tail -3 gpurun_out/r2_pytest_2gpu_unicast.log | cut -c1-200
SCG_LIB_NAME=libscg_b200_vM.so timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -q --tb=short -s > gpurun_out/r2_pytest_2gpu_clobber.log 2>&1; tail -3 gpurun_out/r2_pytest_2gpu_clobber.log | cut -c1-200
grep -h "AssertionError: (" gpurun_out/r2_pytest_2gpu_*.log | cut -c1-300 | head
echo done

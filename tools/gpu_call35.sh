#!/bin/bash
# RAP controller (composition of replayed launches) on the GPU: the one test that was still unverified
mkdir -p gpurun_out
SCG_RUN_UNVERIFIED=1 timeout 45 python -m pytest tests/test_gpu_extended.py -q -k rap --tb=short -x -m gpu > gpurun_out/r2_pytest_rap.log 2>&1
tail -25 gpurun_out/r2_pytest_rap.log | cut -c1-250
echo done

#!/bin/bash
# bisect timing (tools/gpu_call15.sh), then the new extended-surface tests + the SafeExplorer tests with the conditioning-aware bound
bash tools/gpu_call15.sh > /dev/null 2>&1
timeout 900 python -m pytest tests/test_gpu_extended.py tests/test_gpu_safe_explorer.py -m gpu -q --tb=short -x -s > gpurun_out/r2_pytest_ext.log 2>&1
tail -40 gpurun_out/r2_pytest_ext.log
echo done

#!/bin/bash
# TWO GPUs: lock-step replicas + fused gather tests, then the bench at N=2 (device-timed, e2e, train_e2e)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q --tb=short -s > gpurun_out/r2_pytest_2gpu.log 2>&1
tail -12 gpurun_out/r2_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_c3_n2.json 2> gpurun_out/bench_r02_c3_n2.err
tail -c 1500 gpurun_out/bench_r02_c3_n2.json; tail -5 gpurun_out/bench_r02_c3_n2.err
./tools/ubench/fp64_lanes > gpurun_out/ubench_fp64_lanes.log 2>&1; cat gpurun_out/ubench_fp64_lanes.log
echo done

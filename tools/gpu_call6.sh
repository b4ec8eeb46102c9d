#!/bin/bash
# 2-GPU call: lock-step replica test (lagging rank), bench at N=2 with multicast and with peer stores
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_n2.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s --tb=short 2>&1 | tail -30 >> gpurun_out/r2_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_mc.json 2> gpurun_out/r2_bench_n2_mc.err
SCG_NO_MULTICAST=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 --no-train > gpurun_out/r2_bench_n2_peer.json 2> gpurun_out/r2_bench_n2_peer.err
echo done

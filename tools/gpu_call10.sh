#!/bin/bash
# 8-GPU call: scaling of the fused rollout + in-kernel gather (multicast vs peer stores), e2e with NUMA-bound pinned buffers, lock-step test
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_n8_topo.log 2>&1
CUDA_VISIBLE_DEVICES=0 SCG_MLP_ENGINE=ffma2 timeout 200 python - > gpurun_out/r2_regress.log 2>&1 <<'PY'
import sys, os; sys.path.insert(0, 'tools'); sys.path.insert(0, 'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 0, iters=15)
q.run('quadrotor_3D_track_c4', 8192, 200, 64, 'f64', 0, iters=7)
PY
timeout 300 python -m pytest tests/test_gpu_multirank.py -m gpu -q -s --tb=short 2>&1 | tail -40 > gpurun_out/r2_n8_lockstep.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_mc.json 2> gpurun_out/r2_bench_n8_mc.err
SCG_NO_MULTICAST=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 8 --steps 20 --warmup 5 --no-train > gpurun_out/r2_bench_n8_peer.json 2> gpurun_out/r2_bench_n8_peer.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 4 --steps 20 --warmup 5 --no-train > gpurun_out/r2_bench_n4_mc.json 2> gpurun_out/r2_bench_n4_mc.err
echo done

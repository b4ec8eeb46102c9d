#!/bin/bash
# one GPU, re-entry of round 2: the evidence of record (bench lines of every config, launch list, full-set captures), then the GPU suite
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 400 python bench.py > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err
tail -c 600 gpurun_out/bench_r02_c3.json
timeout 300 python bench.py --config c2 --no-train > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err
timeout 300 python bench.py --config c4 --no-train --steps 10 > gpurun_out/bench_r02_c4.json 2> gpurun_out/bench_r02_c4.err
timeout 300 python bench.py --hidden 128 --no-train --no-cpu-baseline --steps 10 > gpurun_out/bench_r02_c3_h128.json 2> gpurun_out/bench_r02_c3_h128.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_c3_reference.json 2> gpurun_out/bench_r02_c3_reference.err
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 1 > gpurun_out/bench_under_ncu_r02.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_c3 python tools/profile_one.py quadrotor_2D_track 4096 64 f64 0 4 > gpurun_out/ncu_r02_c3.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:rollout_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_tc128 python tools/profile_one.py quadrotor_2D_track 4096 128 f64 0 4 > gpurun_out/ncu_r02_tc128.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:ppo_update_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02_ppo_cluster python tools/bench_update.py 65536 64 1 2 > gpurun_out/ncu_r02_ppo_cluster.log 2>&1
timeout 300 python tools/bench_update.py 819200 64 1 20 > gpurun_out/r2_bench_update.log 2>&1
timeout 300 python tools/bench_update.py 819200 32768 10 3 >> gpurun_out/r2_bench_update.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -8 gpurun_out/r2_pytest_gpu_full.log
echo done

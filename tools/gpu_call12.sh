#!/bin/bash
# ncu evidence of round 2 (one GPU): launch list of the bench command + full-set captures of the dominant kernels
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 900 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 1 > gpurun_out/bench_under_ncu_r02.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_c3 python tools/profile_one.py quadrotor_2D_track 4096 64 f64 0 4 > gpurun_out/ncu_r02_c3.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:rollout_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_tc128 python tools/profile_one.py quadrotor_2D_track 4096 128 f64 0 4 > gpurun_out/ncu_r02_tc128.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:ppo_update_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02_ppo_cluster python tools/bench_update.py 65536 64 1 2 > gpurun_out/ncu_r02_ppo_cluster.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:ppo_update_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02_ppo_grid python tools/bench_update.py 819200 32768 1 1 > gpurun_out/ncu_r02_ppo_grid.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done

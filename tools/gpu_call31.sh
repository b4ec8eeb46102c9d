#!/bin/bash
# one GPU: the evidence of record of round 2 (final kernels): smoke, bench lines, launch list, full-set captures, whole GPU suite
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke_r02.log 2>&1; tail -2 gpurun_out/smoke_r02.log
timeout 400 python bench.py > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err
timeout 300 python bench.py --config c2 --no-train > gpurun_out/bench_r02_c2.json 2> gpurun_out/bench_r02_c2.err
timeout 300 python bench.py --config c4 --no-train --steps 10 > gpurun_out/bench_r02_c4.json 2> gpurun_out/bench_r02_c4.err
timeout 300 python bench.py --hidden 128 --no-train --no-cpu-baseline --steps 10 > gpurun_out/bench_r02_c3_h128.json 2> gpurun_out/bench_r02_c3_h128.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_c3_reference.json 2> gpurun_out/bench_r02_c3_reference.err
for f in c3 c2 c4 c3_h128 c3_reference; do python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_$f.json').read().strip().splitlines()[-1]); print('$f', round(d['value']/1e6,2), 'M', round(d['ms_per_step'],4), 'ms  e2e', round(d['e2e']['value']/1e6,1), d['e2e'].get('update_keys',{}).get('value'))"; done
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 1 > gpurun_out/bench_under_ncu_r02.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_c3 python tools/profile_one.py quadrotor_2D_track 4096 64 f64 0 4 > gpurun_out/ncu_r02_c3.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:rollout_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_r02_tc128 python tools/profile_one.py quadrotor_2D_track 4096 128 f64 0 4 > gpurun_out/ncu_r02_tc128.log 2>&1
timeout 400 $NCU --set full --import-source on -k regex:ppo_update_kernel -s 1 -c 1 -f -o gpurun_out/prof_r02_ppo_cluster python tools/bench_update.py 65536 64 1 2 > gpurun_out/ncu_r02_ppo_cluster.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r2_pytest_gpu_full.log 2>&1
tail -6 gpurun_out/r2_pytest_gpu_full.log
echo done

#!/bin/bash
# GPU call 1 of round 2: full GPU test-suite (incl. the new bench-size parity tests), variant timings, phase profile, bench lines.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest_gpu.log
python - > gpurun_out/r2_variants.log 2>&1 <<'PY'
import sys; sys.path.insert(0, 'tools'); sys.path.insert(0, 'tests')
import quick_bench as q
for prec in ('f64', 'f32'):
    for epb in (32, 16):
        q.run('quadrotor_2D_track', 4096, 200, 64, prec, epb)
q.run('quadrotor_2D_track', 32768, 200, 64, 'f64', 32)
q.run('cartpole_c2', 4096, 200, 64, 'f64', 32)
q.run('quadrotor_3D_track_c4', 8192, 200, 64, 'f64', 32)
q.run('quadrotor_3D_track_c4', 8192, 200, 64, 'f64', 16)
q.run('quadrotor_2D_track', 4096, 200, 128, 'f64', 16)
q.run('quadrotor_2D_track', 4096, 200, 128, 'f64', 8)
PY
SCG_LIB_NAME=libscg_b200_prof.so python tools/profile_one.py quadrotor_2D_track 4096 64 f64 32 3 > gpurun_out/r2_prof_phases.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err
python bench.py --config c2 --steps 10 --warmup 3 --no-train > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
python bench.py --config c4 --steps 10 --warmup 3 --no-train > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo done

#!/bin/bash
# timing bisect of the default collector (C3, hidden 64, f64): every round-2 commit's own tree + library, same box, interleaved twice
mkdir -p gpurun_out; rm -f gpurun_out/r2_bisect.log
for rep in 1 2; do
for c in e7bb432 481e089 de1206e 3727f8c 18fe20f 9f09df2 52a0c09 f3b6836; do
  ( cd _bisect/$c && echo "== $c" && SCG_MLP_ENGINE=ffma2 timeout 120 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=25)
q.run('quadrotor_3D_track', 8192, 200, 64, 'f64', 32, iters=9)
" 2>&1 | tail -3 ) >> gpurun_out/r2_bisect.log
done
done
cat gpurun_out/r2_bisect.log

#!/bin/bash
# phase ticks: blocked 2-D sub-step loop (profB) against the sequential one (profS), both with sin/cos theta carried
mkdir -p gpurun_out; rm -f gpurun_out/r2_prof_ticks24.log
for lib in libscg_b200_profB.so libscg_b200_profS.so; do
( echo "== $lib"; SCG_LIB_NAME=$lib SCG_MLP_ENGINE=ffma2 timeout 200 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=1)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=1)
" 2>&1 | grep -v "^$" | tail -24 ) >> gpurun_out/r2_prof_ticks24.log
done
grep "==\|CTA0\|ticks/step\|ms " gpurun_out/r2_prof_ticks24.log | cut -c1-330
echo done

#!/bin/bash
# final check of the rebuilt default library (decoupled critic compiled out): smoke, same-box A/B against the build the table of record was measured with
mkdir -p gpurun_out; rm -f gpurun_out/r2_ab34.log
( echo "== rebuilt" && SCG_MLP_ENGINE=ffma2 timeout 110 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=15)
q.run('quadrotor_3D_track', 8192, 200, 64, 'f64', 32, iters=5)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=15)
import os; os.environ.pop('SCG_MLP_ENGINE', None)
import __graft_entry__ as g; g.smoke(); print('SMOKE OK')
" 2>&1 | tail -5 ) >> gpurun_out/r2_ab34.log
( echo "== measured" && SCG_LIB_NAME=libscg_b200_vprev.so SCG_MLP_ENGINE=ffma2 timeout 60 python -c "
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'tests')
import quick_bench as q
q.run('quadrotor_2D_track', 4096, 200, 64, 'f64', 32, iters=15)
q.run('quadrotor_3D_track', 8192, 200, 64, 'f64', 32, iters=5)
q.run('cartpole', 4096, 200, 64, 'f64', 32, iters=15)
" 2>&1 | tail -3 ) >> gpurun_out/r2_ab34.log
cut -c1-200 gpurun_out/r2_ab34.log
echo done

#!/bin/bash
mkdir -p gpurun_out
SCG_LIB_NAME=libscg_b200_dev.so SCG_MLP_ENGINE=tcgen05 python tools/profile_one.py quadrotor_2D_track 4096 64 f64 0 3 > gpurun_out/r2_tc_phases.log 2>&1
SCG_LIB_NAME=libscg_b200_dev.so SCG_MLP_ENGINE=tcgen05 python tools/profile_one.py cartpole_c2 4096 64 f64 0 2 >> gpurun_out/r2_tc_phases.log 2>&1
echo done

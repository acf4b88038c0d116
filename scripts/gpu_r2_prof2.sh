#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:solve_fused_f2 -s 1 -c 1 -o gpurun_out/r2_prof_f2_final -f python scripts/profile_kernels.py c3small > gpurun_out/r2_ncu_f2.log 2>&1; tail -1 gpurun_out/r2_ncu_f2.log
ncu --set full --clock-control none -k regex:solve_fused_f2 -s 1 -c 1 -o gpurun_out/r2_prof_f2_final_2p24 -f python scripts/profile_kernels.py c3 > gpurun_out/r2_ncu_f2b.log 2>&1; tail -1 gpurun_out/r2_ncu_f2b.log

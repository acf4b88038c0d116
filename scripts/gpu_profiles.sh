#!/bin/bash
# ncu --set full captures of the kernels besides the default workload's (see gpu_checks.sh for that one)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:solve_fused -s 1 -c 1 -o gpurun_out/prof_fused_c3 -f python scripts/profile_kernels.py c3small > gpurun_out/ncu_fused_c3.log 2>&1
tail -1 gpurun_out/ncu_fused_c3.log
ncu --set full --clock-control none --import-source on -k regex:mlp_tanh256 -s 20 -c 1 -o gpurun_out/prof_mlp -f python scripts/profile_kernels.py c4 > gpurun_out/ncu_mlp.log 2>&1
tail -1 gpurun_out/ncu_mlp.log

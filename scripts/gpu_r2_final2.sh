#!/bin/bash
# ncu capture of the step-fused MLP launch after the late round-2 changes, then the round-end validation
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mlp_tanh256 -s 8 -c 1 -o gpurun_out/r2_prof_mlp_step_late -f python scripts/profile_kernels.py c4 > gpurun_out/r2_ncu_mlp.log 2>&1; tail -1 gpurun_out/r2_ncu_mlp.log
python scripts/mlp_step_timing.py
bash scripts/gpu_r2_final.sh

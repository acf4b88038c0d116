#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused_f2.py -x -q > gpurun_out/r2_test_f2.log 2>&1; echo "f2 tests rc=$?"; tail -3 gpurun_out/r2_test_f2.log
echo "=== default"; python scripts/f2_timing.py 2>&1 | tail -1
for lib in build_variants/f2_*.so; do
  echo "=== $lib"; TORCHODE_B200_LIB=$PWD/$lib python scripts/f2_timing.py 2>&1 | tail -1
done
ncu --set full --clock-control none --import-source on -k regex:solve_fused_f2 -s 1 -c 1 -o gpurun_out/r2_prof_f2_c3_v4 -f python scripts/profile_kernels.py c3small > gpurun_out/r2_ncu_f2.log 2>&1
tail -1 gpurun_out/r2_ncu_f2.log

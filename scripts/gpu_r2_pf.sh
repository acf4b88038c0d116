#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused_f2.py -m gpu -x -q 2>&1 | tail -2
echo "=== default (prefetching refill)"; python scripts/f2_timing.py 2>&1 | tail -1
for lib in build_variants/f2_*.so; do
  echo "=== $lib"; TORCHODE_B200_LIB=$PWD/$lib python scripts/f2_timing.py 2>&1 | tail -1
done
echo "=== default again"; python scripts/f2_timing.py 2>&1 | tail -1

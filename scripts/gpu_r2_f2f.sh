#!/bin/bash
cd "$(dirname "$0")/.."
echo "=== default (4 CTAs/SM, refill 8)"; python scripts/f2_timing.py 2>&1 | tail -1
echo "=== grid for 3 CTAs/SM"; TODE_F2_CTAS=3 python scripts/f2_timing.py 2>&1 | tail -1
echo "=== grid for 2 CTAs/SM"; TODE_F2_CTAS=2 python scripts/f2_timing.py 2>&1 | tail -1
for lib in build_variants/f2_*.so; do
  echo "=== $lib"; TORCHODE_B200_LIB=$PWD/$lib python scripts/f2_timing.py 2>&1 | tail -1
done

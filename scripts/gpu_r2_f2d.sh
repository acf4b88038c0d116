#!/bin/bash
cd "$(dirname "$0")/.."
echo "=== default"; python scripts/f2_timing.py 2>&1 | tail -1
for lib in build_variants/f2_*.so; do
  echo "=== $lib"; TORCHODE_B200_LIB=$PWD/$lib python scripts/f2_timing.py 2>&1 | tail -1
done

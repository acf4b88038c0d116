cd /root/repo
echo "=== default"; python scripts/heat_step_timing.py 2>&1 | tail -3
for lib in build_variants/*.so; do
  [ -f $lib ] || continue
  echo "=== $lib"
  TORCHODE_B200_LIB=$PWD/$lib python scripts/heat_step_timing.py 2>&1 | tail -3
done

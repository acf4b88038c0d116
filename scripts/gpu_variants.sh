cd /root/repo
echo "=== default"; for i in 1 2; do python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C2"; done
for lib in build_variants/*.so; do
  [ -f $lib ] || continue
  echo "=== $lib"
  for i in 1 2; do TORCHODE_B200_LIB=$PWD/$lib python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C2"; done
done

cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== default lib"
python scripts/profile_kernels.py path_a_all 2>&1 | grep -E "finish|NK=1>|NK=6"
for lib in build_variants/*.so; do
  echo "=== $lib"
  TORCHODE_B200_LIB=$PWD/$lib python scripts/profile_kernels.py path_a_all 2>&1 | grep -E "finish"
done

cd /root/repo
echo "=== default (96 B in flight)"; python scripts/profile_kernels.py path_a_all 2>&1 | grep -E "stage"
for lib in build_variants/stage_*.so; do echo "=== $lib"; TORCHODE_B200_LIB=$PWD/$lib python scripts/profile_kernels.py path_a 2>&1 | grep -E "stage"; done

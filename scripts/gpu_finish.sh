cd /root/repo
echo "=== default"; python scripts/profile_kernels.py path_a_all 2>&1 | grep -E "finish"
for lib in build_variants/finish_*.so; do echo "=== $lib"; TORCHODE_B200_LIB=$PWD/$lib python scripts/profile_kernels.py path_a_all 2>&1 | grep -E "finish"; done

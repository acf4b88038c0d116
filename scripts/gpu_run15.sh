cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== default"; for i in 1 2 3; do python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C"; done
for lib in build_variants/*.so; do
  [ -f $lib ] || continue
  echo "=== $lib"
  for i in 1 2; do TORCHODE_B200_LIB=$PWD/$lib python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C2"; done
done
ncu --set full --clock-control none --import-source on -k regex:solve_fused -s 1 -c 1 -o gpurun_out/prof_fused_c2_v3 -f python scripts/profile_kernels.py c2 > gpurun_out/ncu_fused_v3.log 2>&1
tail -2 gpurun_out/ncu_fused_v3.log

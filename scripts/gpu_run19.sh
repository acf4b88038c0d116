cd /root/repo
python -m pytest tests -m gpu -q 2>&1 | tail -6
for i in 1 2; do python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C"; done
ncu --set full --clock-control none --import-source on -k regex:solve_fused -s 1 -c 1 -o gpurun_out/prof_fused_c2_v4 -f python scripts/profile_kernels.py c2 > gpurun_out/ncu_fused_v4.log 2>&1
tail -2 gpurun_out/ncu_fused_v4.log

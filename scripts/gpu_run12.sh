cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -12
ncu --set full --clock-control none --import-source on -k regex:solve_fused -s 1 -c 1 -o gpurun_out/prof_fused_c3 -f python scripts/profile_kernels.py c3small > gpurun_out/ncu_fused_c3.log 2>&1
tail -3 gpurun_out/ncu_fused_c3.log
for i in 1 2 3; do python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C3"; done

cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for i in 1 2 3; do python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C"; done
python scripts/profile_kernels.py path_a 2>&1 | grep -E "finish"

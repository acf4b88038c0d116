cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python scripts/profile_kernels.py path_a_all 2>&1 | grep -E "finish|NK=6" | tail -8
python scripts/quick_time.py staged 2>&1 | tail -12

cd /root/repo
python -m pytest tests/test_gpu_backsolve.py -x -q 2>&1 | tail -25
python -m pytest tests -m gpu -x -q 2>&1 | tail -3

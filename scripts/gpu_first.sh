set -x
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -30
python scripts/quick_time.py 2>&1 | tail -20

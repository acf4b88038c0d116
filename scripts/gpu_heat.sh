cd /root/repo
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "heat" 2>&1 | tail -4
python bench.py --workload c5 --steps 3 --warmup 3 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c5', d['ms_per_step'], d['value'], d['route'], d['roofline']['achieved'], d['roofline']['frac'], d.get('cpu_baseline'))"

cd /root/repo
python -m pytest tests/test_gpu_kernels.py -x -q -k "fast_scalar" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for i in 1 2 3; do python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C"; done
python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/bench_c2_run13.json 2> gpurun_out/bench_c2_run13.err; cat gpurun_out/bench_c2_run13.json | python -c "import sys,json; d=json.load(sys.stdin); print(d['ms_per_step'], d['value'], d['e2e'])"

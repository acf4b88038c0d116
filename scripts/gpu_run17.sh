cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C"; done
python bench.py > gpurun_out/bench_c2_run17.json 2> gpurun_out/bench_c2_run17.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_run17.json')); print(d['ms_per_step'], d['value'], d['e2e'], d.get('fp64_issue'), d.get('cpu_baseline'))"

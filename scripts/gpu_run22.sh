cd /root/repo
python -m pytest tests -m gpu -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_c2_run22.json 2> gpurun_out/bench_c2_run22.err; tail -2 gpurun_out/bench_c2_run22.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_run22.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])"

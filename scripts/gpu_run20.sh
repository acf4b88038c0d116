cd /root/repo
python -m pytest tests -m gpu -q 2>&1 | tail -3
for i in 1 2; do python scripts/quick_time.py fusedonly 2>&1 | grep -E "^C"; done
ncu --set full --clock-control none --import-source on -k regex:solve_fused -s 1 -c 1 -o gpurun_out/prof_fused_c2_v5 -f python scripts/profile_kernels.py c2 > gpurun_out/ncu_fused_v5.log 2>&1
tail -1 gpurun_out/ncu_fused_v5.log
python bench.py > gpurun_out/bench_c2_run20.json 2> gpurun_out/bench_c2_run20.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_run20.json')); print(d['ms_per_step'], d['value'], d['e2e'], d.get('fp64_issue'), d.get('cpu_baseline'))"
python bench.py --impl reference --steps 2 --warmup 1 | cut -c1-300

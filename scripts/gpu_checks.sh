#!/bin/bash
# What a round's GPU validation runs (one B200): tests, smoke, the bench lines of every workload,
# launch list + full ncu capture of the dominant kernel.  Usage (from the repo root, on a GPU box):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_checks.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for w in c1 c3 c4 c5; do
  python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
python bench.py --workload c3 --batch 1048576 --steps 5 --warmup 3 > gpurun_out/bench_c3_2p20.json 2> gpurun_out/bench_c3_2p20.err
for w in c2 c1 c3 c3_2p20 c4 c5; do
  python -c "
import json; d=json.load(open('gpurun_out/bench_$w.json')); e=d['e2e']
print('$w', 'ms/step %.3f' % d['ms_per_step'], 'value %.4g' % d['value'], 'e2e %.4g (%.3f ms)' % (e['value'], e['ms_per_step']), d['route'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_c2.csv \
  python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_fused -s 1 -c 1 -o gpurun_out/prof_fused_c2 -f \
  python scripts/profile_kernels.py c2 > gpurun_out/ncu_fused.log 2>&1
tail -1 gpurun_out/ncu_fused.log
# configs[4]: launch list of one step-fused solve and a full capture of heat_step_kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_c5.csv \
  python scripts/profile_kernels.py c5 > gpurun_out/ncu_c5_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:heat_step -s 10 -c 1 -o gpurun_out/prof_heat_step -f \
  python scripts/profile_kernels.py c5 > gpurun_out/ncu_heat_step.log 2>&1
tail -1 gpurun_out/ncu_heat_step.log

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -x -q -k "heat" 2>&1 | tail -3
python -m pytest tests/test_gpu_vs_reference.py -m gpu -x -q -s 2>&1 | grep -E "C4 B|passed|failed"
python bench.py --workload c5 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c5.json"))
print("c5 ms", round(d["ms_per_step"], 3), "frac", round(d["roofline"]["frac"], 3), d["route"])
PY
ncu --set full --clock-control none --import-source on -k regex:heat_step_kernel -s 5 -c 1 -o gpurun_out/r2_prof_heat_v5 -f python scripts/profile_kernels.py c5 > gpurun_out/r2_ncu_heat.log 2>&1
tail -1 gpurun_out/r2_ncu_heat.log

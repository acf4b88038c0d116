#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mlp_tanh256 -s 8 -c 1 -o gpurun_out/r2_prof_mlp_step -f python scripts/profile_kernels.py c4 > gpurun_out/r2_ncu_mlp.log 2>&1; tail -1 gpurun_out/r2_ncu_mlp.log
ncu --set full --clock-control none -k regex:fused_f2_init -s 1 -c 1 -o gpurun_out/r2_prof_f2_init -f python scripts/profile_kernels.py c3small > gpurun_out/r2_ncu_f2i.log 2>&1; tail -1 gpurun_out/r2_ncu_f2i.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -2 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_default.json"))
print("main ms", round(d["ms_per_step"], 3), "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "launches", d["gpu_launches"], "parity", d.get("parity_checked"))
for k, v in d.get("per_config", {}).items():
    if "error" in v: print(k, "ERROR", v["error"]); continue
    print(k, "ms", round(v["ms_per_step"], 3), "value %.3e" % v["value"], "frac", round(v["roofline"]["frac"], 3), v["route"].get("route"))
PY

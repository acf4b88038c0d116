#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_default.json"))
print("main", d["config"]["workload"][:40], "ms", round(d["ms_per_step"], 3), "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "launches", d["gpu_launches"])
for k, v in d.get("per_config", {}).items():
    if "error" in v: print(k, "ERROR", v["error"]); continue
    print(k, "ms", round(v["ms_per_step"], 3), "value %.3e" % v["value"], "frac", round(v["roofline"]["frac"], 3), v["route"], v["config"].get("rows_redrawn"), "status!=0:", v["config"]["samples_with_failure_status"])
for k in ("parity_checked", "parity", "cpu_baseline", "cpu_baseline_port", "reference_cuda", "fp64_issue"):
    print(k, d.get(k))
PY
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"; tail -2 gpurun_out/r2_bench_reference.err; cat gpurun_out/r2_bench_reference.json | cut -c1-1500
python -m pytest tests -m gpu -x -q > gpurun_out/r2_test_gpu_all.log 2>&1; echo "all gpu tests rc=$?"; tail -5 gpurun_out/r2_test_gpu_all.log
python -m pytest tests/test_gpu_vs_reference.py -m gpu -x -q -s 2>&1 | grep -E "C2 B|C3 B|same-count|passed|failed" 

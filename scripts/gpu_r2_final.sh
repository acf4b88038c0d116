#!/bin/bash
# what the driver runs at round end, on one GPU: the GPU test-suite, smoke(), both bench arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_test_gpu_all.log 2>&1; echo "all gpu tests rc=$?"; tail -3 gpurun_out/r2_test_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --impl reference --steps 10 --warmup 3 ) > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"; grep real gpurun_out/r2_bench_reference.err
( time python bench.py ) > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$?"; grep real gpurun_out/r2_bench_default.err
python - <<'PY'
import json
r = json.load(open("gpurun_out/r2_bench_reference.json"))
print("reference arm: value %.3e" % r["value"], r["config"]["reference_path"], "threads", r["config"]["threads"], {k: (("%.3e" % v["value"]) if "value" in v else v) for k, v in r["reference_legs"].items()})
d = json.load(open("gpurun_out/r2_bench_default.json"))
print("main ms", round(d["ms_per_step"], 3), "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"], "launches", d["gpu_launches"], "parity", d.get("parity_checked"), "clocks", d.get("clocks"))
for k, v in d.get("per_config", {}).items():
    if "error" in v: print(k, "ERROR", v["error"]); continue
    print(k, "ms", round(v["ms_per_step"], 3), "value %.3e" % v["value"], "frac", round(v["roofline"]["frac"], 3), v["route"].get("route"))
print("cpu_baseline %.3e" % d["cpu_baseline"]["value"], "port %.3e" % d["cpu_baseline_port"]["value"], "reference_cuda %.3e" % d["reference_cuda"]["value"], d["reference_cuda"]["parity_vs_this_repo"])
print([ (k["kernel"][-30:], round(k["frac"],3)) for k in d["roofline_kernels"]])
PY

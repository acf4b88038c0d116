#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mlp_field.py -m gpu -x -q 2>&1 | tail -3
python -m pytest tests/test_gpu_vs_reference.py -m gpu -x -q -s -k c4 2>&1 | grep -E "C4 B|passed|failed"
python bench.py --workload c4 --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c4.json"))
print("c4 ms", round(d["ms_per_step"], 3), "value %.3e" % d["value"], d["route"])
PY
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest -m gpu -q -x tests/test_gpu_mlp_field.py -k "kernel_matches_fp32_reference and (64 or 1000 or 37)" > gpurun_out/r2_sanitizer_mlp2_$tool.log 2>&1
  echo "mlp2 $tool: exit $?  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_mlp2_$tool.log | tail -1)  $(grep -E ' passed| failed' gpurun_out/r2_sanitizer_mlp2_$tool.log | tail -1)"
done

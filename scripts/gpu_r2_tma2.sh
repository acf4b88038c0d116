#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp_field.py -m gpu -x -q 2>&1 | tail -2
for tool in racecheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest -m gpu -q -x tests/test_gpu_mlp_field.py -k "(kernel_matches_fp32_reference and (64 or 1000 or 37)) or all_stages_in_one_launch and 700" > gpurun_out/r2_sanitizer_mlp3_$tool.log 2>&1
  echo "mlp3 $tool: exit $?  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_mlp3_$tool.log | tail -1)  $(grep -E ' passed| failed' gpurun_out/r2_sanitizer_mlp3_$tool.log | tail -1)"
done
grep -m3 -A2 "Warning\|Error" gpurun_out/r2_sanitizer_mlp3_racecheck.log | cut -c1-250

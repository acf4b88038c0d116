#!/bin/bash
# compute-sanitizer memcheck + racecheck over the tcgen05 MLP field after the late round-2 changes (partial stage
# sums in shared memory, y tile in TMEM, newest operand out of the staged output tile, per-half weight barriers)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest -m gpu -q -x tests/test_gpu_mlp_field.py \
    -k "(kernel_matches_fp32_reference and not 19000 and not 8192) or (stage_fused_evaluation and 37) or (all_stages and (700 or 1000 or 37))" \
    > gpurun_out/r2_sanitizer_mlp2_$tool.log 2>&1
  rc=$?
  echo "mlp $tool: exit $rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_mlp2_$tool.log | tail -1)  $(grep -E ' passed| failed' gpurun_out/r2_sanitizer_mlp2_$tool.log | tail -1)"
done

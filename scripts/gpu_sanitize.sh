#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the kernels added in round 1 session 3: the step-fused heat
# route (shared-memory neighbour exchange, cp.async staging) and the split-mode initial step.
#   gpurun --timeout 600 -- 'bash scripts/gpu_sanitize.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL='(step_fused_heat_route_is_bit_identical and (3-8-dtype0 or 4-4100-dtype2 or 3-2054-dtype4 or 2-2050-dtype7) and False) or (init_kernels and (9000-f32-f32 or 8193-f64 or dt0-16388-f32-f32))'
for tool in memcheck racecheck; do
  timeout 280 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$SEL" \
    > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done

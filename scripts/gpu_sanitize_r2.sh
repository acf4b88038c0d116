#!/bin/bash
# round 2: compute-sanitizer memcheck + racecheck over the kernels round 1 left unchecked: the fused
# whole-solve kernels (general and the packed-fp32 persistent one), stage / finish incl. the split mode, the
# step-fused heat route at multi-chunk rows, and the tcgen05 MLP field (mbarrier / TMEM).
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize_r2.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name tool timeout pytest-args...
  local name=$1 tool=$2 to=$3; shift 3
  timeout $to compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest -m gpu -q -x "$@" > gpurun_out/r2_sanitizer_${name}_$tool.log 2>&1
  local rc=$?
  echo "$name $tool: exit $rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_${name}_$tool.log | tail -1)  $(grep -E ' passed| failed' gpurun_out/r2_sanitizer_${name}_$tool.log | tail -1)"
}
for tool in memcheck racecheck; do
  run fused $tool 300 tests/test_gpu_parity.py -k "fused_matches_oracle and (c1_readme_dopri5 or c2_vdp or lv_dtmax or linear_f4 or lv_data32)"
  run f2 $tool 300 tests/test_gpu_fused_f2.py -k "(variants and (bidir or rows or t1)) or (sizes and (1-- or 5- or 33)) or (equals_oracle and lv) or nonfinite"
  run stage_finish $tool 300 tests/test_gpu_kernels.py -k "(stage_kernel and f32-f32 and Tsit5) or (finish_kernel_one_iteration and (True-12-f32-f32 or True-520-f64 or False-9000-f32-f32 or True-2-f32-f64))"
  run split_heat $tool 300 tests/test_gpu_parity.py -k "heat_equation_routes and (large_heat_f32_tsit5_F16384 or heat_f32_tsit5_teval)"
  run mlp $tool 300 tests/test_gpu_mlp_field.py -k "kernel_matches_fp32_reference or stage_fused_evaluation"
done

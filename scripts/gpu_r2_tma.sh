#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp_field.py -m gpu -x -q 2>&1 | tail -3
for lib in "" build_variants/mlp_cpasync.so; do
  echo "=== ${lib:-default (TMA weights)}"
  TORCHODE_B200_LIB=${lib:+$PWD/$lib} python - <<'PY'
import os, sys, torch
if not os.environ.get("TORCHODE_B200_LIB"): os.environ.pop("TORCHODE_B200_LIB", None)
sys.path.insert(0, ".")
import bench
f = bench._mlp_field("cuda")
for B in (8192, 18944):
    y = torch.randn(B, 256, device="cuda")
    for _ in range(3): f(None, y)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): f(None, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"B={B}: {e0.elapsed_time(e1) / 20 * 1e3:.2f} us per eval (graph replay of 20)")
w = bench.C4("c4", 8192)
prob = bench.make_problem(w.host_inputs(0, 8192), "cuda")
field, method, ctrl = w.components("cuda")
solver = bench.to.AutoDiffAdjoint(method, ctrl)
with torch.no_grad():
    for _ in range(3): sol = solver.solve(prob)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): sol = solver.solve(prob)
    e1.record(); torch.cuda.synchronize()
print("C4 ms per solve", e0.elapsed_time(e1) / 10, solver.last_run["route"])
PY
done
for tool in memcheck racecheck; do
  timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest -m gpu -q -x tests/test_gpu_mlp_field.py -k "(kernel_matches_fp32_reference and (64 or 1000 or 37)) or all_stages_in_one_launch and 700" > gpurun_out/r2_sanitizer_mlp3_$tool.log 2>&1
  echo "mlp3 $tool: exit $?  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r2_sanitizer_mlp3_$tool.log | tail -1)  $(grep -E ' passed| failed' gpurun_out/r2_sanitizer_mlp3_$tool.log | tail -1)"
done

#!/bin/bash
# MLP field after the round-2 late changes (biases out of the critical path, partial stage sums under the MMAs,
# 16-byte copy-out): tests, then ms per solve of configs[3] with all stages in one launch / one launch per stage
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_mlp_field.py -x -q 2>&1 | tail -3
python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench, torchode_b200 as to
w = bench.C4("c4", 8192)
prob = bench.make_problem(w.host_inputs(0, 8192), "cuda")
field, method, ctrl = w.components("cuda")
for mode in (True, "stages"):
    solver = to.AutoDiffAdjoint(method, ctrl); solver.use_step_fusion = mode
    with torch.no_grad():
        for _ in range(3): sol = solver.solve(prob)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): sol = solver.solve(prob)
        e1.record(); torch.cuda.synchronize()
    print(mode, "ms per solve", e0.elapsed_time(e1) / 10, solver.last_run["route"], solver.last_run["iterations"])
PY

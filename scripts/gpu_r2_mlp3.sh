#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_c4.csv python scripts/profile_kernels.py c4 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2_launches_c4.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ik][:60]].append(float(r[iv].replace(",", "")))
    except ValueError: pass
for k, v in agg.items():
    print(f"{k:62s} n={len(v):3d} mean {sum(v)/len(v)/1e3:8.2f} us  min {min(v)/1e3:8.2f}")
PY
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
        for _ in range(5): sol = solver.solve(prob)
        e1.record(); torch.cuda.synchronize()
    print(mode, "ms per solve", e0.elapsed_time(e1) / 5, solver.last_run)
PY

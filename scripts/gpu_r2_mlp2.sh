#!/bin/bash
cd "$(dirname "$0")/.."
TORCHODE_B200_LIB=$PWD/build_variants/mlp_timing.so python scripts/mlp_timing.py
python - <<'PY'
import sys, torch
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
PY

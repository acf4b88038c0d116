#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_reference_suite.py -m gpu -x -q 2>&1 | tail -3
python -m pytest tests/test_gpu_mlp_field.py -m gpu -x -q 2>&1 | tail -2
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
python bench.py --workload c4 --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 ms', round(d['ms_per_step'],3), d['route'])"
echo "=== f2 default (LDS)"; python scripts/f2_timing.py 2>&1 | tail -1
echo "=== f2 generic load"; TORCHODE_B200_LIB=$PWD/build_variants/f2_nolds.so python scripts/f2_timing.py 2>&1 | tail -1
echo "=== f2 default (LDS)"; python scripts/f2_timing.py 2>&1 | tail -1

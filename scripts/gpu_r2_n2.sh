#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_abort.py tests/test_gpu_autodiff.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/test_symmetric_multi.py > gpurun_out/r2_sym_2.log 2>&1; echo "sym rc=$?"
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/r2_sym_2.log | tail -12
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 ) > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; tail -4 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2.json") if l.startswith("{")][-1])
    print("main ms", round(d["ms_per_step"], 3), "value %.3e" % d["value"], "solve_only", d.get("solve_only_ms"), d["route"])
    for k, v in d.get("per_config", {}).items():
        if "error" in v: print(k, "ERROR", v["error"]); continue
        print(k, "ms", round(v["ms_per_step"], 3), "solve_only", round(v.get("solve_only_ms") or 0, 3), "value %.3e" % v["value"], v["route"].get("route"), v.get("exchange"))
except Exception as e:
    print("parse failed", e)
PY

#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 4 2; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 ) > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; echo "bench n$n rc=$?"; tail -3 gpurun_out/r2_bench_n$n.err | grep real
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n$n.json") if l.startswith("{")][-1])
    print("N=$n main ms", round(d["ms_per_step"], 3), "value %.3e" % d["value"], "solve_only", round(d.get("solve_only_ms") or 0, 3), d["route"])
    for k, v in d.get("per_config", {}).items():
        if "error" in v: print(k, "ERROR", v["error"]); continue
        ex = v.get("exchange") or {}
        print(k, "ms", round(v["ms_per_step"], 3), "solve_only", round(v.get("solve_only_ms") or 0, 3), "value %.3e" % v["value"], "sharded %.3e" % v.get("value_results_left_sharded", 0), v["route"].get("route"), "nvlink GB/s", round(ex.get("nvlink_gbs_per_rank_over_the_step", 0)))
except Exception as e:
    print("parse failed", e)
PY
done
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29549 bench.py --impl reference --gpus 4 --steps 5 --warmup 2 ) > gpurun_out/r2_bench_reference_n4.json 2> gpurun_out/r2_bench_reference_n4.err; echo "ref n4 rc=$?"; grep real gpurun_out/r2_bench_reference_n4.err; cut -c1-300 gpurun_out/r2_bench_reference_n4.json

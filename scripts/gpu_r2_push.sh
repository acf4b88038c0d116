#!/bin/bash
# dense-output exchange at N GPUs: copy engines vs NCCL all-gather vs the SM push kernel (tode_peer_push)
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/test_symmetric_multi.py > gpurun_out/r2_sym_${N}_sm.log 2>&1; echo "sym rc=$?"
grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" gpurun_out/r2_sym_${N}_sm.log | tail -8
for mode in sm nccl copy; do
  TORCHODE_B200_PUSH=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload c3 --batch 2097152 --steps 5 --warmup 3 --no-extras > gpurun_out/r2_push_${mode}_n$N.json 2> gpurun_out/r2_push_${mode}_n$N.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_push_${mode}_n$N.json") if l.startswith("{")][-1])
    ex = d.get("exchange", {})
    print("$mode N=$N: ms", round(d["ms_per_step"], 3), "solve_only", round(d["solve_only_ms"], 3), "GB/s per rank over the step", round(ex.get("nvlink_gbs_per_rank_over_the_step", 0)), d["route"])
except Exception as e:
    print("$mode failed", e)
PY
done

#!/bin/bash
# round 2, first GPU trip: the packed-fp32 persistent dense-output kernel (erk_fused_f2.cuh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused_f2.py -x -q > gpurun_out/r2_test_f2.log 2>&1; echo "f2 tests rc=$?"; tail -5 gpurun_out/r2_test_f2.log
for nb in 20 22; do
  python bench.py --workload c3 --batch $((1<<nb)) --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_c3_2p${nb}_f2.json 2> gpurun_out/r2_bench_c3_2p${nb}_f2.err
  TODE_NO_F2=1 python bench.py --workload c3 --batch $((1<<nb)) --steps 10 --warmup 3 --no-extras > gpurun_out/r2_bench_c3_2p${nb}_general.json 2> gpurun_out/r2_bench_c3_2p${nb}_general.err
  python - <<PY
import json
for k in ("f2", "general"):
    try:
        d = json.load(open("gpurun_out/r2_bench_c3_2p${nb}_%s.json" % k))
        print("c3 2^${nb}", k, "ms", round(d["ms_per_step"], 4), "value %.3e" % d["value"], "frac", round(d["roofline"]["frac"], 4), d["route"], "e2e ms", round(d["e2e"]["ms_per_step"], 3))
    except Exception as e:
        print("c3 2^${nb}", k, "failed", e)
PY
done
ncu --set full --clock-control none --import-source on -k regex:solve_fused_f2 -s 1 -c 1 -o gpurun_out/r2_prof_f2_c3 -f python scripts/profile_kernels.py c3small > gpurun_out/r2_ncu_f2.log 2>&1
tail -2 gpurun_out/r2_ncu_f2.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_test_gpu_all.log 2>&1; echo "all gpu tests rc=$?"; tail -5 gpurun_out/r2_test_gpu_all.log

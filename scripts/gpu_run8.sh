cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 600 gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
for w in c1 c3 c4 c5; do python bench.py --workload $w --steps 5 --no-extras > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$w.json"))
    print("$w", "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], d["route"], "roofline frac %.3f" % d["roofline"]["frac"])
except Exception as e:
    print("$w failed", e); print(open("gpurun_out/bench_$w.err").read()[-1500:])
PY
done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 400 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_bench_c4.csv python bench.py --workload c4 --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_bench_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:solve_fused -s 1 -c 1 -o gpurun_out/prof_fused_c2 -f python scripts/profile_kernels.py c2 > gpurun_out/ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:erk_finish -s 1 -c 1 -o gpurun_out/prof_finish -f python scripts/profile_kernels.py path_a > gpurun_out/ncu_finish.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:erk_stage -s 13 -c 1 -o gpurun_out/prof_stage6 -f python scripts/profile_kernels.py path_a > gpurun_out/ncu_stage.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_tanh256 -s 2 -c 1 -o gpurun_out/prof_mlp -f python scripts/profile_kernels.py c4 > gpurun_out/ncu_mlp.log 2>&1
ls -la gpurun_out | tail -20

set -x
cd /root/repo
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
python bench.py --workload c3 --no-extras --steps 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 1500 gpurun_out/bench_c3.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; tail -c 1200 gpurun_out/bench_ref.json
nproc; lscpu | grep -E "Model name|^CPU\(s\)" 
# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_c2.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:solve_fused -s 1 -c 1 -o gpurun_out/prof_fused_c2 -f python scripts/profile_kernels.py c2 > gpurun_out/ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:erk_finish -s 1 -c 1 -o gpurun_out/prof_finish -f python scripts/profile_kernels.py path_a > gpurun_out/ncu_finish.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:erk_stage -s 13 -c 1 -o gpurun_out/prof_stage6 -f python scripts/profile_kernels.py path_a > gpurun_out/ncu_stage.log 2>&1
ls -la gpurun_out

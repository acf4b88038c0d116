cd /root/repo
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --no-extras | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
python bench.py --workload c4 --steps 10 --no-extras > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c4.json')); print('c4 value %.4g ms %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['route'], d['config']['mean_n_steps'])" || tail -20 gpurun_out/bench_c4.err
python scripts/quick_time.py staged 2>&1 | grep -E "graph=True" | head -5

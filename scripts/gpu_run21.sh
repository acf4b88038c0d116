cd /root/repo
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k solve_from_host 2>&1 | tail -3
python bench.py --no-extras > gpurun_out/bench_c2_run21.json 2> gpurun_out/bench_c2_run21.err; tail -3 gpurun_out/bench_c2_run21.err; python -c "
import json; d=json.load(open('gpurun_out/bench_c2_run21.json')); print(d['ms_per_step'], d['value'], d['e2e'])"
for c in 1 2 4 8 16 32; do python - <<PY
import sys, time, torch
sys.path.insert(0,'.')
import bench, torchode_b200 as to
w = bench.C2('c2', 1<<20); host = {k:(None if v is None else v.pin_memory()) for k,v in w.host_inputs(0, 1<<20).items()}
_, m, c = w.components(); solver = to.AutoDiffAdjoint(m, c)
hp = to.InitialValueProblem(host['y0'], host['t_start'], host['t_end'])
out=None; ts=[]
for i in range(8):
    torch.cuda.synchronize(); t0=time.perf_counter(); out = to.solve_from_host(solver, hp, 'cuda', chunks=$c, out=out); ts.append(time.perf_counter()-t0)
print('chunks', $c, 'ms', ['%.2f' % (1e3*t) for t in ts[2:]])
PY
done

"""configs[4] (64 rows x 2^20 grid values, 268 MB in and out) end to end from host buffers: solve_from_host with the
batch cut into 1 / 2 / 4 / 8 chunks, driven by one or two host threads, against copy-in, solve, copy-out in sequence."""
import statistics
import sys
import time

import torch

sys.path.insert(0, ".")
import bench
import torchode_b200 as to

w = bench.C5("c5", 64)
host = {k: (None if v is None else v.pin_memory()) for k, v in w.host_inputs(0, 64).items()}
field, method, ctrl = w.components("cuda")
solver = to.AutoDiffAdjoint(method, ctrl)
te = host["t_eval"]
if te is not None and te.ndim == 1:
    te = te.expand(64, -1)
hp = to.InitialValueProblem(host["y0"], host["t_start"], host["t_end"], te)
moved = hp.y0.numel() * hp.y0.element_size() * 2
with torch.no_grad():
    for workers in (1, 2):
        for chunks in (1, 2, 4, 8):
            out, times = None, []
            for i in range(6):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = to.solve_from_host(solver, hp, "cuda", chunks=chunks, min_chunk_bytes=max(1, moved // chunks),
                                         workers=workers, out=out)
                if i >= 2:
                    times.append(time.perf_counter() - t0)
            print(f"host threads {workers}, chunks asked {chunks} run {solver.last_run['chunks']}: "
                  f"{statistics.median(times) * 1e3:.2f} ms (accepted {int(out.stats['n_accepted'].sum())})")

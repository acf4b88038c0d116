"""Device time of one C3 solve (Lotka-Volterra, Dopri5 + I(1e-6,1e-3), fp32, 100 shared t_eval points) per
library variant: TORCHODE_B200_LIB selects the .so (scripts/build_variant.sh).  L2 flushed between solves."""
import statistics
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    dev = torch.device("cuda", 0)
    w = bench.C3("c3", B)
    prob = bench.make_problem(w.host_inputs(0, B), dev)
    _, method, ctrl = w.components()
    solver = bench.to.AutoDiffAdjoint(method, ctrl)
    l2 = torch.zeros(128 << 20, dtype=torch.float32, device=dev)
    times = []
    with torch.no_grad():
        for i in range(13):
            l2.add_(1)
            e0, e1 = bench.ev_pair()
            e0.record()
            sol = solver.solve(prob)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                times.append(e0.elapsed_time(e1))
    acc = int(sol.stats["n_accepted"].sum())
    ms = statistics.median(times)
    print(f"B={B} median {ms:.4f} ms min {min(times):.4f} ms  {acc / ms * 1e3:.3e} acc-steps/s  "
          f"HBM frac {B * 848 / ms / 1e6 / 6447.8:.3f}  route {solver.last_run}")

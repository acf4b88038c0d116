"""Runs each hot kernel a few times so that ncu can capture it (see scripts/gpu_profile.sh)."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "path_a"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    if which == "path_a":
        for r in bench.measure_path_a_kernels(dev, bench.peaks()[0], reps=2):
            print(r)
    else:
        cls, batch = bench.WORKLOADS[which]
        w = cls(which, batch)
        prob = bench.make_problem(w.host_inputs(0, w.batch), dev)
        _, method, ctrl = w.components()
        solver = bench.to.AutoDiffAdjoint(method, ctrl)
        with torch.no_grad():
            for _ in range(2):
                sol = solver.solve(prob)
        torch.cuda.synchronize()
        print(which, int(sol.stats["n_accepted"].sum()))

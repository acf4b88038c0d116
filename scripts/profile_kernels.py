"""Runs each hot kernel a few times so that ncu can capture it (see scripts/gpu_profile.sh)."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "path_a"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    if which.startswith("path_a"):
        shapes = {"path_a": [(1 << 24, 2, torch.float32)],
                  "path_a_all": [(1 << 24, 2, torch.float32), (1 << 23, 2, torch.float64),
                                 (1 << 18, 256, torch.float32), (1 << 17, 256, torch.float64),
                                 (1 << 22, 8, torch.float32), (1 << 14, 4096, torch.float32),
                                 (64, 1 << 20, torch.float32)]}[which]
        for B, F, dt in shapes:
            for r in bench.measure_path_a_kernels(dev, bench.peaks()[0], reps=3, B=B, F=F, dtype=dt):
                print(f"{r['kernel']:55s} {r['ms']:8.4f} ms {r['achieved']:8.1f} GB/s frac {r['frac']:.3f}"
                      + (f" acc {r['accepted_fraction']:.2f}" if 'accepted_fraction' in r else ""))
    else:
        small = which.endswith("small")
        which = which.replace("small", "")
        cls, batch = bench.WORKLOADS[which]
        if small:
            batch = 1 << 20
        w = cls(which, batch)
        prob = bench.make_problem(w.host_inputs(0, w.batch), dev)
        _, method, ctrl = w.components(dev) if getattr(w, "staged", False) else w.components()
        solver = bench.to.AutoDiffAdjoint(method, ctrl)
        with torch.no_grad():
            for _ in range(2):
                sol = solver.solve(prob)
        torch.cuda.synchronize()
        print(which, int(sol.stats["n_accepted"].sum()))

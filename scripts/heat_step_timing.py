"""C5 (heat equation, 64 x 2^20, Tsit5 + I) on the step-fused route against the stage-wise route:
time per solve, iterations, equality of the results.  Usage: python scripts/heat_step_timing.py [B N]"""
import sys

import torch

sys.path.insert(0, ".")
import torchode_b200 as to
from torchode_b200.fields import Heat1D


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None else min(best, ms)
    return best, out


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    g = torch.Generator().manual_seed(1234)
    x = torch.linspace(0, 1, N)
    amp = torch.rand(B, 3, generator=g)
    y0 = sum(amp[:, k - 1:k] * torch.sin(k * torch.pi * x)[None] for k in (1, 2, 3)).cuda()
    prob = to.InitialValueProblem(y0, torch.zeros(B, device="cuda"), torch.ones(B, device="cuda"))
    term = to.ODETerm(Heat1D(25.0))
    sols = {}
    with torch.no_grad():
        for fusion in (True, False):
            solver = to.AutoDiffAdjoint(to.Tsit5(term), to.IntegralController(1e-6, 1e-3, term=term))
            solver.use_step_fusion = fusion
            ms, sol = timed(lambda: solver.solve(prob))
            run = solver.last_run
            acc = int(sol.stats["n_accepted"].sum())
            att = int(sol.stats["n_steps"].sum())
            rows = 4 * att if fusion else 56 * att
            print(f"{run['route']:>10}: {ms:8.2f} ms / solve, {run['iterations']} iterations ({ms / run['iterations']:.3f} ms each), "
                  f"accepted {acc}, {acc / ms * 1e3:.3e} acc-steps/s, algorithmic {rows * N * 4 / ms / 1e6:.0f} GB/s")
            sols[fusion] = sol
    same = torch.equal(sols[True].ys, sols[False].ys) and all(
        torch.equal(sols[True].stats[k], sols[False].stats[k]) for k in ("n_steps", "n_accepted"))
    print("bit-identical:", same)


if __name__ == "__main__":
    main()

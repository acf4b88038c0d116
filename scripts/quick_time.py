"""Quick device timings used while developing (not the bench contract; see bench.py)."""
import sys
import time

import torch

sys.path.insert(0, ".")
import torchode_b200 as to
from torchode_b200.fields import LotkaVolterra, VanDerPol


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


GRAPH = False


def c2(B=1 << 20, staged=False):
    g = torch.Generator().manual_seed(1234)
    y0 = (torch.rand(B, 2, generator=g, dtype=torch.float64) * 4 - 2).cuda()
    field = VanDerPol(10.0)
    f = (lambda t, y: field(t, y)) if staged else field
    term = to.ODETerm(f)
    solver = to.AutoDiffAdjoint(to.Tsit5(term), to.PIDController(1e-8, 1e-8, 0.2, 0.5, 0.0, term=term))
    solver.use_cuda_graph = GRAPH
    prob = to.InitialValueProblem(y0, torch.zeros(B, dtype=torch.float64, device="cuda"),
                                  torch.full((B,), 20.0, dtype=torch.float64, device="cuda"))
    ms, sol = timed(lambda: solver.solve(prob), n=2 if staged else 3)
    acc = int(sol.stats["n_accepted"].sum())
    print(f"C2 vdp B={B} staged={staged} graph={GRAPH and staged}: {ms:.2f} ms, accepted {acc}, {acc / ms * 1e3:.3e} acc-steps/s, "
          f"iters {int(sol.stats['n_f_evals'][0] - 2) // 6}, mean n_steps {sol.stats['n_steps'].float().mean():.1f}")


def c3(B=1 << 22, staged=False, T=100):
    g = torch.Generator().manual_seed(1234)
    y0 = (1 + torch.rand(B, 2, generator=g)).cuda()
    t_eval = torch.linspace(0, 10, T).cuda().expand(B, T)
    field = LotkaVolterra()
    f = (lambda t, y: field(t, y)) if staged else field
    term = to.ODETerm(f)
    solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
    solver.use_cuda_graph = GRAPH
    prob = to.InitialValueProblem(y0, t_eval=t_eval)
    ms, sol = timed(lambda: solver.solve(prob), n=2 if staged else 3)
    acc = int(sol.stats["n_accepted"].sum())
    byts = B * (8 + 8 + T * 8 + 32)
    print(f"C3 lv B={B} T={T} staged={staged} graph={GRAPH and staged}: {ms:.2f} ms, accepted {acc}, {acc / ms * 1e3:.3e} acc-steps/s, "
          f"algorithmic {byts / ms / 1e6:.1f} GB/s, mean n_steps {sol.stats['n_steps'].float().mean():.1f}")


def heat(B=64, N=1 << 20):
    """C5: 1-D heat equation, method of lines, Tsit5 + I(1e-6, 1e-3), fp32, staged route."""
    g = torch.Generator().manual_seed(1234)
    x = torch.linspace(0, 1, N)
    amp = torch.rand(B, 3, generator=g)
    y0 = sum(amp[:, k - 1:k] * torch.sin(k * torch.pi * x)[None] for k in (1, 2, 3)).cuda()
    kappa = 25.0

    def f(t, y):
        out = torch.zeros_like(y)
        out[:, 1:-1] = kappa * ((y[:, 2:] - 2 * y[:, 1:-1]) + y[:, :-2])
        return out

    term = to.ODETerm(f)
    solver = to.AutoDiffAdjoint(to.Tsit5(term), to.IntegralController(1e-6, 1e-3, term=term))
    solver.use_cuda_graph = GRAPH
    prob = to.InitialValueProblem(y0, torch.zeros(B, device="cuda"), torch.ones(B, device="cuda"))
    ms, sol = timed(lambda: solver.solve(prob), n=2)
    acc = int(sol.stats["n_accepted"].sum())
    iters = (int(sol.stats["n_f_evals"][0]) - 2) // 6
    solver_bytes = iters * B * N * 4 * 44  # solver-owned traffic per attempted step: 44 F e
    print(f"C5 heat B={B} N={N} graph={GRAPH}: {ms:.2f} ms, iters {iters}, accepted {acc}, "
          f"{acc / ms * 1e3:.3e} acc-steps/s, solver-owned {solver_bytes / ms / 1e6:.0f} GB/s (f excluded)")


if __name__ == "__main__":
    import sys
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    with torch.no_grad():
        if which == "all":
            c2(1 << 16)
            c2(1 << 20)
            c3(1 << 20)
        if which == "fusedonly":
            c2(1 << 20)
            c3(1 << 20)
            sys.exit(0)
        for GRAPH in (False, True):
            c2(1 << 10, staged=True)
            c2(1 << 14, staged=True)
            c3(1 << 16, staged=True)
            c3(1 << 20, staged=True)
            heat(64, 1 << 20)

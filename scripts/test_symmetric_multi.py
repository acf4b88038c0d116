"""N-GPU check of the peer-store gather (run under torchrun on a GPU box): the fused kernel writing
the gathered Solution into every rank's symmetric buffers must equal the single-GPU solve of the whole
batch bit for bit -- also when a sample of ONE shard fails ("any failure stops the whole batch" holds
across GPUs on this path) -- and, without failures, the NCCL-gathered Solution; the dense-output block is
pushed whole or in four chunks under the solve; plus timings."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchode_b200 as to  # noqa: E402
from torchode_b200.distributed import SymmetricWorkspace, shard_problem, solve_sharded  # noqa: E402
from torchode_b200.fields import LotkaVolterra, VanDerPol  # noqa: E402


def same(a, b, n_init=None):
    """bit equality; with ``n_init`` only the first n_init[b] evaluation points of sample b count
    (the rest of a stopped sample's row is uninitialised memory by the reference's contract)"""
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    eq = a.contiguous().view(torch.uint8) == b.contiguous().view(torch.uint8)
    if n_init is not None:
        valid = torch.arange(a.shape[1], device=a.device)[None, :] < n_init[:, None]
        eq = eq.view(a.shape[0], a.shape[1], -1) | ~valid[:, :, None]
    return bool(eq.all())


def check(name, solver, problem, rank, world):
    B, F, Tn = problem.batch_size, problem.n_features, problem.n_evaluation_points
    ws = SymmetricWorkspace(B // world, Tn, F, problem.data_dtype, problem.device)
    full = solver.solve(problem)  # the whole batch on this GPU: what the reference semantics define
    n_fail = int((full.status != 0).sum())
    for rep, chunks in enumerate((1, 4, 1)):  # the workspace is reused; whole-block and chunked pushes
        got = solve_sharded(solver, problem, workspace=ws, chunks=chunks)
        ok = same(got.ys, full.ys, full.stats["n_initialized"]) and same(got.status, full.status)
        for k in ("n_steps", "n_accepted", "n_initialized", "n_f_evals"):
            ok = ok and got.stats[k].tolist() == full.stats[k].tolist()
        assert ok, f"{name}: peer-store gather differs from the single-GPU solve (rank {rank}, rep {rep}, chunks {chunks})"
    if n_fail == 0:  # the NCCL path stops a failing shard only
        ref = solve_sharded(solver, problem)
        ok = same(got.ys, ref.ys, ref.stats["n_initialized"]) and same(got.status, ref.status)
        for k in ("n_steps", "n_accepted", "n_initialized", "n_f_evals"):
            ok = ok and got.stats[k].tolist() == ref.stats[k].tolist()
        assert ok, f"{name}: peer-store gather differs from the NCCL gather (rank {rank})"
    torch.cuda.synchronize()
    dist.barrier()

    def timed(fn, n=5):
        out = []
        for _ in range(n):
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            out.append(time.perf_counter() - t0)
        return sorted(out)[len(out) // 2] * 1e3

    local = shard_problem(problem, rank, world)
    t_local = timed(lambda: solver.solve(local))
    t_nccl = timed(lambda: solve_sharded(solver, problem))
    t_sym = timed(lambda: solve_sharded(solver, problem, workspace=ws))
    t_sym4 = timed(lambda: solve_sharded(solver, problem, workspace=ws, chunks=4))
    if rank == 0:
        print(f"{name:28s} B={B} failures={n_fail}: local solve {t_local:.3f} ms, +NCCL gather {t_nccl:.3f} ms, "
              f"peer stores {t_sym:.3f} ms, in 4 chunks {t_sym4:.3f} ms", flush=True)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    g = torch.Generator().manual_seed(99)
    with torch.no_grad():
        B = 4096 * world
        y0 = (torch.rand(B, 2, generator=g, dtype=torch.float64) * 4 - 2).to(dev)
        term = to.ODETerm(VanDerPol(10.0))
        solver = to.AutoDiffAdjoint(to.Tsit5(term), to.PIDController(1e-8, 1e-8, 0.2, 0.5, 0.0, term=term))
        check("vdp f64 no t_eval", solver, to.InitialValueProblem(
            y0, torch.zeros(B, dtype=torch.float64, device=dev), torch.full((B,), 5.0, dtype=torch.float64, device=dev)),
            rank, world)
        y0 = (1 + torch.rand(B, 2, generator=g)).to(dev)
        te = torch.linspace(0, 10, 100).to(dev).expand(B, -1)
        term = to.ODETerm(LotkaVolterra())
        solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
        check("lv f32 100 t_eval", solver, to.InitialValueProblem(y0, t_eval=te), rank, world)
        # a shard with a failing sample: max_steps cuts everyone off
        solver2 = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term), max_steps=7)
        check("lv f32 max_steps=7", solver2, to.InitialValueProblem(y0, t_eval=te), rank, world)
        # failure in ONE shard only: a non-finite initial condition in the last row
        y0b = y0.clone()
        y0b[-1, 0] = float("inf")
        check("lv f32 inf in last shard", solver, to.InitialValueProblem(y0b, t_eval=te), rank, world)
        # configs[2] at 2^20 samples per rank: the dense-output block (838 MB per rank) rides under the solve
        B = (1 << 20) * world
        g = torch.Generator().manual_seed(1234)
        y0 = (1 + torch.rand(B, 2, generator=g)).to(dev)
        te = torch.linspace(0, 10, 100).to(dev).expand(B, -1)
        check("C3 2^20 per rank", solver, to.InitialValueProblem(y0, t_eval=te), rank, world)
        B = (1 << 20) * world
        g = torch.Generator().manual_seed(1234)
        y0 = (torch.rand(B, 2, generator=g, dtype=torch.float64) * 4 - 2).to(dev)
        term = to.ODETerm(VanDerPol(10.0))
        solver = to.AutoDiffAdjoint(to.Tsit5(term), to.PIDController(1e-8, 1e-8, 0.2, 0.5, 0.0, term=term))
        check("C2 2^20 per rank", solver, to.InitialValueProblem(
            y0, torch.zeros(B, dtype=torch.float64, device=dev), torch.full((B,), 20.0, dtype=torch.float64, device=dev)),
            rank, world)
    if rank == 0:
        print("symmetric multi-GPU checks passed", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

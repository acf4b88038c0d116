"""Peer-store gather vs NCCL gather for a dense-output workload (C3 shape: 100 t_eval points)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchode_b200 as to
from torchode_b200.distributed import SymmetricWorkspace, solve_sharded
from torchode_b200.fields import LotkaVolterra

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
with torch.no_grad():
    for logB in (16, 20):
        B = (1 << logB) * world
        g = torch.Generator().manual_seed(7)
        y0 = (1 + torch.rand(B, 2, generator=g)).to(dev)
        te = torch.linspace(0, 10, 100).to(dev).expand(B, -1)
        term = to.ODETerm(LotkaVolterra())
        solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
        prob = to.InitialValueProblem(y0, t_eval=te)
        ws = SymmetricWorkspace(B // world, 100, 2, torch.float32, dev)
        def timed(fn, n=4):
            out = []
            for _ in range(n):
                torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
                t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); out.append(time.perf_counter() - t0)
            return sorted(out)[len(out) // 2] * 1e3
        a = timed(lambda: solve_sharded(solver, prob, gather=False))
        b = timed(lambda: solve_sharded(solver, prob))
        c = timed(lambda: solve_sharded(solver, prob, workspace=ws))
        # the same step, phase by phase (host clock, device synchronised after every phase)
        from torchode_b200.distributed import shard_problem
        local = shard_problem(prob, rank, world)
        for rep in range(3):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            marks = [time.perf_counter()]
            def mark():
                torch.cuda.synchronize(); marks.append(time.perf_counter())
            ws.glob.zero_(); ws.barrier(); mark()
            ctx = solver._fused_launch(local, term, term.f, None, peers=ws); mark()
            ws.push_ys(); mark()
            ws.barrier(); mark()
            ws.glob.tolist(); ctx["summary"].tolist(); mark()
        if rank == 0:
            names = ["zero+barrier", "kernel", "push", "barrier", "reads"]
            print("   phases (ms): " + "  ".join(f"{n} {1e3 * (marks[i + 1] - marks[i]):.3f}" for i, n in enumerate(names)), flush=True)
        if rank == 0:
            print(f"2^{logB} per rank, T=100: local {a:.3f} ms, NCCL gather {b:.3f} ms, peer stores {c:.3f} ms", flush=True)
        del ws
dist.barrier(); dist.destroy_process_group()

"""Breakdown of one symmetric (peer-store) step for a dense-output workload (2 GPUs)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchode_b200 as to
from torchode_b200.distributed import SymmetricWorkspace, shard_problem
from torchode_b200.fields import LotkaVolterra

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B = 1 << 20
g = torch.Generator().manual_seed(7 + rank)
y0 = (1 + torch.rand(B, 2, generator=g)).to(dev)
te = torch.linspace(0, 10, 100).to(dev).expand(B, -1)
term = to.ODETerm(LotkaVolterra())
solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
prob = to.InitialValueProblem(y0, t_eval=te)
ws = SymmetricWorkspace(B, 100, 2, torch.float32, dev)
field = term.f

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

with torch.no_grad():
    for rep in range(4):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        e = [ev()]
        ws.glob.zero_(); ws.barrier(); e.append(ev())
        ctx = solver._fused_launch(prob, term, field, None, peers=ws); e.append(ev())
        ws.push_ys(); e.append(ev())
        ws.barrier(); e.append(ev())
        gl = ws.glob.tolist(); e.append(ev())
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ctx2 = solver._fused_launch(prob, term, field, None); e2 = ev(); torch.cuda.synchronize()
        if rank == 0 and rep >= 2:
            names = ["zero+barrier", "kernel (+peer stats)", "push ys", "barrier", "tolist"]
            print("  ".join(f"{n} {e[i].elapsed_time(e[i + 1]):.3f}" for i, n in enumerate(names)), f" wall {wall:.3f} ms; plain kernel {e[5].elapsed_time(e2):.3f}", flush=True)
dist.barrier(); dist.destroy_process_group()

"""Diagnostics of the multi-GPU step (run under torchrun): where does the time of a sharded C2
step go (solve vs the gathers), which transport NCCL picked, and whether symmetric (peer-mapped)
memory is available on the box.  Measurement aid, not part of the product path."""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torchode_b200 as to  # noqa: E402
from torchode_b200.distributed import gather_solution  # noqa: E402
from torchode_b200.fields import VanDerPol  # noqa: E402


def ev():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timeit(fn, n=5):
    out = []
    for _ in range(n):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = ev()
        a.record()
        fn()
        b.record()
        b.synchronize()
        out.append(a.elapsed_time(b))
    return sorted(out)[len(out) // 2], out


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    B = 1 << 20
    g = torch.Generator().manual_seed(1234 + rank)
    y0 = (torch.rand(B, 2, generator=g, dtype=torch.float64) * 4 - 2).to(dev)
    prob = to.InitialValueProblem(y0, torch.zeros(B, dtype=torch.float64, device=dev),
                                  torch.full((B,), 20.0, dtype=torch.float64, device=dev))
    term = to.ODETerm(VanDerPol(10.0))
    solver = to.AutoDiffAdjoint(to.Tsit5(term), to.PIDController(1e-8, 1e-8, 0.2, 0.5, 0.0, term=term))
    res = {}
    with torch.no_grad():
        sol = solver.solve(prob)
        gather_solution(sol, B * world)
        res["solve"] = timeit(lambda: solver.solve(prob))
        sol = solver.solve(prob)
        res["gather_solution"] = timeit(lambda: gather_solution(sol, B * world))
        res["solve+gather"] = timeit(lambda: gather_solution(solver.solve(prob), B * world))
        for mb in (1, 16, 64, 256):
            x = torch.empty(mb << 20, dtype=torch.uint8, device=dev)
            o = torch.empty(world * (mb << 20), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(o, x)
            res[f"all_gather {mb} MiB/rank"] = timeit(lambda: dist.all_gather_into_tensor(o, x))
        one = torch.ones(1, device=dev)
        res["all_reduce 4 B"] = timeit(lambda: dist.all_reduce(one, op=dist.ReduceOp.MAX))

        def host_sync_reduce():
            dist.all_reduce(one, op=dist.ReduceOp.MAX)
            one.item()
        res["all_reduce 4 B + item()"] = timeit(host_sync_reduce)
    if rank == 0:
        for k, (med, all_) in res.items():
            print(f"{k:32s} median {med:8.3f} ms   {['%.3f' % v for v in all_]}", flush=True)

    # symmetric memory probe
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(1 << 20, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD)
        t.fill_(float(rank))
        hdl.barrier()
        peer = (rank + 1) % world
        pt = hdl.get_buffer(peer, (1 << 20,), torch.float32)
        seen = float(pt[0])
        hdl.barrier()
        pt[1:2].fill_(100.0 + rank)  # peer store
        hdl.barrier()
        torch.cuda.synchronize()
        print(f"[rank {rank}] symm_mem ok: peer {peer} value {seen}, my[1]={float(t[1])}, "
              f"buffer_ptrs={[hex(p) for p in hdl.buffer_ptrs]}, multicast_ptr={hex(hdl.multicast_ptr)}",
              flush=True)
        a, b = ev()
        big = symm.empty(64 << 20, dtype=torch.uint8, device=dev)
        h2 = symm.rendezvous(big, dist.group.WORLD)
        pb = h2.get_buffer(peer, (64 << 20,), torch.uint8)
        src = torch.ones(64 << 20, dtype=torch.uint8, device=dev)
        pb.copy_(src)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            pb.copy_(src)
        b.record()
        b.synchronize()
        print(f"[rank {rank}] peer copy 64 MiB: {64 / 1024 * 5 / (a.elapsed_time(b) * 1e-3):.1f} GiB/s", flush=True)
    except Exception as exc:  # noqa: BLE001
        print(f"[rank {rank}] symm_mem unavailable: {type(exc).__name__}: {exc}", flush=True)
    time.sleep(0.5)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""N > 1 path on CPU: two gloo ranks shard a batch, solve their slices independently (generic
plug-in route with scripted components -- the kernels need a GPU) and all-gather the Solution."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import torchode_b200 as to
from torchode_b200.distributed import gather_solution, shard_bounds, shard_problem, solve_sharded

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_host_logic import ExactStep, Scripted, exact  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _full_problem(B):
    t_start = torch.linspace(0.0, 1.0, B)
    t_end = t_start + torch.linspace(0.5, 2.0, B)  # ragged spans: different trip counts per rank
    t_eval = t_start[:, None] + (t_end - t_start)[:, None] * torch.linspace(0, 1, 5)[None]
    return to.InitialValueProblem(exact(t_start), t_start, t_end, t_eval)


def _worker(rank, world, port, B, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        problem = _full_problem(B)
        solver = to.AutoDiffAdjoint(ExactStep(), Scripted(0.11, 0.11))
        sol = solve_sharded(solver, problem)
        if rank == 0:
            single = to.AutoDiffAdjoint(ExactStep(), Scripted(0.11, 0.11)).solve(problem)
            result["ys_equal"] = bool(torch.equal(sol.ys, single.ys))
            result["n_steps"] = sol.stats["n_steps"].tolist() == single.stats["n_steps"].tolist()
            result["n_accepted"] = sol.stats["n_accepted"].tolist() == single.stats["n_accepted"].tolist()
            result["n_init"] = sol.stats["n_initialized"].tolist() == single.stats["n_initialized"].tolist()
            result["status"] = sol.status.tolist() == single.status.tolist()
            result["shape"] = tuple(sol.ys.shape)
            result["ts_is_t_eval"] = sol.ts is problem.t_eval
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])  # even split and ragged split (7 = 4 + 3)
def test_two_rank_shard_solve_gather_equals_single_process(B):
    port = _free_port()
    with mp.Manager() as mgr:
        result = mgr.dict()
        mp.spawn(_worker, args=(2, port, B, result), nprocs=2, join=True)
        assert result["ys_equal"] and result["n_steps"] and result["n_accepted"]
        assert result["n_init"] and result["status"] and result["ts_is_t_eval"]
        assert result["shape"] == (B, 5, 1)


def test_shard_bounds_cover_the_batch_exactly():
    for B in (1, 7, 8, 1 << 20):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_shard_problem_slices_every_field():
    p = _full_problem(10)
    s = shard_problem(p, 1, 3)  # rows [4, 7)
    assert s.batch_size == 3 and torch.equal(s.y0, p.y0[4:7]) and torch.equal(s.t_eval, p.t_eval[4:7])

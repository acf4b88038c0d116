"""Multi-GPU: shard independent samples across ranks, gather solutions and statistics.

Samples of a batch never interact inside the solve loop (every reduction is over the feature
dimension), so the batch dimension shards with no collective in the step loop: one process
per GPU solves its contiguous slice with its own adaptive step sizes and its own loop trip
count.  Collectives run once, after the solve (NCCL over NVLink on the B200 box, gloo in
the CPU tests):

* ``all_gather`` of ``ys``, ``n_steps``, ``n_accepted``, ``n_initialized``, ``status``;
* ``all_reduce(MAX)`` of the per-rank loop iteration count, because the reference's
  ``n_f_evals`` is batch-uniform (``n_init + 6 * max_b n_steps``, terms.py:54-58).

Known difference to a single-device solve: "any failure stops the whole batch"
(adjoints.py:186-190) holds per shard, not across shards (a failing sample only cuts off the
samples of its own rank).
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from .problems import InitialValueProblem
from .solution import Solution


def shard_bounds(batch_size: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, near-equal slices: the first ``batch_size % world_size`` ranks get one more row."""
    base, extra = divmod(batch_size, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_problem(problem: InitialValueProblem, rank: int, world_size: int) -> InitialValueProblem:
    lo, hi = shard_bounds(problem.batch_size, rank, world_size)
    t_eval = None if problem.t_eval is None else problem.t_eval[lo:hi]
    return InitialValueProblem(problem.y0[lo:hi], problem.t_start[lo:hi], problem.t_end[lo:hi], t_eval)


def _gather_rows_async(x: torch.Tensor, sizes, group):
    """all_gather of row blocks that may differ by one row between ranks.  Returns
    ``(work, finish)``: wait for ``work``, then ``finish()`` yields the gathered tensor."""
    world = len(sizes)
    x = x.contiguous()
    if len(set(sizes)) == 1:
        out = x.new_empty((world * sizes[0],) + tuple(x.shape[1:]))
        work = dist.all_gather_into_tensor(out, x, group=group, async_op=True)
        return work, lambda: out
    pad = max(sizes)
    padded = x.new_zeros((pad,) + tuple(x.shape[1:]))
    padded[: x.shape[0]] = x
    out = x.new_empty((world * pad,) + tuple(x.shape[1:]))
    work = dist.all_gather_into_tensor(out, padded, group=group, async_op=True)
    return work, lambda: torch.cat([out[r * pad: r * pad + n] for r, n in enumerate(sizes)])


def _gather_rows(x: torch.Tensor, sizes, group) -> torch.Tensor:
    work, finish = _gather_rows_async(x, sizes, group)
    work.wait()
    return finish()


def gather_solution(local: Solution, global_batch: int, ts: Optional[torch.Tensor] = None,
                    group=None) -> Solution:
    """Assemble the full-batch Solution on every rank from the per-rank ones.  Every tensor is
    gathered straight into its final buffer (no packing / unpacking passes over the gathered
    data); the collectives are enqueued back to back and waited for once; the iteration count
    goes through one all_reduce(MAX)."""
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in (shard_bounds(global_batch, r, world) for r in range(world))]
    keys = [k for k in ("n_steps", "n_accepted", "n_initialized") if k in local.stats]
    pending = [_gather_rows_async(local.ys, sizes, group),
               _gather_rows_async(local.status.to(torch.long), sizes, group)]
    pending += [_gather_rows_async(local.stats[k].to(torch.long), sizes, group) for k in keys]
    if ts is None:
        pending.append(_gather_rows_async(local.ts, sizes, group))
    n = None
    if "n_f_evals" in local.stats:
        # batch-uniform in the reference: every sample is charged the evaluations of the
        # longest-running one -> MAX over ranks
        n = local.stats["n_f_evals"][:1].to(local.ys.device, copy=True)
        dist.all_reduce(n, op=dist.ReduceOp.MAX, group=group)
    for work, _ in pending:
        work.wait()
    gathered = [finish() for _, finish in pending]
    ys, status = gathered[0], gathered[1]
    stats = {k: gathered[2 + i] for i, k in enumerate(keys)}
    if n is not None:
        stats["n_f_evals"] = torch.full((1,), int(n.item()), dtype=torch.long).expand(global_batch)
    if ts is None:
        ts = gathered[-1]
    return Solution(ts=ts, ys=ys, stats=stats, status=status)


def solve_sharded(solver, problem: InitialValueProblem, *, dt0: Optional[torch.Tensor] = None,
                  args=None, group=None, gather: bool = True) -> Solution:
    """Every rank holds (or can build) the full problem; each solves its slice.

    With ``gather=False`` the local Solution is returned (statistics of the slice only)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(problem.batch_size, rank, world)
    local = solver.solve(shard_problem(problem, rank, world),
                         dt0=None if dt0 is None else dt0[lo:hi], args=args)
    if not gather:
        return local
    return gather_solution(local, problem.batch_size, ts=problem.t_eval, group=group)

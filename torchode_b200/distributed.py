"""Multi-GPU: shard independent samples across ranks, gather solutions and statistics.

Samples of a batch never interact inside the solve loop (every reduction is over the feature
dimension), so the batch dimension shards with no collective in the step loop: one process
per GPU solves its contiguous slice with its own adaptive step sizes and its own loop trip
count.  Collectives run once, after the solve (NCCL over NVLink on the B200 box, gloo in
the CPU tests):

* ``all_gather`` of ``ys``, ``n_steps``, ``n_accepted``, ``n_initialized``, ``status``;
* ``all_reduce(MAX)`` of the per-rank loop iteration count, because the reference's
  ``n_f_evals`` is batch-uniform (``n_init + 6 * max_b n_steps``, terms.py:54-58).

"Any failure stops the whole batch" (adjoints.py:186-190) across shards: on the fused route
(``solve_sharded_symmetric``) every shard publishes its first failing iteration with a system-scope
atomic and every shard whose samples ran past the batch-wide first failure replays with that cap --
the gathered Solution equals the single-device one.  On the NCCL path (``solve_sharded``, opaque
vector fields) it holds per shard (a failing sample only cuts off the samples of its own rank).
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


class SymmetricWorkspace:
    """Gathered result buffers of a sharded fused solve in symmetric (peer-mapped) memory.

    One ``torch.distributed._symmetric_memory`` allocation per rank holds the gathered ``ys``, the
    four gathered int64 statistics and a 4-word global block; every rank maps every other rank's
    allocation over NVLink.  The fused solve kernel of rank r writes the results of its samples
    into rows ``[r * B_local, (r + 1) * B_local)`` of EVERY rank's buffers while it solves (with
    ``t_eval`` only the statistics: the dense-output block follows as one bulk copy per peer, see
    ``bulk_ys``) and publishes its iteration count with system-scope atomics: the all-gather that
    used to follow the solve is gone -- what is left between the ranks is two barriers around the
    launch (signal pads of the symmetric allocation, stream-ordered, no host sync).

    Equal shard sizes, one node (<= 8 ranks), built-in analytic fields (the fused route) only; the
    returned Solution aliases the workspace -- it is overwritten by the next solve that uses it."""

    MAX_PEERS = 8

    def __init__(self, local_batch: int, n_points: int, n_features: int, dtype: torch.dtype, device,
                 group=None):
        import torch.distributed._symmetric_memory as symm

        self.group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        assert self.world <= self.MAX_PEERS, "one node: at most 8 ranks"
        self.local_batch, self.n_points, self.n_features, self.dtype = local_batch, max(n_points, 1), n_features, dtype
        G = self.world * local_batch
        esz = torch.empty((), dtype=dtype).element_size()

        def up(n):
            return (n + 255) // 256 * 256

        self._off_ys = 0
        self._off_stats = up(G * self.n_points * n_features * esz)
        self._off_global = self._off_stats + up(4 * G * 8)
        total = self._off_global + 256
        self.buf = symm.empty(total, dtype=torch.uint8, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.buf.zero_()
        self.ys = self.buf[self._off_ys: self._off_ys + G * self.n_points * n_features * esz].view(dtype).view(
            G, self.n_points, n_features)
        stats = self.buf[self._off_stats: self._off_stats + 4 * G * 8].view(torch.long).view(4, G)
        self.n_steps, self.n_accepted, self.n_initialized, self.status = stats[0], stats[1], stats[2], stats[3]
        self.glob = self.buf[self._off_global: self._off_global + 16].view(torch.int32)
        self.bases = [int(p) for p in self.hdl.buffer_ptrs]
        # Dense output writes many small rows per sample: 8-byte stores scattered over NVLink are slow
        # (C3 shape, 2^20 samples per rank, 2 GPUs: 13.4 ms against 3.8 ms for the NCCL gather), so
        # with t_eval the kernel replicates only the statistics and this rank's finished ys block is
        # pushed to the peers in bulk (one device-to-device copy per peer, each on its own stream).
        self.bulk_ys = self.n_points > 1
        # ... over copy-engine pushes between TWO GPUs (750 GB/s; NCCL's all-gather: 515 GB/s).  With more
        # ranks every GPU has N - 1 pushes in flight in each direction and the copy engines get 300-340 GB/s per
        # rank (measured at N = 4 and 8), where NCCL's in-place all-gather of the same blocks reaches 590 GB/s:
        # from three ranks on the dense-output block goes through NCCL (profiles/r02_symmetric_8gpu.txt)
        # ... or through this repo's SM-store push kernel (tode_peer_push: every 16-byte vector read once from
        # HBM and stored to all peer mappings, every link busy at once); TORCHODE_B200_PUSH = copy | nccl | sm
        # overrides the choice (experiments)
        import os

        row_bytes = self.n_points * n_features * esz
        default = "copy" if self.world == 2 else ("sm" if row_bytes % 16 == 0 else "nccl")
        self.push_mode = os.environ.get("TORCHODE_B200_PUSH", default) if self.bulk_ys else "none"
        if self.push_mode == "sm" and row_bytes % 16 != 0:
            self.push_mode = "nccl"
        self.push_by_copy = self.push_mode in ("copy", "sm")  # pushes are per row block and need no collective
        self._ys_flat = self.buf[self._off_ys: self._off_ys + G * self.n_points * n_features * esz].view(dtype)
        if self.bulk_ys:
            lo = self.rank * local_batch
            self._ys_local = self.ys[lo: lo + local_batch]
            self._ys_peer = [None if p == self.rank else self.hdl.get_buffer(
                p, (local_batch, self.n_points, n_features), dtype, storage_offset=self._off_ys // esz + lo * self.n_points * n_features)
                for p in range(self.world)]
            self._push_streams = [torch.cuda.Stream(device) for _ in range(self.world - 1)]
        self.barrier()

    def matches(self, local_batch, n_points, n_features, dtype) -> bool:
        return (self.local_batch, self.n_points, self.n_features, self.dtype) == (
            local_batch, max(n_points, 1), n_features, dtype)

    def barrier(self):
        """Cross-GPU barrier on the current stream (signal pads of the symmetric allocation)."""
        self.hdl.barrier()

    def push_ys(self, rows=None):
        """Bulk copy of rows ``[a, b)`` (default: all) of this rank's ys block into every peer's gathered
        buffer, one device-to-device copy per peer on that peer's push stream, ordered after what is on
        the current stream now.  The current stream does NOT wait: the next chunk's solve overlaps the
        copies; ``wait_pushes`` joins them."""
        if not self.bulk_ys or not self.push_by_copy:
            return
        a, b = (0, self.local_batch) if rows is None else rows
        cur = torch.cuda.current_stream(self.buf.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        peers = [t for t in self._ys_peer if t is not None]
        if self.push_mode == "sm":
            import ctypes as C

            from . import _cabi, _launch

            stream = self._push_streams[0]
            stream.wait_event(ready)
            src = self._ys_local[a:b]
            dst = (C.c_void_p * len(peers))(*[t[a:b].data_ptr() for t in peers])
            with torch.cuda.stream(stream):
                _cabi.check(_cabi.lib().tode_peer_push(src.data_ptr(), dst, len(peers),
                                                       src.numel() * src.element_size(),
                                                       _launch.stream_ptr(self.buf.device)), "tode_peer_push")
            return
        for stream, dst in zip(self._push_streams, peers):
            stream.wait_event(ready)
            with torch.cuda.stream(stream):
                dst[a:b].copy_(self._ys_local[a:b], non_blocking=True)

    def wait_pushes(self):
        """The current stream waits for every bulk copy enqueued by ``push_ys`` -- or, with more than two
        ranks, runs the in-place NCCL all-gather of the ranks' dense-output blocks (this rank's block already
        sits at its place in the gathered buffer: no staging copy)."""
        if not self.bulk_ys:
            return
        if not self.push_by_copy:
            block = self.local_batch * self.n_points * self.n_features
            dist.all_gather_into_tensor(self._ys_flat, self._ys_flat[self.rank * block:(self.rank + 1) * block],
                                        group=self.group)
            return
        cur = torch.cuda.current_stream(self.buf.device)
        for stream in self._push_streams:
            done = torch.cuda.Event()
            done.record(stream)
            cur.wait_event(done)

    def _check(self, B, n_points, F, dtype, rows):
        a, b = (0, self.local_batch) if rows is None else rows
        assert (b - a == B and 0 <= a <= b <= self.local_batch
                and (self.n_points, self.n_features, self.dtype) == (max(n_points, 1), F, dtype)), \
            "workspace was built for another problem shape"
        return a, b

    def own_rows(self, B, n_points, F, dtype, rows=None):
        """Rows ``[a, b)`` (default: all) of this rank's block of its own gathered buffers (ys, n_steps,
        n_accepted, n_initialized, status): the primary outputs of the kernel launch."""
        a, b = self._check(B, n_points, F, dtype, rows)
        lo = self.rank * self.local_batch
        return (self.ys[lo + a:lo + b], self.n_steps[lo + a:lo + b], self.n_accepted[lo + a:lo + b],
                self.n_initialized[lo + a:lo + b], self.status[lo + a:lo + b])

    def fill(self, sol, B, n_points, F, dtype, rows=None):
        """Set the peer_* fields of a ``tode_solution`` whose primary outputs are ``own_rows`` (called by
        ``AutoDiffAdjoint._fused_launch``): the other ranks' replicas, and everybody's global block."""
        a, _ = self._check(B, n_points, F, dtype, rows)
        G = self.world * self.local_batch
        sol.n_peers, sol.peer_row0 = self.world, self.rank * self.local_batch + a
        for p, base in enumerate(self.bases):
            remote = p != self.rank
            sol.peer_ys[p] = base + self._off_ys if (remote and not self.bulk_ys) else None
            for k, name in enumerate(("peer_n_steps", "peer_n_accepted", "peer_n_initialized", "peer_status")):
                getattr(sol, name)[p] = base + self._off_stats + k * G * 8 if remote else None
            sol.peer_global[p] = base + self._off_global


def solve_sharded_symmetric(solver, local_problem: InitialValueProblem, ws: SymmetricWorkspace, *,
                            dt0: Optional[torch.Tensor] = None, ts: Optional[torch.Tensor] = None,
                            chunks: int = 1) -> Solution:
    """Solve this rank's shard with the fused kernel writing straight into every rank's gathered
    buffers (``ws``); returns the full-batch Solution (aliasing ``ws``).  Collective: every rank
    of the group must call it.  ``ts``: the full-batch evaluation times if the caller has them
    (else this rank's ``ts`` is gathered with NCCL).

    ``chunks`` (dense-output workloads): the shard is solved in that many row blocks, and the ys rows
    of block i travel to the peers (bulk copies on their own streams) while block i + 1 solves.

    A failure anywhere stops the WHOLE batch at that iteration (adjoints.py:186-190): every launch
    publishes its iteration count and its first failing iteration to all ranks (system-scope atomics
    of the kernel's epilogue); a shard whose samples ran past the batch-wide first failure replays
    with that cap."""
    from .adjoints import _INT32_MAX

    term_ = solver.step_method.term
    field = solver._fused_eligible(local_problem, term_) if term_ is not None else None
    if field is None:
        raise NotImplementedError("solve_sharded_symmetric needs a built-in analytic field (the fused route); "
                                  "use solve_sharded for opaque vector fields")
    B = local_problem.batch_size
    n_blocks = max(1, min(int(chunks), B)) if (ws.bulk_ys and ws.push_by_copy) else 1
    bounds = [shard_bounds(B, i, n_blocks) for i in range(n_blocks)]

    def block(a, b):
        if n_blocks == 1:
            return local_problem, dt0
        te = None if local_problem.t_eval is None else local_problem.t_eval[a:b]
        return (InitialValueProblem(local_problem.y0[a:b], local_problem.t_start[a:b], local_problem.t_end[a:b], te),
                None if dt0 is None else dt0[a:b])

    with torch.no_grad(), torch.cuda.device(local_problem.device):
        ws.glob.zero_()
        ws.barrier()  # every rank has reset its global block and is done with the previous results
        ctxs = []
        for a, b in bounds:
            prob_i, dt0_i = block(a, b)
            ctxs.append(solver._fused_launch(prob_i, term_, field, dt0_i, peers=ws, rows=(a, b)))
            ws.push_ys((a, b))
        ws.wait_pushes()
        ws.barrier()  # every rank's kernels (and their peer stores / atomics / bulk copies) have completed
        # one host read: this rank's global block followed by the summaries of its launches
        words = torch.cat([ws.glob[:4]] + [c["summary"][:4] for c in ctxs]).tolist()
        (g_iters, _, g_fail_enc, _), summaries = words[:4], [words[4 + 4 * i: 8 + 4 * i] for i in range(len(ctxs))]
        if g_fail_enc:
            # some sample of the batch failed: the reference stops everybody at that iteration
            g_first_fail = _INT32_MAX - g_fail_enc
            replay = any(sm[0] > g_first_fail for sm in summaries)
            if replay:
                for (a, b), c in zip(bounds, ctxs):
                    c["run"](g_first_fail)
                    ws.push_ys((a, b))
            if replay or not ws.push_by_copy:
                # (the NCCL gather of the dense-output blocks is a collective: every rank takes part, whether
                # its own shard replayed or not -- all ranks read the same global block, so all are here)
                ws.wait_pushes()
            ws.barrier()
            g_iters = min(g_iters, g_first_fail)
        if any(sm[2] for sm in summaries):
            raise NotImplementedError("non-monotone t_eval rows need the stage-wise route: use solve_sharded")
        launches = sum(sm[3] for sm in summaries)
    G = ws.world * ws.local_batch
    stats = {"n_steps": ws.n_steps, "n_accepted": ws.n_accepted, "n_initialized": ws.n_initialized}
    if getattr(term_, "with_stats", True):
        stats["n_f_evals"] = torch.full((1,), ctxs[0]["n_init_evals"] + ctxs[0]["n_stage_evals"] * g_iters,
                                        dtype=torch.long).expand(G)
    if ts is None:
        sizes = [ws.local_batch] * ws.world
        lts = local_problem.t_eval if local_problem.t_eval is not None else local_problem.t_end[:, None]
        ts = _gather_rows(lts, sizes, ws.group)
    solver.last_run = {"route": "fused+peer-stores", "iterations": g_iters, "blocks": n_blocks,
                       "kernel_launches": launches}
    return Solution(ts=ts, ys=ws.ys, stats=stats, status=ws.status)


def solve_sharded(solver, problem: InitialValueProblem, *, dt0: Optional[torch.Tensor] = None,
                  args=None, group=None, gather: bool = True,
                  workspace: Optional[SymmetricWorkspace] = None, chunks: int = 1) -> Solution:
    """Every rank holds (or can build) the full problem; each solves its slice.

    With ``gather=False`` the local Solution is returned (statistics of the slice only).  With a
    ``workspace`` (equal slices, built-in field) the fused kernel writes the gathered Solution
    itself (``solve_sharded_symmetric``) instead of the NCCL all-gathers."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(problem.batch_size, rank, world)
    local_problem = shard_problem(problem, rank, world)
    dt0_local = None if dt0 is None else dt0[lo:hi]
    if workspace is not None and gather:
        return solve_sharded_symmetric(solver, local_problem, workspace, dt0=dt0_local, ts=problem.t_eval,
                                       chunks=chunks)
    local = solver.solve(local_problem, dt0=dt0_local, args=args)
    if not gather:
        return local
    return gather_solution(local, problem.batch_size, ts=problem.t_eval, group=group)

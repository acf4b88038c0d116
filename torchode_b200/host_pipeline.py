"""Solve a problem whose tensors live in HOST memory, overlapping the PCIe transfers with the solve.

Not part of the reference's API (torchode solves where the tensors are); this is the call for a
user whose initial conditions arrive in host buffers and whose results are consumed on the host.
The batch is cut into contiguous chunks; every chunk gets its own CUDA stream on which its inputs
are copied in, the solve is enqueued and its results are copied out to pinned host memory, so the
copy-in of chunk i+1 and the copy-out of chunk i-1 run while chunk i computes, and the chunks'
kernels overlap at their tails.  One host synchronisation at the end.

Semantics: as ``solve_sharded`` -- every chunk is an independent solve ("any failure stops the
whole batch", adjoints.py:186-190, holds per chunk); ``n_f_evals`` is the maximum over the chunks
(the reference charges every sample the evaluations of the longest-running one).
"""
from typing import Any, Dict, List, Optional

import torch

from . import _cabi
from .adjoints import _INT32_MAX, AutoDiffAdjoint
from .distributed import shard_bounds
from .problems import InitialValueProblem
from .solution import Solution


def _pinned_like(shape, dtype) -> torch.Tensor:
    return torch.empty(shape, dtype=dtype, device="cpu", pin_memory=True)


def chunk_count(B: int, F: int, Tn: int, itemsize: int, chunks: int, min_chunk: int, min_chunk_bytes: int) -> int:
    """How many chunks ``solve_from_host`` runs: at most ``chunks``, each worth a stream of its own either by samples
    (``min_chunk``) or by bytes moved over PCIe (``min_chunk_bytes``: y0 in, ys out), never more than samples."""
    if B <= 0:
        return 1
    moved = B * F * (1 + max(Tn, 1)) * itemsize
    worth = max(B // max(1, int(min_chunk)), moved // max(1, int(min_chunk_bytes)))
    return max(1, min(int(chunks), worth, B))


def _kernel_field(solver, term_) -> bool:
    """Fields of this package whose solve is host-driven, whose code is ours (safe to run from two threads) and whose
    route records no CUDA graph (a capture in one thread does not tolerate allocations in another): Heat1D."""
    from .fields import Heat1D

    return isinstance(getattr(term_, "f", None), Heat1D) and not solver.use_cuda_graph


def _solve_chunks_threaded(solver, staged, bounds, streams, device, args, workers, outs):
    """Chunk i is solved by thread i % workers on stream i with that thread's own AutoDiffAdjoint (same step
    method and controller objects -- they are immutable --, own plans / poll rings); results leave on the chunk's
    stream.  ctypes and torch release the GIL around the calls that block."""
    import threading

    clones = solver.__dict__.setdefault("_host_workers", [])
    while len(clones) < workers:
        clones.append(AutoDiffAdjoint(solver.step_method, solver.step_size_controller))
    for c in clones[:workers]:  # follow the caller's settings of this call
        c.max_steps, c.lookahead = solver.max_steps, solver.lookahead
        c.use_cuda_graph, c.use_step_fusion = solver.use_cuda_graph, solver.use_step_fusion
        c.backprop_through_step_size_control = solver.backprop_through_step_size_control
    pending: List[Optional[Dict[str, Any]]] = [None] * len(bounds)
    errors: List[BaseException] = []

    def work(tid):
        try:
            with torch.no_grad(), torch.cuda.device(device):
                for i in range(tid, len(bounds), workers):
                    lo, hi = bounds[i]
                    prob_i, dt0_i = staged[i]
                    with torch.cuda.stream(streams[i]):
                        sol_i = clones[tid].solve(prob_i, dt0=dt0_i, args=args)
                        if hi > lo:
                            ys, status, n_steps, n_accepted, n_init = outs
                            for dst, src in ((ys, sol_i.ys), (status, sol_i.status),
                                             (n_steps, sol_i.stats["n_steps"]),
                                             (n_accepted, sol_i.stats["n_accepted"]),
                                             (n_init, sol_i.stats["n_initialized"])):
                                dst[lo:hi].copy_(src, non_blocking=True)
                    pending[i] = dict(sol=sol_i, copied=hi > lo, prob=prob_i, dt0=dt0_i)
        except BaseException as e:  # re-raised by the caller's thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(t,), name=f"torchode_b200-host-{t}") for t in range(workers)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    solver.last_run = dict(clones[0].last_run)
    return pending


def solve_from_host(solver: AutoDiffAdjoint, problem: InitialValueProblem, device, *, chunks: int = 8,
                    min_chunk: int = 4096, min_chunk_bytes: int = 128 << 20, workers: int = 2,
                    dt0: Optional[torch.Tensor] = None, args: Any = None,
                    out: Optional[Solution] = None) -> Solution:
    """``problem``: an InitialValueProblem over CPU tensors (pinned memory makes the copies
    asynchronous).  Returns a Solution over pinned CPU tensors.  ``out``: the Solution of an earlier
    call with the same shapes, whose buffers are reused (allocating pinned memory costs more than a
    solve).  A chunk is worth a stream of its own if it has ``min_chunk`` samples (small batches of narrow
    states are launch-bound: they run as one chunk) OR moves ``min_chunk_bytes`` over PCIe (few wide samples --
    64 x 4 MB rows of a method-of-lines grid -- are cut by bytes: their copies are what there is to hide).  The
    default of 128 MB per chunk is what configs[4] measures (scripts/c5_e2e_chunks.py, 268 MB each way, B200): one
    host thread 22.0 ms in one piece, 20.7 in two chunks, 22.3 in four, 27.1 in eight -- a chunk whose solve drives
    its loop from the host pays that loop's latency again; two host threads 20.2 / 19.6 / 22.9 ms in two / four /
    eight chunks.
    ``workers``: host threads driving such chunks (``fields.Heat1D`` only: this package's own code on a route that
    records no CUDA graph; a user's ``f`` is never called from a second thread)."""
    device = torch.device(device)
    term_ = solver.step_method.term
    assert term_ is not None, "solve_from_host needs the ODE term on the step method"
    B, F, Tn = problem.batch_size, problem.n_features, problem.n_evaluation_points
    D = problem.data_dtype
    chunks = chunk_count(B, F, Tn, torch.empty((), dtype=D).element_size(), chunks, min_chunk, min_chunk_bytes)
    bounds = [shard_bounds(B, i, chunks) for i in range(chunks)]
    reuse = (out is not None and out.ys.shape == (B, max(Tn, 1), F) and out.ys.dtype == D
             and out.ys.is_pinned() and out.status.shape == (B,))
    if reuse:
        ys, status = out.ys, out.status
        n_steps, n_accepted, n_init = (out.stats[k] for k in ("n_steps", "n_accepted", "n_initialized"))
    else:
        ys = _pinned_like((B, max(Tn, 1), F), D)
        status, n_steps, n_accepted, n_init = (_pinned_like((B,), torch.long) for _ in range(4))
    summaries = torch.empty((chunks, _cabi.SUMMARY_WORDS), dtype=torch.int32, device="cpu", pin_memory=True)
    te_host = problem.t_eval
    te_broadcast = te_host is not None and te_host.stride(0) == 0

    def to_dev(t, lo, hi):
        return None if t is None else t[lo:hi].to(device, non_blocking=True)

    pending: List[Optional[Dict[str, Any]]] = []
    # the chunk streams live with the solver: the caching allocator keeps one pool per stream, fresh
    # streams would mean fresh cudaMallocs on every call
    streams = solver._host_streams.setdefault(str(device), [])
    while len(streams) < chunks:
        streams.append(torch.cuda.Stream(device))
    with torch.no_grad(), torch.cuda.device(device):
        te_dev_row = te_host[:1].to(device, non_blocking=True) if te_broadcast else None
        ready = torch.cuda.Event()
        ready.record()
        staged: List[Any] = [None] * chunks

        def stage(i):
            if staged[i] is None:
                lo, hi = bounds[i]
                with torch.cuda.stream(streams[i]):
                    streams[i].wait_event(ready)
                    t_eval = te_dev_row.expand(hi - lo, -1) if te_broadcast else to_dev(te_host, lo, hi)
                    staged[i] = (InitialValueProblem(to_dev(problem.y0, lo, hi), to_dev(problem.t_start, lo, hi),
                                                     to_dev(problem.t_end, lo, hi), t_eval), to_dev(dt0, lo, hi))
            return staged[i]

        # A chunk whose solve drives its loop from the host (opaque f, the kernel-backed fields) would hold back the
        # copy-ins of all later chunks: those are queued before the first solve.  Whole-solve kernels are launched
        # without a host synchronisation: their chunks keep copy-in, launch, copy-out interleaved in issue order
        # (queueing eight chunks' copies first delays the first launch by ~0.3 ms of host time).
        host_driven = B > 0 and solver._fused_eligible(stage(0)[0], term_) is None
        if host_driven:
            for i in range(chunks):
                stage(i)
        if host_driven and chunks > 1 and workers > 1 and _kernel_field(solver, term_):
            # the solve of a kernel-backed field drives its loop from the host (look-ahead launches, a polled control
            # block, a few synchronisations around it): two host threads, each with a solver of its own over the
            # same components, keep two chunks' loops going so that one chunk's host latency is the other's GPU time
            pending = _solve_chunks_threaded(solver, staged, bounds, streams, device, args, workers,
                                             (ys, status, n_steps, n_accepted, n_init))
            staged_done = True
        else:
            staged_done = False
        for i, (lo, hi) in enumerate(bounds):
            if staged_done:
                break
            prob_i, dt0_i = stage(i)
            with torch.cuda.stream(streams[i]):
                field = solver._fused_eligible(prob_i, term_) if hi > lo else None
                if field is None:
                    # opaque f / plug-ins / empty chunk: the solve synchronises with the host itself; its results
                    # leave on the chunk's stream while the next chunk's loop runs
                    sol_i = solver.solve(prob_i, dt0=dt0_i, args=args)
                    ctx = dict(sol=sol_i, copied=hi > lo)
                    if hi > lo:
                        for dst, src in ((ys, sol_i.ys), (status, sol_i.status), (n_steps, sol_i.stats["n_steps"]),
                                         (n_accepted, sol_i.stats["n_accepted"]),
                                         (n_init, sol_i.stats["n_initialized"])):
                            dst[lo:hi].copy_(src, non_blocking=True)
                else:
                    ctx = solver._fused_launch(prob_i, term_, field, dt0_i)
                    summaries[i].copy_(ctx["summary"], non_blocking=True)
                    for dst, src in ((ys, ctx["ys"]), (status, ctx["status"]), (n_steps, ctx["n_steps"]),
                                     (n_accepted, ctx["n_accepted"]), (n_init, ctx["n_init"])):
                        dst[lo:hi].copy_(src, non_blocking=True)
                ctx["prob"], ctx["dt0"] = prob_i, dt0_i
                pending.append(ctx)
        n_f_evals = 0
        for i, (lo, hi) in enumerate(bounds):
            streams[i].synchronize()
            ctx = pending[i]
            sol_i = ctx.get("sol")
            if sol_i is None:
                iters, first_fail, nonmono = summaries[i].tolist()[:3]
                if nonmono or (first_fail != _INT32_MAX and first_fail < iters):
                    # rare: replay after a failure / general t_eval mode -> finish this chunk in order
                    with torch.cuda.stream(streams[i]):
                        sol_i = solver._fused_finish(ctx)
                        if sol_i is None:
                            sol_i = solver._solve_staged(ctx["prob"], term_, ctx["dt0"], args)
                    streams[i].synchronize()
                else:
                    n_f_evals = max(n_f_evals, ctx["n_init_evals"] + ctx["n_stage_evals"] * iters)
            if sol_i is not None and hi > lo:
                if not ctx.get("copied"):  # a replayed chunk: its results are final only now
                    ys[lo:hi].copy_(sol_i.ys)
                    status[lo:hi].copy_(sol_i.status)
                    n_steps[lo:hi].copy_(sol_i.stats["n_steps"])
                    n_accepted[lo:hi].copy_(sol_i.stats["n_accepted"])
                    n_init[lo:hi].copy_(sol_i.stats["n_initialized"])
                if "n_f_evals" in sol_i.stats:
                    n_f_evals = max(n_f_evals, int(sol_i.stats["n_f_evals"][0]))
    stats: Dict[str, Any] = {}
    if getattr(term_, "with_stats", True):
        stats["n_f_evals"] = torch.full((1,), n_f_evals, dtype=torch.long).expand(B)
    stats["n_steps"], stats["n_accepted"], stats["n_initialized"] = n_steps, n_accepted, n_init
    ts = problem.t_eval if problem.t_eval is not None else problem.t_end[:, None]
    solver.last_run = {"route": "host-pipelined", "chunks": chunks, "host_threads": workers if staged_done else 1,
                       "kernel_launches": 2 * chunks}
    return Solution(ts=ts, ys=ys, stats=stats, status=status)

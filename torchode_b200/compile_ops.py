"""torch.compile support (the reference advertises JIT compatibility, README.md:5-7).

The solve loop itself cannot be traced -- it launches kernels through ctypes and polls a device
control block -- and it does not need to be: inside a compiled region ``AutoDiffAdjoint.solve``
dispatches to ONE opaque custom operator, ``torchode_b200::solve``, whose fake (meta) kernel tells
the tracer the output shapes and whose real kernel is the eager CUDA route.  A model that
contains a solve therefore compiles without a graph break around it (``fullgraph=True`` works);
nothing of the hot path goes through the compiler.

The operator is forward-only: under ``torch.compile`` a problem that needs gradients raises
(train in eager mode, where the recorded forward / recompute backward of autodiff.py applies).
"""
import weakref
from typing import List, Optional, Tuple

import torch

from .problems import InitialValueProblem
from .solution import Solution

_SOLVERS = weakref.WeakValueDictionary()  # handle -> AutoDiffAdjoint (the op takes plain ints)


def solver_handle(solver) -> int:
    """Registered when the solver is constructed (not while tracing)."""
    h = id(solver)
    _SOLVERS[h] = solver
    return h


@torch.library.custom_op("torchode_b200::solve", mutates_args=())
def _solve_op(y0: torch.Tensor, t_start: torch.Tensor, t_end: torch.Tensor, t_eval: Optional[torch.Tensor],
              dt0: Optional[torch.Tensor], handle: int) -> List[torch.Tensor]:
    solver = _SOLVERS[handle]
    with torch.no_grad():
        sol = solver.solve(InitialValueProblem(y0, t_start, t_end, t_eval), dt0=dt0)
    n_f = sol.stats.get("n_f_evals")
    n_f = torch.zeros(0, dtype=torch.long) if n_f is None else n_f.contiguous()
    # custom-op outputs must not alias inputs or each other
    return [sol.ys.clone() if sol.ys.data_ptr() == y0.data_ptr() else sol.ys, sol.stats["n_steps"],
            sol.stats["n_accepted"], sol.stats["n_initialized"], sol.status, n_f]


@_solve_op.register_fake
def _(y0, t_start, t_end, t_eval, dt0, handle):
    B, F = y0.shape
    T = 1 if t_eval is None else t_eval.shape[1]
    solver = _SOLVERS[handle]
    with_stats = getattr(solver.step_method.term, "with_stats", True)

    def counts():
        return torch.empty((B,), dtype=torch.long, device=y0.device)

    return [y0.new_empty((B, T, F)), counts(), counts(), counts(), counts(),
            torch.empty((B if with_stats else 0,), dtype=torch.long, device="cpu")]


def _no_backward(ctx, *grads):
    raise NotImplementedError(
        "torchode_b200::solve is forward-only under torch.compile; differentiate through the solver in "
        "eager mode (AutoDiffAdjoint records the CUDA forward and recomputes for the backward pass)")


torch.library.register_autograd("torchode_b200::solve", _no_backward)


def solve_compiled(solver, problem: InitialValueProblem, dt0: Optional[torch.Tensor]) -> Solution:
    """What ``AutoDiffAdjoint.solve`` runs while a tracer is compiling the caller."""
    ys, n_steps, n_accepted, n_init, status, n_f = torch.ops.torchode_b200.solve(
        problem.y0, problem.t_start, problem.t_end, problem.t_eval, dt0, solver._compile_handle)
    stats = {"n_steps": n_steps, "n_accepted": n_accepted, "n_initialized": n_init}
    if n_f.shape[0]:
        stats["n_f_evals"] = n_f
    ts = problem.t_eval if problem.t_eval is not None else problem.t_end[:, None]
    return Solution(ts=ts, ys=ys, stats=stats, status=status)

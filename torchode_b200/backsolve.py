"""Continuous-adjoint ("backsolve") gradients on top of the CUDA solve loop
(API of torchode/adjoints.py:343-681: ``BacksolveAdjoint``, ``JointBacksolveAdjoint``).

Both the forward solve and the backward solve of the augmented system are plain forward solves
under ``no_grad`` -- exactly what the kernels provide -- so this is the way to train through the
B200 solve loop: the forward pass takes the fused / stage-wise kernels, the backward pass
integrates ``[a_t, y, a_y, a_theta]`` backwards in time through the stage-wise kernels with the
vector-Jacobian products of the user's ``f`` as the (opaque) vector field.
"""
from typing import Any, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .adjoints import AutoDiffAdjoint
from .problems import InitialValueProblem
from .solution import Solution
from .terms import ODETerm


def _pack(parts: Sequence[torch.Tensor]) -> Tuple[List[torch.Size], torch.Tensor]:
    """Concatenate per-sample tensors along the feature axis; remember their trailing shapes."""
    return [p.shape[1:] for p in parts], torch.cat([p.reshape(p.shape[0], -1) for p in parts], dim=1)


def _unpack(shapes: Sequence[torch.Size], flat: torch.Tensor) -> List[torch.Tensor]:
    sizes = [int(torch.Size(s).numel()) for s in shapes]
    return [chunk.reshape((-1, *shape)) for shape, chunk in zip(shapes, torch.split(flat, sizes, dim=1))]


class _AdjointSolve(torch.autograd.Function):
    """forward: solve; backward: integrate the augmented adjoint system from t_end to t_start,
    interval by interval when intermediate evaluation points carry gradients."""

    @staticmethod
    def forward(ctx, fwd_loop, bwd_loop, term, aug_term, y0, t_start, t_end, t_eval, dt0, bwd_dt0, args,
                *params):
        with torch.no_grad():
            sol = fwd_loop.solve(InitialValueProblem(y0, t_start, t_end, t_eval), term=term, dt0=dt0, args=args)
        ctx.bwd_loop, ctx.aug_term, ctx.args, ctx.stats = bwd_loop, aug_term, args, sol.stats
        ctx.save_for_backward(t_start, t_end, t_eval, bwd_dt0, sol.ys, *params)
        ctx.mark_non_differentiable(sol.status)
        return sol.ts, sol.ys, sol.stats, sol.status

    @staticmethod
    def backward(ctx, _g_ts, g_ys, _g_stats, _g_status):
        t_start, t_end, t_eval, dt0, ys, *params = ctx.saved_tensors
        loop, aug_term, stats = ctx.bwd_loop, ctx.aug_term, ctx.stats
        log = stats.setdefault("backsolve", [])
        B, n_eval, F = ys.shape
        shapes, state = _pack([ys.new_zeros((B, 1)), ys[:, -1], g_ys[:, -1]]
                              + [torch.zeros_like(p).expand(B, *p.shape) for p in params])
        # segment boundaries, walked from the last evaluation point back to the first
        if t_eval is None:
            segments = [(t_end, t_start, None)]
        else:
            segments = [(t_eval[:, i], t_eval[:, i - 1], i - 1) for i in range(n_eval - 1, 0, -1)]
        with torch.no_grad():
            for seg_start, seg_end, landed_on in segments:
                sol = loop.solve(InitialValueProblem(state.contiguous(), seg_start, seg_end), term=aug_term,
                                 dt0=dt0, args=(shapes, ctx.args))
                log.append(sol.stats)
                state = sol.ys[:, -1].clone()
                if landed_on is not None:
                    # re-anchor y on the stored forward solution and pick up that point's gradient
                    state[:, 1:1 + F] = ys[:, landed_on]
                    state[:, 1 + F:1 + 2 * F] += g_ys[:, landed_on]
        _a_t, _y, a_y0, *a_params = _unpack(shapes, state)
        return (None, None, None, None, a_y0, None, None, None, None, None, None,
                *[a.sum(dim=0) for a in a_params])


class _PerSampleAugmentedField:
    """f_aug(t, [a_t, y, a_y, a_theta], (shapes, args)) with per-sample VJPs (torch.func.vmap)."""

    def __init__(self, term: ODETerm, vmap_args_dims, vmap_randomness: str):
        f = term.f
        assert isinstance(f, nn.Module), "BacksolveAdjoint needs the dynamics as an nn.Module"
        names = list(dict(f.named_parameters()).keys())
        values = tuple(dict(f.named_parameters()).values())
        buffers = dict(f.named_buffers())
        with_args = term.with_args

        def one_sample(t_i, y_i, a_y_i, arg_i):
            def call(theta, t_, y_):
                inputs = (t_, y_, arg_i) if with_args else (t_, y_)
                return torch.func.functional_call(f, (dict(zip(names, theta)), buffers), inputs)

            dy, pullback = torch.func.vjp(call, values, t_i, y_i)
            g_theta, g_t, g_y = pullback(-a_y_i)
            return dy, g_t, g_y, g_theta

        self._vjp = torch.func.vmap(one_sample, in_dims=(0, 0, 0, vmap_args_dims), randomness=vmap_randomness)

    def __call__(self, t, state, packed_args):
        shapes, args = packed_args
        _a_t, y, a_y, *_rest = _unpack(shapes, state)
        dy, g_t, g_y, g_theta = self._vjp(t, y, a_y, args)
        return _pack([g_t[:, None], dy, g_y, *g_theta])[1]


class BacksolveAdjoint(nn.Module):
    """Gradients w.r.t. ``y0`` and the parameters of ``term.f`` by solving the adjoint equation
    backwards in time; every sample keeps its own adaptive step size in both directions."""

    def __init__(self, term: ODETerm, step_method, step_size_controller, vmap_args_dims=None,
                 vmap_randomness: str = "error"):
        super().__init__()
        self.term = term
        self.augmented_term = ODETerm(_PerSampleAugmentedField(term, vmap_args_dims, vmap_randomness),
                                      with_stats=term.with_stats, with_args=True)
        self.forward_adjoint = AutoDiffAdjoint(step_method, step_size_controller)
        self.backward_adjoint = AutoDiffAdjoint(step_method, step_size_controller)

    def solve(self, problem: InitialValueProblem, term: Optional[ODETerm] = None,
              dt0: Optional[torch.Tensor] = None, args: Any = None,
              backward_dt0: Optional[torch.Tensor] = None) -> Solution:
        if backward_dt0 is None and dt0 is not None:
            backward_dt0 = -dt0
        ts, ys, stats, status = _AdjointSolve.apply(
            self.forward_adjoint, self.backward_adjoint, self.term, self.augmented_term, problem.y0,
            problem.t_start, problem.t_end, problem.t_eval, dt0, backward_dt0, args,
            *list(self.term.parameters()))
        return Solution(ts, ys, stats, status)

    def __repr__(self):
        return (f"BacksolveAdjoint(term={self.term!r}, forward_adjoint={self.forward_adjoint!r}, "
                f"backward_adjoint={self.backward_adjoint!r})")


class _WholeBatchField(nn.Module):
    """Presents the whole batch as ONE sample on the time axis of the first instance: the other
    instances' intervals are mapped onto it linearly and dy is rescaled by the slope
    (substitution rule).  ``args = (batch, intercept, slope, inner_args)``."""

    def __init__(self, term: ODETerm):
        super().__init__()
        self.term = term

    def forward(self, t, y, args):
        batch, intercept, slope, inner = args
        t_inner = torch.addcmul(intercept, slope, t)
        y_inner = y[0].reshape(batch, -1)
        dy = self.term.f(t_inner, y_inner, inner) if self.term.with_args else self.term.f(t_inner, y_inner)
        return (dy * slope[:, None]).flatten()[None]


class _JointAugmentedField:
    """f_aug for the whole-batch formulation, VJPs through ordinary autograd."""

    def __init__(self, whole: _WholeBatchField):
        self.whole = whole
        self.params = list(whole.parameters())

    def __call__(self, t, state, packed_args):
        shapes, inner = packed_args
        _a_t, y, a_y, *_rest = _unpack(shapes, state)
        with torch.enable_grad():
            t_, y_ = t.detach().requires_grad_(), y.detach().requires_grad_()
            dy = self.whole(t_, y_, inner)
            grads = torch.autograd.grad(dy, [t_, y_] + self.params, -a_y, allow_unused=True)
        g_t = grads[0] if grads[0] is not None else torch.zeros_like(t)
        g_y = grads[1] if grads[1] is not None else torch.zeros_like(y)
        g_p = [g if g is not None else torch.zeros_like(p) for p, g in zip(self.params, grads[2:])]
        return _pack([g_t[:, None], dy.detach(), g_y, *[g[None] for g in g_p]])[1]


class JointBacksolveAdjoint(nn.Module):
    """Backsolve adjoint that integrates the whole batch as a single ODE (one shared step size),
    for dynamics that do not vmap (adjoints.py:613-681)."""

    def __init__(self, term: ODETerm, step_method, step_size_controller):
        super().__init__()
        self.whole = _WholeBatchField(term)
        self.term = ODETerm(self.whole, with_stats=term.with_stats, with_args=True)
        self.augmented_term = ODETerm(_JointAugmentedField(self.whole), with_stats=term.with_stats,
                                      with_args=True)
        self.forward_loop = AutoDiffAdjoint(step_method, step_size_controller)
        self.backward_loop = AutoDiffAdjoint(step_method, step_size_controller)

    def solve(self, problem: InitialValueProblem, term: Optional[ODETerm] = None,
              dt0: Optional[torch.Tensor] = None, args: Any = None,
              backward_dt0: Optional[torch.Tensor] = None) -> Solution:
        t_eval = problem.t_eval
        if t_eval is not None:
            steps = torch.diff(t_eval, dim=1)
            rel = steps / torch.maximum(steps[:, :1], t_eval.new_tensor(1e-8))
            assert (rel - rel[0]).abs().max() < 1e-8, (
                "JointBacksolveAdjoint can only be applied if all instances in the batch are evaluated "
                "at the same points in time")
            t_eval = t_eval[:1]
        span = problem.t_end - problem.t_start
        slope = span / span[0]
        intercept = problem.t_start - slope * problem.t_start[0]
        if backward_dt0 is None and dt0 is not None:
            backward_dt0 = -dt0
        _, ys, stats, status = _AdjointSolve.apply(
            self.forward_loop, self.backward_loop, self.term, self.augmented_term,
            problem.y0.flatten()[None], problem.t_start[:1], problem.t_end[:1], t_eval, dt0, backward_dt0,
            (problem.batch_size, intercept, slope, args), *list(self.whole.parameters()))
        ys = ys[0].unflatten(dim=1, sizes=(problem.batch_size, -1)).transpose(1, 0)
        ts = problem.t_end if problem.t_eval is None else problem.t_eval
        return Solution(ts, ys, stats, status)

    def __repr__(self):
        return (f"JointBacksolveAdjoint(term={self.term!r}, forward_loop={self.forward_loop!r}, "
                f"backward_loop={self.backward_loop!r})")

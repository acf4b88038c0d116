"""Gradients through ``AutoDiffAdjoint.solve`` by recompute-based backward.

The forward pass is the ordinary CUDA solve loop (stage-wise route) run under ``no_grad`` while
recording the solver state at the start of every iteration.  The backward pass walks the recorded
iterations in reverse: each one is re-traced from its recorded inputs with differentiable PyTorch
ops (stage combinations, the user's ``f``, error ratio, step-size controller, commit masks, dense
output -- the arithmetic of adjoints.py:135-255, runge_kutta.py:246-279 and
step_size_controllers.py:394-429 / 598-620) and differentiated with ``torch.autograd.grad``.
All discrete decisions (accept, running, which t_eval points were crossed) are taken from the
forward record, so the gradient is the exact derivative of the computation the kernels performed:
the same quantity the reference obtains by back-propagating through its eager loop, including the
dependence on the adaptive step sizes unless ``backprop_through_step_size_control=False``.
"""
from typing import Any, Dict, List, Optional

import torch

from . import _cabi
from .problems import InitialValueProblem
from .solution import Solution


class _Record:
    """Solver state at the start of every launched iteration (+ the final state)."""

    FIELDS = ("t", "dt", "y", "f0", "r1", "r2", "running", "n_accepted", "cursor", "status")

    def __init__(self):
        self.snaps: List[Dict[str, Optional[torch.Tensor]]] = []
        self.init: Dict[str, Any] = {}

    def snapshot(self, st):
        self.snaps.append({k: (None if getattr(st, k) is None else getattr(st, k).clone()) for k in self.FIELDS})


def _norm(x, kind):
    # the reference's own norm functions: their backward is well defined at x == 0
    from .step_size_controllers import max_norm, rms_norm

    return max_norm(x) if kind == _cabi.NORM_MAX else rms_norm(x)


def _quartic(tab, method_interp, dtD, y, y1, k):
    """Coefficients (a, b, c, d, e) of the dense output, dopri5.py:54-60 / tsit5.py:124-139."""
    if method_interp == _cabi.INTERP_DOPRI5:
        f0, f1 = dtD * k[0], dtD * k[-1]
        ymid = y + dtD * sum(w * ks for w, ks in zip(tab["w"][0], k))
        a = 2 * (f1 - f0) - 8 * (y1 + y) + 16 * ymid
        b = 5 * f0 - 3 * f1 + 18 * y + 14 * y1 - 32 * ymid
        c = f1 - 4 * f0 - 11 * y - 5 * y1 + 16 * ymid
        return a, b, c, f0, y
    rows = [dtD * sum(w * ks for w, ks in zip(tab["w"][r], k)) for r in range(3)]
    return rows[2], rows[1], rows[0], dtD * k[0], y


def _eval_quartic(co, x):
    a, b, c, d, e = co
    return (((a * x + b) * x + c) * x + d) * x + e


class _ReplaySolve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, solver, term_, problem, dt0, args, n_params, y0, *params):
        rec = _Record()
        with torch.no_grad():
            detached = InitialValueProblem(y0.detach(), problem.t_start, problem.t_end, problem.t_eval)
            sol = solver._solve_staged(detached, term_, dt0, args, record=rec)
        if solver.last_run.get("general"):
            raise NotImplementedError("gradients with t_eval rows that are not monotone in time")
        ctx.solver, ctx.term, ctx.problem, ctx.dt0, ctx.args, ctx.rec = solver, term_, problem, dt0, args, rec
        ctx.n_iters = solver.last_run["iterations"]
        ctx.params = list(params[:n_params])
        ctx.save_for_backward(y0)
        solver._last_solution = sol
        ctx.mark_non_differentiable(sol.status)
        return sol.ys, sol.status

    @staticmethod
    def backward(ctx, g_ys, _g_status):
        solver, term_, problem, rec, args = ctx.solver, ctx.term, ctx.problem, ctx.rec, ctx.args
        (y0,) = ctx.saved_tensors
        params = ctx.params
        method, ctrl = solver.step_method, solver.step_size_controller
        diff_dt = solver.backprop_through_step_size_control
        D, Tt, dev = problem.data_dtype, problem.time_dtype, problem.device
        B, F, Tn = problem.batch_size, problem.n_features, problem.n_evaluation_points
        tb = method.tableau
        S = tb.n_stages
        tab = {
            "a": tb.a.to(dev, D), "c": tb.c.to(dev, Tt), "b_err": tb.b_err.to(dev, D),
            "w": None if tb.b_other is None else [[w for w in row] for row in tb.b_other.to(dev, D)],
        }
        cab_c = ctrl.to_cabi(method.convergence_order(), D, solver.max_steps)
        interp_kind = method.INTERP_ID
        t_start, t_end, t_eval = problem.t_start, problem.t_end, problem.t_eval
        sign = problem.time_direction.to(Tt)
        lo, hi = torch.minimum(t_start, t_end), torch.maximum(t_start, t_end)
        atol, rtol = float(cab_c.atol), float(cab_c.rtol)
        order = method.convergence_order()
        stats: Dict[str, Any] = {}
        term_.init(problem, stats)

        def vf(t, y):
            return term_.vf(t, y, stats, args)

        def controller(dt, ratio, r1, r2):
            factor = cab_c.safety * ratio ** cab_c.exp_ratio
            if cab_c.pid:
                factor = factor * r1 ** cab_c.exp_prev * r2 ** cab_c.exp_prev2
            # clamp without routing a zero gradient through pow at the almost_zero floor (0 * inf)
            inside = (factor >= cab_c.factor_min) & (factor <= cab_c.factor_max)
            factor = torch.where(inside, factor, torch.clamp(factor.detach(), cab_c.factor_min, cab_c.factor_max))
            dt_next = dt * factor.to(Tt)
            if cab_c.has_dt_min or cab_c.has_dt_max:
                mag = torch.clamp(dt_next.abs(), cab_c.dt_min if cab_c.has_dt_min else None,
                                  cab_c.dt_max if cab_c.has_dt_max else None)
                dt_next = torch.sign(dt_next) * mag
            return dt_next

        # adjoints of the loop-carried state
        z = lambda *shape, dtype=D: torch.zeros(*shape, dtype=dtype, device=dev)
        a_t, a_dt, a_y, a_f0, a_r1, a_r2 = z(B, dtype=Tt), z(B, dtype=Tt), z(B, F), z(B, F), z(B), z(B)
        a_params = [torch.zeros_like(p) for p in params]
        snaps = rec.snaps
        for i in range(ctx.n_iters - 1, -1, -1):
            s0, s1 = snaps[i], snaps[i + 1]
            running = s0["running"].bool()
            if not bool(running.any()):
                continue
            upd = (s1["n_accepted"] > s0["n_accepted"])
            running_new = s1["running"].bool()
            with torch.enable_grad():
                t = s0["t"].clone().requires_grad_()
                dt = s0["dt"].clone().requires_grad_()
                y = s0["y"].clone().requires_grad_()
                f0 = s0["f0"].clone().requires_grad_()
                r1 = (s0["r1"] if cab_c.pid else torch.ones(B, dtype=D, device=dev)).clone().requires_grad_()
                r2 = (s0["r2"] if cab_c.pid else torch.ones(B, dtype=D, device=dev)).clone().requires_grad_()
                dtD = dt.to(D)[:, None]
                k = [f0]
                y_i = y
                for s in range(1, S):
                    acc = sum(tab["a"][s, j] * k[j] for j in range(s))
                    y_i = y + dtD * acc
                    k.append(vf(t + tab["c"][s] * dt, y_i))
                y1 = y_i
                err = dtD * sum(tab["b_err"][s] * k[s] for s in range(S))
                bounds = atol + rtol * torch.maximum(y.abs(), y1.abs())
                ratio = torch.clamp(_norm(err.abs() / bounds, cab_c.norm), min=cab_c.almost_zero)
                dt_next = controller(dt, ratio, r1, r2)
                if not diff_dt:
                    dt_next = dt_next.detach()
                t_new = torch.where(upd, t + dt, t)
                y_new = torch.where(upd[:, None], y1, y)
                f0_new = torch.where(upd[:, None], k[-1], f0)
                keep = running_new & upd
                r1_new = torch.where(keep, ratio, r1)
                r2_new = torch.where(keep, r1, r2)
                dt_sel = torch.where(running_new, dt_next, dt)
                dt_new = torch.maximum(torch.minimum(dt_sel, hi - t_new), lo - t_new)
                # frozen (already finished) samples: identity
                t_new = torch.where(running, t_new, t)
                dt_new = torch.where(running, dt_new, dt)
                y_new = torch.where(running[:, None], y_new, y)
                f0_new = torch.where(running[:, None], f0_new, f0)
                r1_new = torch.where(running, r1_new, r1)
                r2_new = torch.where(running, r2_new, r2)
                outs = [t_new, dt_new, y_new, f0_new, r1_new, r2_new]
                gouts = [a_t, a_dt, a_y, a_f0, a_r1, a_r2]
                # dense output produced by this iteration
                if Tn > 0:
                    cols = torch.arange(Tn, device=dev)[None, :]
                    hit = (cols >= s0["cursor"][:, None]) & (cols < s1["cursor"][:, None]) & running[:, None]
                    rows, cj = hit.nonzero(as_tuple=True)
                    if rows.numel():
                        co = _quartic(tab, interp_kind, dtD, y, y1, k)
                        h = (t + dt) - t
                        h = torch.where(h.abs() > 0, h, torch.ones_like(h))
                        x = ((t_eval[rows, cj] - t[rows]) / h[rows]).to(D)[:, None]
                        outs.append(_eval_quartic([c_[rows] for c_ in co], x))
                        gouts.append(g_ys[rows, cj])
                else:
                    ended = running & (~running_new | (s1["status"] != 0))
                    rows = ended.nonzero(as_tuple=True)[0]
                    if rows.numel():
                        co = _quartic(tab, interp_kind, dtD, y, y1, k)
                        h = (t + dt) - t
                        h = torch.where(h.abs() > 0, h, torch.ones_like(h))
                        x = ((t_end[rows] - t[rows]) / h[rows]).to(D)[:, None]
                        outs.append(_eval_quartic([c_[rows] for c_ in co], x))
                        gouts.append(g_ys[rows, 0])
                leaves = [t, dt, y, f0, r1, r2] + params
                grads = torch.autograd.grad(outs, leaves, gouts, allow_unused=True)
            g = [gr if gr is not None else torch.zeros_like(lf) for gr, lf in zip(grads, leaves)]
            a_t, a_dt, a_y, a_f0, a_r1, a_r2 = g[:6]
            for ap, gp in zip(a_params, g[6:]):
                ap += gp

        # ---- initialisation: f0 = f(t0, y0), initial step size (step_size_controllers.py:431-490)
        with torch.enable_grad():
            y0_ = y0.detach().clone().requires_grad_()
            f0 = vf(t_start, y0_)
            outs, gouts = [f0, y0_ * 1.0], [a_f0, a_y]
            if ctx.dt0 is None and diff_dt:
                inv = 1.0 / (atol + rtol * y0_.abs())
                d0, d1 = _norm(y0_ * inv, cab_c.norm), _norm(f0 * inv, cab_c.norm)
                small = (d0 < 1e-5) | (d1 < 1e-5)
                h0 = torch.where(small, torch.full_like(d0, 1e-6), 0.01 * d0 / d1)
                h0 = torch.minimum(h0, (t_end - t_start).abs().to(D))
                y1 = y0_ + (sign.to(D) * h0)[:, None] * f0
                f1 = vf(t_start + sign * h0.to(Tt), y1)
                d2 = _norm((f1 - f0) * inv, cab_c.norm) / h0
                m = torch.maximum(d1, d2)
                h1 = torch.where(m <= 1e-15, torch.clamp(h0 * 1e-3, min=1e-6), (0.01 / m) ** (1.0 / order))
                dt_init = (sign.to(D) * torch.minimum(100 * h0, h1)).to(Tt)
                dt_init = torch.maximum(torch.minimum(dt_init, hi - t_start), lo - t_start)
                outs.append(dt_init)
                gouts.append(a_dt)
            grads = torch.autograd.grad(outs, [y0_] + params, gouts, allow_unused=True)
        g_y0 = grads[0] if grads[0] is not None else torch.zeros_like(y0)
        for ap, gp in zip(a_params, grads[1:]):
            if gp is not None:
                ap += gp
        if Tn > 0:  # evaluation exactly at t_start copies y0 (adjoints.py:123-126)
            at_start = t_eval[:, 0] == t_start
            g_y0 = g_y0 + torch.where(at_start[:, None], g_ys[:, 0], torch.zeros_like(g_ys[:, 0]))
        return (None, None, None, None, None, None, g_y0, *a_params)


def grad_leaves(term_, problem: InitialValueProblem, args) -> List[torch.Tensor]:
    """Every leaf tensor requiring grad that ONE evaluation of the vector field reaches, other than the
    state itself: the term's parameters, a model captured by a closure (``ODETerm(lambda t, y: net(y))``),
    tensors inside ``args`` (the reference differentiates all of them through its eager loop).  Found by
    tracing f once under ``enable_grad`` and walking the autograd graph (one extra f evaluation)."""
    from .fields import BuiltinField

    if isinstance(term_.f, BuiltinField) and args is None:
        return []  # analytic fields: plain floats, nothing to differentiate but y0
    stats: Dict[str, Any] = {}
    term_.init(problem, stats)
    try:
        with torch.enable_grad():
            y = problem.y0.detach().clone().requires_grad_()
            out = term_.vf(problem.t_start, y, stats, args)
    except Exception:  # f cannot be evaluated here (e.g. a test double that must never be called): nothing to trace
        return []
    if not isinstance(out, torch.Tensor):
        return []
    leaves, seen, stack = [], set(), [out.grad_fn]
    while stack:
        node = stack.pop()
        if node is None or node in seen:
            continue
        seen.add(node)
        v = getattr(node, "variable", None)
        if v is not None and v is not y and v.requires_grad and all(v is not w for w in leaves):
            leaves.append(v)
        stack.extend(fn for fn, _ in node.next_functions)
    return leaves


def solve_with_grad(solver, problem: InitialValueProblem, term_, dt0, args,
                    leaves: Optional[List[torch.Tensor]] = None) -> Solution:
    """``AutoDiffAdjoint.solve`` for inputs / parameters that require gradients.  ``leaves``: what f
    depends on differentiably besides the state (``grad_leaves``)."""
    params = grad_leaves(term_, problem, args) if leaves is None else leaves
    ys, status = _ReplaySolve.apply(solver, term_, problem, dt0, args, len(params), problem.y0, *params)
    ts = problem.t_eval if problem.t_eval is not None else problem.t_end[:, None]
    return Solution(ts=ts, ys=ys, stats=solver._last_solution.stats, status=status)

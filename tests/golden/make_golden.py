#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the REAL reference.

Run in the build container only (the reference tree is not available on the GPU box):

    mkdir -p /tmp/refstub/torchtyping && printf 'class _M(type):\\n    def __getitem__(c, i):\\n        return c\\nclass TensorType(metaclass=_M):\\n    pass\\nis_float = object()\\n' > /tmp/refstub/torchtyping/__init__.py
    PYTHONPATH=/tmp/refstub:/root/reference:/root/repo python tests/golden/make_golden.py

(torchtyping is annotation-only in the reference; the 6-line stub above is all it needs.)
Every case stores the inputs, the reference's Solution, and -- for the lock-step
cases -- a per-iteration trace of what the reference's step method / controller saw and
produced, recorded by thin wrappers (no reference code is modified or copied).
The fixtures pin the oracle (tests/test_oracle_golden.py) and, through it and directly,
the CUDA path (tests/test_gpu_parity.py).
"""
import os
import sys

import numpy as np
import torch

import torchode as to  # the reference, from /root/reference

from torchode_b200.fields import LinearDecay, LotkaVolterra, VanDerPol  # plain torch modules

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)


def field_of(name, params):
    return {"linear": LinearDecay, "vdp": VanDerPol, "lv": LotkaVolterra}[name](*params)


class Recorder:
    """Wraps f, the step method and the controller of the reference to log one solve."""

    def __init__(self):
        self.f_calls = []  # (t, y, out)
        self.iters = []  # dict per iteration

    def wrap_f(self, f):
        def g(t, y):
            out = f(t, y)
            self.f_calls.append((t.clone(), y.clone(), out.clone()))
            return out
        return g


def make_controller(spec, term):
    kind = spec["kind"]
    kw = dict(term=term)
    for k in ("dt_min", "dt_max", "safety", "factor_min", "factor_max"):
        if k in spec:
            kw[k] = spec[k]
    if spec.get("norm") == "max":
        kw["norm"] = to.step_size_controllers.max_norm
    if kind == "integral":
        return to.IntegralController(atol=spec["atol"], rtol=spec["rtol"], **kw)
    return to.PIDController(atol=spec["atol"], rtol=spec["rtol"], pcoeff=spec["pcoeff"],
                            icoeff=spec["icoeff"], dcoeff=spec["dcoeff"], **kw)


def run_case(name, *, field, params, method, ctrl, y0, t_start=None, t_end=None, t_eval=None,
             dt0=None, max_steps=None, trace=False):
    rec = Recorder()
    f = field_of(field, params)
    term = to.ODETerm(rec.wrap_f(f) if trace else f)
    step_method = {"dopri5": to.Dopri5, "tsit5": to.Tsit5}[method](term=term)
    controller = make_controller(ctrl, term)
    iters = []
    if trace:
        orig_adapt = controller.adapt_step_size

        def adapt(t0, dt, y0_, step_result, state, stats):
            out = orig_adapt(t0, dt, y0_, step_result, state, stats)
            accept, dt_next, state_next, status = out
            r1 = getattr(state, "prev_error_ratio", None)
            r2 = getattr(state, "prev_prev_error_ratio", None)
            iters.append(dict(
                t0=t0.clone(), dt=dt.clone(), y0=y0_.clone(), y1=step_result.y.clone(),
                err=step_result.error_estimate.clone(), accept=accept.clone(),
                dt_next=dt_next.clone(), status=status.clone(),
                r1=None if r1 is None else r1.clone(), r2=None if r2 is None else r2.clone()))
            return out

        controller.adapt_step_size = adapt
        orig_step = step_method.step

        def step(term_, running, y0_, t0, dt, state, *, stats, args):
            out = orig_step(term_, running, y0_, t0, dt, state, stats=stats, args=args)
            iters_k.append(out[1].k.clone())
            iters_running.append(running.clone())
            return out

        iters_k, iters_running = [], []
        step_method.step = step
    solver = to.AutoDiffAdjoint(step_method, controller, max_steps=max_steps)
    problem = to.InitialValueProblem(y0=y0, t_start=t_start, t_end=t_end, t_eval=t_eval)
    with torch.no_grad():
        sol = solver.solve(problem, dt0=dt0)
    out = dict(
        field=np.array(field), params=np.array(params, dtype=np.float64), method=np.array(method),
        ctrl_keys=np.array(sorted(ctrl.keys())),
        ctrl_vals=np.array([str(ctrl[k]) for k in sorted(ctrl.keys())]),
        y0=y0.numpy(), t_start=problem.t_start.numpy(), t_end=problem.t_end.numpy(),
        ys=sol.ys.numpy(), ts=sol.ts.numpy(), status=sol.status.numpy(),
        n_steps=sol.stats["n_steps"].numpy(), n_accepted=sol.stats["n_accepted"].numpy(),
        n_f_evals=sol.stats["n_f_evals"].numpy(), n_initialized=sol.stats["n_initialized"].numpy(),
        max_steps=np.array(-1 if max_steps is None else max_steps),
    )
    if t_eval is not None:
        out["t_eval"] = t_eval.numpy()
    if dt0 is not None:
        out["dt0"] = dt0.numpy()
    if trace:
        n = len(iters)
        out["trace_n"] = np.array(n)
        for key in ("t0", "dt", "y0", "y1", "err", "accept", "dt_next", "status"):
            out["trace_" + key] = np.stack([it[key].numpy() for it in iters])
        if iters[0]["r1"] is not None:
            out["trace_r1"] = np.stack([it["r1"].numpy() for it in iters])
            out["trace_r2"] = np.stack([it["r2"].numpy() for it in iters])
        out["trace_k"] = np.stack([k.numpy() for k in iters_k])  # (n, S, B, F)
        out["trace_running"] = np.stack([r.numpy() for r in iters_running])
        # the stage inputs the reference handed to f: first 2 calls are the initial-step
        # heuristic (if dt0 is None), then 6 per iteration
        n_init = len(rec.f_calls) - 6 * n
        out["trace_n_init_calls"] = np.array(n_init)
        out["trace_init_t"] = np.stack([c[0].numpy() for c in rec.f_calls[:n_init]])
        out["trace_init_y"] = np.stack([c[1].numpy() for c in rec.f_calls[:n_init]])
        out["trace_init_f"] = np.stack([c[2].numpy() for c in rec.f_calls[:n_init]])
        st_t = np.stack([c[0].numpy() for c in rec.f_calls[n_init:]])
        st_y = np.stack([c[1].numpy() for c in rec.f_calls[n_init:]])
        out["trace_stage_t"] = st_t.reshape(n, 6, *st_t.shape[1:])
        out["trace_stage_y"] = st_y.reshape(n, 6, *st_y.shape[1:])
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: iters={int(sol.stats['n_f_evals'][0])} n_steps={sol.stats['n_steps'][:6].tolist()} "
          f"n_acc={sol.stats['n_accepted'][:6].tolist()} status={sol.status[:6].tolist()} "
          f"-> {os.path.getsize(path) / 1024:.1f} KiB")


def main():
    g = torch.Generator().manual_seed(1234)
    # reference tableaux (float64) -- pins torchode_b200/tableaus.py incl. the sympy-derived Tsit5 weights
    tabs = {}
    for nm, cls in (("dopri5", to.Dopri5), ("tsit5", to.Tsit5)):
        tb = cls.TABLEAU
        for k in ("c", "a", "b", "b_err", "b_other"):
            tabs[f"{nm}_{k}"] = getattr(tb, k).numpy()
        tabs[f"{nm}_fsal"] = np.array(tb.fsal)
        tabs[f"{nm}_ssal"] = np.array(tb.ssal)
    np.savez_compressed(os.path.join(HERE, "tableaus.npz"), **tabs)

    I63 = dict(kind="integral", atol=1e-6, rtol=1e-3)
    PID8 = dict(kind="pid", atol=1e-8, rtol=1e-8, pcoeff=0.2, icoeff=0.5, dcoeff=0.0)

    # C1: README example (README.md:36-59)
    y0 = torch.tensor([[1.2], [5.0]])
    t_eval = torch.stack((torch.linspace(0, 5, 10), torch.linspace(3, 4, 10)))
    for m in ("dopri5", "tsit5"):
        run_case(f"c1_readme_{m}", field="linear", params=[-0.5], method=m, ctrl=I63, y0=y0,
                 t_eval=t_eval, trace=True)

    # Appendix B.2 + 60 seeded samples: Lotka-Volterra fp32, Dopri5 + I(1e-6,1e-3), 100 t_eval
    y0 = torch.cat((torch.tensor([[1., 1.], [1.5, .5], [2., 2.], [.5, 1.75]]),
                    1 + torch.rand(60, 2, generator=g)))
    t_eval = torch.linspace(0, 10, 100).repeat(64, 1)
    run_case("c3_lv_f32_dopri5", field="lv", params=[1.5, 1.0, 1.0, 3.0], method="dopri5", ctrl=I63,
             y0=y0, t_eval=t_eval, trace=True)
    run_case("lv_f32_tsit5_pid", field="lv", params=[1.5, 1.0, 1.0, 3.0], method="tsit5",
             ctrl=dict(kind="pid", atol=1e-6, rtol=1e-5, pcoeff=0.2, icoeff=0.5, dcoeff=0.0),
             y0=y0[:16], t_eval=t_eval[:16], trace=True)
    run_case("lv_f64_dopri5_tol8", field="lv", params=[1.5, 1.0, 1.0, 3.0], method="dopri5",
             ctrl=dict(kind="integral", atol=1e-8, rtol=1e-8), y0=y0[:16].double(),
             t_eval=t_eval[:16].double(), trace=False)

    # Appendix B.3 + 28 seeded samples: Van der Pol mu=10 fp64, Tsit5 + PID(1e-8), t in [0,20], no t_eval
    y0 = torch.cat((torch.tensor([[2., 0.], [-1., 1.], [.5, -.5], [.1, 0.]], dtype=torch.float64),
                    torch.rand(28, 2, generator=g, dtype=torch.float64) * 4 - 2))
    B = y0.shape[0]
    run_case("c2_vdp_f64_tsit5_pid", field="vdp", params=[10.0], method="tsit5", ctrl=PID8, y0=y0,
             t_start=torch.zeros(B, dtype=torch.float64),
             t_end=torch.full((B,), 20.0, dtype=torch.float64), trace=False)
    run_case("vdp_f64_tsit5_pid_short_trace", field="vdp", params=[10.0], method="tsit5", ctrl=PID8,
             y0=y0[:8], t_start=torch.zeros(8, dtype=torch.float64),
             t_end=torch.full((8,), 1.0, dtype=torch.float64), trace=True)
    run_case("vdp_f64_dopri5_pidd", field="vdp", params=[2.0], method="dopri5",
             ctrl=dict(kind="pid", atol=1e-7, rtol=1e-6, pcoeff=0.3, icoeff=0.4, dcoeff=0.1),
             y0=y0[:8], t_start=torch.zeros(8, dtype=torch.float64),
             t_end=torch.full((8,), 5.0, dtype=torch.float64), trace=True)

    # opposite time directions in one batch, ragged spans, t_eval rows per sample (adjoint_test.py:350-368)
    y0 = torch.tensor([[1.0, 2.0], [0.5, 0.25], [3.0, -1.0]])
    t_eval = torch.stack((torch.linspace(0, 2, 7), torch.linspace(2, -1, 7), torch.linspace(1, 1.5, 7)))
    for nm, ctrl in (("integral", I63), ("pid", dict(kind="pid", atol=1e-6, rtol=1e-4, pcoeff=0.2,
                                                      icoeff=0.5, dcoeff=0.0))):
        run_case(f"linear_bidir_{nm}", field="linear", params=[-0.7], method="dopri5", ctrl=ctrl, y0=y0,
                 t_eval=t_eval, trace=True)

    # mixed dtypes (dtype_stability_test.py:15-42)
    y0 = 1 + torch.rand(8, 2, generator=g)
    te = torch.linspace(0, 3, 11).repeat(8, 1)
    run_case("lv_data32_time64", field="lv", params=[1.5, 1.0, 1.0, 3.0], method="tsit5", ctrl=I63,
             y0=y0, t_eval=te.double(), trace=True)
    run_case("lv_data64_time32", field="lv", params=[1.5, 1.0, 1.0, 3.0], method="dopri5", ctrl=I63,
             y0=y0.double(), t_eval=te, trace=True)

    # user dt0, max_steps, dt_min / dt_max, max_norm, F = 3 / 4
    run_case("lv_dt0_maxsteps", field="lv", params=[1.5, 1.0, 1.0, 3.0], method="dopri5", ctrl=I63, y0=y0,
             t_eval=te, dt0=torch.full((8,), 0.05), max_steps=9, trace=True)
    run_case("lv_dtmin", field="lv", params=[1.5, 1.0, 1.0, 3.0], method="dopri5",
             ctrl=dict(kind="integral", atol=1e-9, rtol=1e-9, dt_min=0.02), y0=y0, t_eval=te, trace=True)
    run_case("lv_dtmax_maxnorm", field="lv", params=[1.5, 1.0, 1.0, 3.0], method="tsit5",
             ctrl=dict(kind="pid", atol=1e-4, rtol=1e-3, pcoeff=0.1, icoeff=0.6, dcoeff=0.05, dt_max=0.11,
                       norm="max", safety=0.8, factor_min=0.3, factor_max=4.0),
             y0=y0, t_eval=te, trace=True)
    y0 = torch.randn(6, 3, generator=g)
    run_case("linear_f3_no_teval", field="linear", params=[-1.3], method="tsit5", ctrl=I63, y0=y0,
             t_start=torch.zeros(6), t_end=torch.linspace(0.0, 4.0, 6), trace=True)
    y0 = torch.randn(5, 4, generator=g, dtype=torch.float64)
    run_case("linear_f4_f64", field="linear", params=[0.4], method="dopri5", ctrl=PID8, y0=y0,
             t_eval=torch.linspace(0, 2, 5, dtype=torch.float64).repeat(5, 1), trace=True)
    # non-finite derivative -> INFINITE_NORM stops the whole batch (adjoint_test.py:215-239)
    y0 = torch.tensor([[1.0, 1.0], [float("inf"), 1.0], [2.0, 0.5]])
    run_case("lv_infinite_norm", field="lv", params=[1.5, 1.0, 1.0, 3.0], method="dopri5", ctrl=I63, y0=y0,
             t_eval=torch.linspace(0, 1, 4).repeat(3, 1), trace=False)


if __name__ == "__main__":
    sys.exit(main())

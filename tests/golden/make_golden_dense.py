#!/usr/bin/env python
"""Golden fixtures with mid-size feature dimensions (F = 16, 100) from the REAL reference: a dense linear
field f(t, y) = y A^T, so that the per-sample norm reduction over F > 4 features (lane groups of the
canonical order, odd F) is pinned on the reference too.  Run in the build container only (same stub as
make_golden.py):

    PYTHONPATH=/tmp/refstub:/root/reference:/root/repo python tests/golden/make_golden_dense.py
"""
import os

import numpy as np
import torch

import torchode as to  # the reference, from /root/reference

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)


def run(name, B, F, dtype, method, ctrl_kind, seed, n_eval, reverse_half):
    g = torch.Generator().manual_seed(seed)
    Q = torch.randn(F, F, generator=g, dtype=torch.float64)
    A = (-(Q @ Q.T) / F - 0.3 * torch.eye(F, dtype=torch.float64) + 0.5 * (Q - Q.T) / F ** 0.5).to(dtype)  # damped rotation
    y0 = torch.randn(B, F, generator=g, dtype=dtype)
    t_start = torch.zeros(B, dtype=dtype)
    t_end = 1.0 + torch.rand(B, generator=g, dtype=dtype)
    if reverse_half:
        t_start, t_end = torch.where(torch.arange(B) % 2 == 0, t_start, t_end), torch.where(torch.arange(B) % 2 == 0, t_end, t_start)
    t_eval = None
    if n_eval:
        frac = torch.sort(torch.rand(B, n_eval, generator=g, dtype=dtype), dim=1).values
        frac[::2, 0] = 0.0
        t_eval = t_start[:, None] + (t_end - t_start)[:, None] * frac
    term = to.ODETerm(lambda t, y: y @ A.T)
    step = {"dopri5": to.Dopri5, "tsit5": to.Tsit5}[method](term=term)
    if ctrl_kind == "pid":
        ctrl = to.PIDController(atol=1e-8, rtol=1e-6, pcoeff=0.2, icoeff=0.5, dcoeff=0.0, term=term)
    else:
        ctrl = to.IntegralController(atol=1e-7 if dtype == torch.float32 else 1e-9, rtol=1e-4 if dtype == torch.float32 else 1e-7,
                                     term=term)
    solver = to.AutoDiffAdjoint(step, ctrl)
    with torch.no_grad():
        sol = solver.solve(to.InitialValueProblem(y0=y0, t_start=t_start, t_end=t_end, t_eval=t_eval))
    out = dict(A=A.numpy(), y0=y0.numpy(), t_start=t_start.numpy(), t_end=t_end.numpy(), method=np.array(method),
               ctrl=np.array(ctrl_kind), ys=sol.ys.numpy(), status=sol.status.numpy(),
               n_steps=sol.stats["n_steps"].numpy(), n_accepted=sol.stats["n_accepted"].numpy(),
               n_f_evals=sol.stats["n_f_evals"].numpy(), n_initialized=sol.stats["n_initialized"].numpy())
    if t_eval is not None:
        out["t_eval"] = t_eval.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n_steps", out["n_steps"].tolist(), "n_accepted", out["n_accepted"].tolist(),
          "n_f_evals", int(out["n_f_evals"][0]), "status", out["status"].tolist())


if __name__ == "__main__":
    run("dense_f16_f64_dopri5_teval", 5, 16, torch.float64, "dopri5", "integral", 11, 7, False)
    run("dense_f100_f64_tsit5_pid_bidir", 4, 100, torch.float64, "tsit5", "pid", 12, 0, True)
    # (an fp32 case with F = 33 was tried and dropped: one accept decision of the reference flips under the
    # rounding-level difference between its matmul and numpy's -- the fp32 noise floor of SURVEY.md Appendix C)

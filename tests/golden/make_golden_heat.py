#!/usr/bin/env python
"""Golden fixtures of BASELINE.json configs[4] in miniature (method-of-lines heat equation) from the REAL
reference.  Run in the build container only (same stub as make_golden.py):

    PYTHONPATH=/tmp/refstub:/root/reference:/root/repo python tests/golden/make_golden_heat.py

f is the plain PyTorch expression of fields.Heat1D (forward_reference) on the CPU.  Stored: inputs, the
reference's Solution (ys, n_steps, n_accepted, n_f_evals, n_initialized, status).  The fixtures pin the
oracle (tests/test_oracle_golden.py) and the step-fused / stage-wise CUDA routes (tests/test_gpu_parity.py).
"""
import os

import numpy as np
import torch

import torchode as to  # the reference, from /root/reference

from torchode_b200.fields import Heat1D  # plain torch module

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)
KAPPA = 20.0  # stencil without the 1/dx^2 factor: spectral radius 4 kappa = 80, non-stiff


def problem(B, N, dtype, seed, with_t_eval):
    g = torch.Generator().manual_seed(seed)
    x = torch.linspace(0, 1, N, dtype=dtype)
    amp = torch.rand(B, 3, generator=g, dtype=dtype)
    y0 = sum(amp[:, k - 1:k] * torch.sin(k * torch.pi * x)[None] for k in (1, 2, 3))
    y0 = y0 + 0.01 * torch.randn(B, N, generator=g, dtype=dtype)
    t_start = torch.zeros(B, dtype=dtype)
    t_end = 0.3 + 0.4 * torch.rand(B, generator=g, dtype=dtype)
    t_eval = None
    if with_t_eval:
        frac = torch.sort(torch.rand(B, 9, generator=g, dtype=dtype), dim=1).values
        frac[::2, 0] = 0.0
        t_eval = t_end[:, None] * frac
    return y0, t_start, t_end, t_eval


def run(name, B, N, dtype, method, seed, with_t_eval):
    y0, t_start, t_end, t_eval = problem(B, N, dtype, seed, with_t_eval)
    field = Heat1D(KAPPA)
    term = to.ODETerm(field.forward_reference)
    step = {"dopri5": to.Dopri5, "tsit5": to.Tsit5}[method](term=term)
    ctrl = to.IntegralController(atol=1e-6, rtol=1e-3, term=term)
    solver = to.AutoDiffAdjoint(step, ctrl)
    with torch.no_grad():
        sol = solver.solve(to.InitialValueProblem(y0=y0, t_start=t_start, t_end=t_end, t_eval=t_eval))
    out = dict(y0=y0.numpy(), t_start=t_start.numpy(), t_end=t_end.numpy(), kappa=np.float64(KAPPA),
               method=np.array(method), ys=sol.ys.numpy(), status=sol.status.numpy(),
               n_steps=sol.stats["n_steps"].numpy(), n_accepted=sol.stats["n_accepted"].numpy(),
               n_f_evals=sol.stats["n_f_evals"].numpy(), n_initialized=sol.stats["n_initialized"].numpy())
    if t_eval is not None:
        out["t_eval"] = t_eval.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "n_steps", out["n_steps"].tolist(), "n_accepted", out["n_accepted"].tolist(),
          "n_f_evals", int(out["n_f_evals"][0]), "status", out["status"].tolist())


if __name__ == "__main__":
    run("heat_f64_tsit5", 4, 1024, torch.float64, "tsit5", 1, False)
    run("heat_f64_dopri5_teval", 3, 2052, torch.float64, "dopri5", 2, True)
    run("heat_f32_tsit5", 4, 4100, torch.float32, "tsit5", 3, False)
    run("heat_f32_tsit5_teval", 3, 1024, torch.float32, "tsit5", 4, True)

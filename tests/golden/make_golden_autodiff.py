#!/usr/bin/env python
"""Golden gradients of the REAL reference's AutoDiffAdjoint (back-propagation through its eager
loop, CPU fp64), with and without back-propagation through the step-size control.

    PYTHONPATH=/tmp/refstub:/root/reference:/root/repo python tests/golden/make_golden_autodiff.py
"""
import os

import numpy as np
import torch

import torchode as to  # the reference

from make_golden_backsolve import TanhField

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    B, n = 6, 3
    out = {}
    for method_name, method_cls in (("dopri5", to.Dopri5), ("tsit5", to.Tsit5)):
        for through in (True, False):
            for with_t_eval in (False, True):
                for ctrl_name in ("integral", "pid"):
                    model = TanhField(n, 8)
                    y0 = torch.randn(B, n, generator=torch.Generator().manual_seed(11), dtype=torch.float64).requires_grad_()
                    w = torch.randn(B, n, generator=torch.Generator().manual_seed(12), dtype=torch.float64)
                    term = to.ODETerm(model)
                    if ctrl_name == "integral":
                        ctrl = to.IntegralController(1e-6, 1e-5, term=term)
                    else:
                        ctrl = to.PIDController(1e-6, 1e-5, 0.2, 0.5, 0.1, term=term)
                    solver = to.AutoDiffAdjoint(method_cls(term), ctrl, backprop_through_step_size_control=through)
                    if with_t_eval:
                        t_eval = torch.linspace(0.0, 2.0, 5, dtype=torch.float64).repeat(B, 1)
                        sol = solver.solve(to.InitialValueProblem(y0=y0, t_eval=t_eval))
                        loss = (sol.ys[:, -1] * w).sum() + (sol.ys[:, 2] ** 2).sum() + (sol.ys[:, 0] * w).sum()
                    else:
                        sol = solver.solve(to.InitialValueProblem(y0=y0, t_start=torch.zeros(B, dtype=torch.float64),
                                                               t_end=torch.full((B,), 2.0, dtype=torch.float64)))
                        loss = (sol.ys[:, -1] * w).sum()
                    grads = torch.autograd.grad(loss, [y0] + list(model.parameters()))
                    key = f"{method_name}_{ctrl_name}_{'through' if through else 'detached'}_{'teval' if with_t_eval else 'tend'}"
                    out[f"{key}_ys"] = sol.ys.detach().numpy()
                    out[f"{key}_n_steps"] = sol.stats["n_steps"].numpy()
                    out[f"{key}_grad_y0"] = grads[0].numpy()
                    for i, gp in enumerate(grads[1:]):
                        out[f"{key}_grad_p{i}"] = gp.numpy()
                    print(key, float(loss.detach()), sol.stats["n_steps"].tolist(), float(grads[0].abs().max()))
    out["y0"] = y0.detach().numpy()
    out["w"] = w.numpy()
    for i, p in enumerate(TanhField(n, 8).parameters()):
        out[f"param{i}"] = p.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "autodiff_gradients.npz"), **out)


if __name__ == "__main__":
    main()

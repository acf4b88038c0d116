#!/usr/bin/env python
"""Golden gradients of the REAL reference's BacksolveAdjoint / JointBacksolveAdjoint (CPU, fp64).

    PYTHONPATH=/tmp/refstub:/root/reference:/root/repo python tests/golden/make_golden_backsolve.py

(see make_golden.py for the torchtyping stub).  Stored: inputs, the module's parameters, ys and
the gradients of a fixed scalar loss w.r.t. y0 and every parameter."""
import os

import numpy as np
import torch

import torchode as to  # the reference

HERE = os.path.dirname(os.path.abspath(__file__))


class TanhField(torch.nn.Module):
    def __init__(self, n, hidden):
        super().__init__()
        torch.manual_seed(3)
        self.l1 = torch.nn.Linear(n, hidden).double()
        self.l2 = torch.nn.Linear(hidden, n).double()

    def forward(self, t, y):
        return self.l2(torch.tanh(self.l1(y))) * (1 + 0.1 * torch.sin(t)[..., None])


def main():
    g = torch.Generator().manual_seed(11)
    B, n = 6, 3
    out = {}
    for name, cls in (("backsolve", to.BacksolveAdjoint), ("joint", to.JointBacksolveAdjoint)):
        for with_t_eval in (False, True):
            model = TanhField(n, 8)
            y0 = torch.randn(B, n, generator=torch.Generator().manual_seed(11), dtype=torch.float64).requires_grad_()
            w = torch.randn(B, n, generator=torch.Generator().manual_seed(12), dtype=torch.float64)
            term = to.ODETerm(model)
            adj = cls(term, to.Tsit5(term), to.IntegralController(1e-9, 1e-9, term=term))
            if with_t_eval:
                t_eval = torch.linspace(0.0, 2.0, 5, dtype=torch.float64).repeat(B, 1)
                sol = adj.solve(to.InitialValueProblem(y0=y0, t_eval=t_eval))
                loss = (sol.ys[:, -1] * w).sum() + (sol.ys[:, 2] ** 2).sum()
            else:
                sol = adj.solve(to.InitialValueProblem(y0=y0, t_start=torch.zeros(B, dtype=torch.float64),
                                                       t_end=torch.full((B,), 2.0, dtype=torch.float64)))
                loss = (sol.ys[:, -1] * w).sum()
            params = list(model.parameters())
            grads = torch.autograd.grad(loss, [y0] + params)
            key = f"{name}_{'teval' if with_t_eval else 'tend'}"
            out[f"{key}_ys"] = sol.ys.detach().numpy()
            out[f"{key}_loss"] = loss.detach().numpy()
            out[f"{key}_grad_y0"] = grads[0].numpy()
            for i, gp in enumerate(grads[1:]):
                out[f"{key}_grad_p{i}"] = gp.numpy()
            print(key, float(loss), [float(x.abs().max()) for x in grads])
    out["y0"] = y0.detach().numpy()
    out["w"] = w.numpy()
    for i, p in enumerate(TanhField(n, 8).parameters()):
        out[f"param{i}"] = p.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "backsolve_gradients.npz"), **out)


if __name__ == "__main__":
    main()

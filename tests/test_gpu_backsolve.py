"""Continuous-adjoint gradients (BacksolveAdjoint / JointBacksolveAdjoint, adjoints.py:343-681)
on top of the CUDA solve loop, checked against closed-form gradients of a linear ODE."""
import pytest
import torch

import torchode_b200 as to

pytestmark = pytest.mark.gpu
DEV = "cuda"


class Linear(torch.nn.Module):
    """y' = A y + b with trainable A, b."""

    def __init__(self, n):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.A = torch.nn.Parameter(0.5 * torch.randn(n, n, generator=g, dtype=torch.float64))
        self.b = torch.nn.Parameter(0.1 * torch.randn(n, generator=g, dtype=torch.float64))

    def forward(self, t, y):
        return y @ self.A.T + self.b


def closed_form(model, y0, t):
    """y(t) = e^{At} y0 + A^{-1}(e^{At} - I) b, differentiable through torch.matrix_exp."""
    n = y0.shape[1]
    out = []
    for i in range(y0.shape[0]):
        E = torch.matrix_exp(model.A * t[i])
        out.append(E @ y0[i] + torch.linalg.solve(model.A, (E - torch.eye(n, dtype=E.dtype, device=E.device)) @ model.b))
    return torch.stack(out)


@pytest.mark.parametrize("adjoint_cls", [to.BacksolveAdjoint, to.JointBacksolveAdjoint])
@pytest.mark.parametrize("with_t_eval", [False, True])
def test_gradients_match_the_closed_form(adjoint_cls, with_t_eval):
    B, n = 5, 3
    model = Linear(n).to(DEV)
    g = torch.Generator().manual_seed(1)
    y0 = torch.randn(B, n, generator=g, dtype=torch.float64).to(DEV).requires_grad_()
    t0 = torch.zeros(B, dtype=torch.float64, device=DEV)
    t1 = torch.full((B,), 1.5, dtype=torch.float64, device=DEV)
    term = to.ODETerm(model)
    adjoint = adjoint_cls(term, to.Tsit5(term), to.IntegralController(1e-10, 1e-10, term=term))
    w = torch.randn(B, n, generator=g, dtype=torch.float64).to(DEV)
    if with_t_eval:
        t_eval = torch.linspace(0, 1.5, 4, dtype=torch.float64, device=DEV).repeat(B, 1)
        sol = adjoint.solve(to.InitialValueProblem(y0, t_eval=t_eval))
        assert sol.ys.shape == (B, 4, n)
        loss = (sol.ys[:, -1] * w).sum() + (sol.ys[:, 2] * w).sum()
        want = (closed_form(model, y0, t1) * w).sum() + (closed_form(model, y0, t_eval[:, 2]) * w).sum()
    else:
        sol = adjoint.solve(to.InitialValueProblem(y0, t0, t1))
        loss = (sol.ys[:, -1] * w).sum()
        want = (closed_form(model, y0, t1) * w).sum()
    assert (sol.status == 0).all()
    assert torch.allclose(loss, want, rtol=1e-8)
    got = torch.autograd.grad(loss, [y0, model.A, model.b])
    ref = torch.autograd.grad(want, [y0, model.A, model.b])
    for a, b in zip(got, ref):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)
    assert len(sol.stats["backsolve"]) == (3 if with_t_eval else 1)


def test_backsolve_forward_uses_the_fused_kernel_for_builtin_fields():
    B = 64
    y0 = (1 + torch.rand(B, 2, dtype=torch.float64, device=DEV)).requires_grad_()
    term = to.ODETerm(to.fields.LotkaVolterra())
    adjoint = to.BacksolveAdjoint(term, to.Dopri5(term), to.IntegralController(1e-9, 1e-9, term=term))
    t0 = torch.zeros(B, dtype=torch.float64, device=DEV)
    t1 = torch.full((B,), 2.0, dtype=torch.float64, device=DEV)
    sol = adjoint.solve(to.InitialValueProblem(y0, t0, t1))
    assert adjoint.forward_adjoint.last_run["route"] == "fused"
    (gy,) = torch.autograd.grad(sol.ys[:, -1].sum(), [y0])
    # finite-difference check of d sum(y(T)) / d y0 on one coordinate
    eps = 1e-6
    with torch.no_grad():
        yp = y0.detach().clone()
        yp[:, 0] += eps
        plain = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-9, 1e-9, term=term))
        a = plain.solve(to.InitialValueProblem(yp, t0, t1)).ys[:, -1].sum(dim=1)
        b = plain.solve(to.InitialValueProblem(y0.detach(), t0, t1)).ys[:, -1].sum(dim=1)
    assert torch.allclose((a - b) / eps, gy[:, 0], rtol=1e-3, atol=1e-5)


class TanhField(torch.nn.Module):
    """Same module as tests/golden/make_golden_backsolve.py (parameters loaded from the fixture)."""

    def __init__(self, n, hidden):
        super().__init__()
        self.l1 = torch.nn.Linear(n, hidden).double()
        self.l2 = torch.nn.Linear(hidden, n).double()

    def forward(self, t, y):
        return self.l2(torch.tanh(self.l1(y))) * (1 + 0.1 * torch.sin(t)[..., None])


@pytest.mark.parametrize("name,cls", [("backsolve", to.BacksolveAdjoint), ("joint", to.JointBacksolveAdjoint)])
@pytest.mark.parametrize("with_t_eval", [False, True])
def test_gradients_match_the_reference_golden(name, cls, with_t_eval):
    """Gradients of the REAL reference's adjoints (CPU fp64, tests/golden/backsolve_gradients.npz)."""
    import os

    import numpy as np

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "backsolve_gradients.npz"))
    B, n = z["y0"].shape
    model = TanhField(n, 8)
    with torch.no_grad():
        for i, p in enumerate(model.parameters()):
            p.copy_(torch.from_numpy(z[f"param{i}"]))
    model = model.to(DEV)
    y0 = torch.from_numpy(z["y0"]).to(DEV).requires_grad_()
    w = torch.from_numpy(z["w"]).to(DEV)
    term = to.ODETerm(model)
    adj = cls(term, to.Tsit5(term), to.IntegralController(1e-9, 1e-9, term=term))
    if with_t_eval:
        t_eval = torch.linspace(0.0, 2.0, 5, dtype=torch.float64, device=DEV).repeat(B, 1)
        sol = adj.solve(to.InitialValueProblem(y0, t_eval=t_eval))
        loss = (sol.ys[:, -1] * w).sum() + (sol.ys[:, 2] ** 2).sum()
    else:
        sol = adj.solve(to.InitialValueProblem(y0, torch.zeros(B, dtype=torch.float64, device=DEV),
                                               torch.full((B,), 2.0, dtype=torch.float64, device=DEV)))
        loss = (sol.ys[:, -1] * w).sum()
    key = f"{name}_{'teval' if with_t_eval else 'tend'}"
    assert np.allclose(sol.ys.detach().cpu().numpy(), z[f"{key}_ys"], rtol=1e-7, atol=1e-9)
    assert abs(float(loss) - float(z[f"{key}_loss"])) <= 1e-7 * abs(float(z[f"{key}_loss"]))
    grads = torch.autograd.grad(loss, [y0] + list(model.parameters()))
    assert np.allclose(grads[0].cpu().numpy(), z[f"{key}_grad_y0"], rtol=1e-6, atol=1e-8)
    for i, gp in enumerate(grads[1:]):
        assert np.allclose(gp.cpu().numpy(), z[f"{key}_grad_p{i}"], rtol=1e-6, atol=1e-8), f"param {i}"

"""Gradients through AutoDiffAdjoint.solve (recompute-based backward over the recorded CUDA loop)
against the REAL reference's autograd-through-the-eager-loop gradients (CPU fp64 goldens,
tests/golden/autodiff_gradients.npz), with and without back-propagation through the step-size
control, for both methods, both controllers, with and without t_eval."""
import os

import numpy as np
import pytest
import torch

import torchode_b200 as to

from test_gpu_backsolve import TanhField

pytestmark = pytest.mark.gpu
DEV = "cuda"
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "autodiff_gradients.npz"))


def setup(method_name, ctrl_name, through):
    B, n = Z["y0"].shape
    model = TanhField(n, 8)
    with torch.no_grad():
        for i, p in enumerate(model.parameters()):
            p.copy_(torch.from_numpy(Z[f"param{i}"]))
    model = model.to(DEV)
    term = to.ODETerm(model)
    if ctrl_name == "integral":
        ctrl = to.IntegralController(1e-6, 1e-5, term=term)
    else:
        ctrl = to.PIDController(1e-6, 1e-5, 0.2, 0.5, 0.1, term=term)
    method = {"dopri5": to.Dopri5, "tsit5": to.Tsit5}[method_name](term)
    solver = to.AutoDiffAdjoint(method, ctrl, backprop_through_step_size_control=through)
    y0 = torch.from_numpy(Z["y0"]).to(DEV).requires_grad_()
    w = torch.from_numpy(Z["w"]).to(DEV)
    return model, solver, y0, w


@pytest.mark.parametrize("method_name", ["dopri5", "tsit5"])
@pytest.mark.parametrize("ctrl_name", ["integral", "pid"])
@pytest.mark.parametrize("through", [True, False])
@pytest.mark.parametrize("with_t_eval", [False, True])
def test_gradients_match_the_reference(method_name, ctrl_name, through, with_t_eval):
    model, solver, y0, w = setup(method_name, ctrl_name, through)
    B = y0.shape[0]
    if with_t_eval:
        t_eval = torch.linspace(0.0, 2.0, 5, dtype=torch.float64, device=DEV).repeat(B, 1)
        sol = solver.solve(to.InitialValueProblem(y0, t_eval=t_eval))
        loss = (sol.ys[:, -1] * w).sum() + (sol.ys[:, 2] ** 2).sum() + (sol.ys[:, 0] * w).sum()
    else:
        sol = solver.solve(to.InitialValueProblem(y0, torch.zeros(B, dtype=torch.float64, device=DEV),
                                                  torch.full((B,), 2.0, dtype=torch.float64, device=DEV)))
        loss = (sol.ys[:, -1] * w).sum()
    key = f"{method_name}_{ctrl_name}_{'through' if through else 'detached'}_{'teval' if with_t_eval else 'tend'}"
    assert sol.stats["n_steps"].cpu().tolist() == Z[f"{key}_n_steps"].tolist()
    assert np.allclose(sol.ys.detach().cpu().numpy(), Z[f"{key}_ys"], rtol=1e-10, atol=1e-12)
    grads = torch.autograd.grad(loss, [y0] + list(model.parameters()))
    # the two step-size-control modes differ at the 1e-6 level: 1e-8 tells them apart
    assert np.allclose(grads[0].cpu().numpy(), Z[f"{key}_grad_y0"], rtol=1e-8, atol=1e-11)
    for i, gp in enumerate(grads[1:]):
        assert np.allclose(gp.cpu().numpy(), Z[f"{key}_grad_p{i}"], rtol=1e-8, atol=1e-11), f"param {i}"


def test_training_step_on_a_neural_ode_runs_and_reduces_the_loss():
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(2, 16), torch.nn.Tanh(), torch.nn.Linear(16, 2)).to(DEV)

    class F(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = model

        def forward(self, t, y):
            return self.net(y)

    term = to.ODETerm(F())
    solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-5, 1e-4, term=term))
    y0 = torch.randn(64, 2, device=DEV)
    target = torch.zeros(64, 2, device=DEV)
    t0, t1 = torch.zeros(64, device=DEV), torch.ones(64, device=DEV)
    opt = torch.optim.SGD(term.parameters(), lr=0.05)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        sol = solver.solve(to.InitialValueProblem(y0, t0, t1))
        loss = ((sol.ys[:, -1] - target) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(torch.isfinite(p.grad).all() for p in term.parameters())
    assert losses[-1] < losses[0]

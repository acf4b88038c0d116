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


# ---- round 2 (ADVICE): leaves reached through closures / args, refused configurations, kernel fields ----
def _module_and_closure_solvers():
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Linear(3, 8), torch.nn.Tanh(), torch.nn.Linear(8, 3)).double().to(DEV)

    class F(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = net

        def forward(self, t, y):
            return self.net(y)

    def build(f):
        term = to.ODETerm(f)
        return to.AutoDiffAdjoint(to.Tsit5(term), to.IntegralController(1e-7, 1e-6, term=term))

    return net, build(F()), build(lambda t, y: net(y))


@pytest.mark.parametrize("y0_requires_grad", [True, False])
def test_parameters_captured_by_a_closure_get_the_module_gradients(y0_requires_grad):
    """ODETerm(lambda t, y: net(y)): term.parameters() is empty; the leaves are found by tracing f once."""
    net, as_module, as_closure = _module_and_closure_solvers()
    g = torch.Generator().manual_seed(1)
    y0 = torch.randn(7, 3, generator=g, dtype=torch.float64).to(DEV).requires_grad_(y0_requires_grad)
    t0, t1 = torch.zeros(7, dtype=torch.float64, device=DEV), torch.full((7,), 1.5, dtype=torch.float64, device=DEV)
    grads = []
    for solver in (as_module, as_closure):
        sol = solver.solve(to.InitialValueProblem(y0, t0, t1))
        assert sol.ys.requires_grad
        leaves = list(net.parameters()) + ([y0] if y0_requires_grad else [])
        grads.append(torch.autograd.grad((sol.ys ** 2).sum(), leaves))
    for a, b in zip(*grads):
        assert float(a.abs().max()) > 0 and torch.equal(a, b)


def test_tensors_inside_args_get_gradients():
    """docs/extra-args.md of the reference: f(t, y, args) with a tensor in args that requires grad."""
    rate = torch.tensor([-0.3, -0.9], dtype=torch.float64, device=DEV, requires_grad=True)
    term = to.ODETerm(lambda t, y, a: a[None, :] * y, with_args=True)
    solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-9, 1e-9, term=term))
    y0 = torch.tensor([[1.0, 2.0], [0.5, -1.0]], dtype=torch.float64, device=DEV)
    T = 1.25
    sol = solver.solve(to.InitialValueProblem(y0, torch.zeros(2, dtype=torch.float64, device=DEV),
                                              torch.full((2,), T, dtype=torch.float64, device=DEV)), args=rate)
    (g,) = torch.autograd.grad(sol.ys[:, -1].sum(), [rate])
    want = (T * y0 * torch.exp(rate.detach() * T)[None, :]).sum(dim=0)  # d/drate of y0 exp(rate T)
    assert torch.allclose(g, want, rtol=1e-6)


@pytest.mark.parametrize("config", ["heun", "fixed_step", "euler"])
def test_configurations_without_gradient_support_refuse_instead_of_returning_zeros(config):
    net = torch.nn.Linear(2, 2).double().to(DEV)
    term = to.ODETerm(lambda t, y: net(y))
    if config == "heun":
        solver = to.AutoDiffAdjoint(to.Heun(term), to.IntegralController(1e-6, 1e-4, term=term))
    elif config == "euler":
        solver = to.AutoDiffAdjoint(to.Euler(term), to.FixedStepController())
    else:
        solver = to.AutoDiffAdjoint(to.Dopri5(term), to.FixedStepController())
    y0 = torch.ones(4, 2, dtype=torch.float64, device=DEV)
    prob = to.InitialValueProblem(y0, torch.zeros(4, dtype=torch.float64, device=DEV),
                                  torch.ones(4, dtype=torch.float64, device=DEV))
    dt0 = torch.full((4,), 0.1, dtype=torch.float64, device=DEV)
    with pytest.raises(NotImplementedError):
        solver.solve(prob, dt0=dt0)
    with torch.no_grad():  # the forward solve itself is fine
        sol = solver.solve(prob, dt0=dt0)
    assert bool(torch.isfinite(sol.ys).all())


def test_kernel_backed_fields_are_differentiable_in_the_state():
    """fields.Heat1D as f with y0.requires_grad: the replay backward differentiates the field through its
    PyTorch restatement (df/dy is not silently zero)."""
    from torchode_b200.fields import Heat1D

    field = Heat1D(3.0)
    g = torch.Generator().manual_seed(2)
    y0 = torch.rand(3, 64, generator=g, dtype=torch.float64).to(DEV).requires_grad_()
    t0, t1 = torch.zeros(3, dtype=torch.float64, device=DEV), torch.full((3,), 0.05, dtype=torch.float64, device=DEV)
    grads = []
    for f in (field, field.forward_reference):
        term = to.ODETerm(f)
        solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-8, 1e-8, term=term))
        sol = solver.solve(to.InitialValueProblem(y0, t0, t1))
        grads.append(torch.autograd.grad((sol.ys ** 2).sum(), [y0])[0])
    assert float(grads[0].abs().max()) > 0
    assert torch.allclose(grads[0], grads[1], rtol=1e-9, atol=1e-12)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_gradient_solve_on_a_device_that_is_not_current():
    dev1 = torch.device("cuda", 1)
    net = torch.nn.Linear(2, 2).double().to(dev1)
    term = to.ODETerm(lambda t, y: torch.tanh(net(y)))
    solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-7, 1e-6, term=term))
    y0 = torch.ones(5, 2, dtype=torch.float64, device=dev1, requires_grad=True)
    assert torch.cuda.current_device() == 0
    sol = solver.solve(to.InitialValueProblem(y0, torch.zeros(5, dtype=torch.float64, device=dev1),
                                              torch.ones(5, dtype=torch.float64, device=dev1)))
    (g,) = torch.autograd.grad(sol.ys.sum(), [y0])
    assert g.device == dev1 and bool(torch.isfinite(g).all()) and float(g.abs().max()) > 0

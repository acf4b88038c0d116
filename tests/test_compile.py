"""torch.compile support: inside a compiled region the solve is ONE opaque
operator with a fake kernel, so a model around it traces into a full graph."""
import pytest
import torch

import torchode_b200 as to
from torchode_b200.fields import VanDerPol


def _solver():
    term = to.ODETerm(VanDerPol(10.0))
    return to.AutoDiffAdjoint(to.Tsit5(term), to.PIDController(1e-8, 1e-8, 0.2, 0.5, 0.0, term=term))


def _model(solver):
    def fn(y0, t_end):
        prob = to.InitialValueProblem(y0 * 2.0, torch.zeros_like(t_end), t_end)
        sol = solver.solve(prob)
        return sol.ys[:, -1].sum(dim=1) + 1.0, sol.stats["n_steps"], sol.status

    return fn


def test_solve_traces_into_one_operator_without_a_gpu():
    """Dynamo export works on fake tensors: no kernel runs, the fake kernel supplies the shapes."""
    solver = _solver()
    y0, t_end = torch.ones(5, 2, dtype=torch.float64), torch.full((5,), 2.0, dtype=torch.float64)
    gm = torch._dynamo.export(_model(solver))(y0, t_end).graph_module
    targets = [str(n.target) for n in gm.graph.nodes if n.op == "call_function"]
    assert sum("torchode_b200.solve" in t for t in targets) == 1, targets
    fake_out = [n for n in gm.graph.nodes if n.op == "output"][0]
    assert fake_out is not None


@pytest.mark.gpu
def test_compiled_model_equals_eager():
    solver = _solver()
    g = torch.Generator().manual_seed(3)
    y0 = (torch.rand(64, 2, generator=g, dtype=torch.float64) * 2 - 1).cuda()
    t_end = torch.full((64,), 2.0, dtype=torch.float64, device="cuda")
    fn = _model(solver)
    want = fn(y0, t_end)
    got = torch.compile(fn, backend="eager", fullgraph=True)(y0, t_end)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_compiled_solve_refuses_gradients():
    solver = _solver()
    y0 = torch.ones(4, 2, dtype=torch.float64, device="cuda", requires_grad=True)
    t_end = torch.full((4,), 1.0, dtype=torch.float64, device="cuda")
    out = torch.compile(_model(solver), backend="eager")(y0, t_end)[0]
    with pytest.raises(NotImplementedError):
        out.sum().backward()

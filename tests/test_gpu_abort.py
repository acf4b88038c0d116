"""A failing sample aborts the whole batch (adjoints.py:186-190); without t_eval the samples that were still
running get the last iteration's interpolant evaluated at t_end (adjoints.py:298-301).  Round 1 returned the
last written value on the stage-wise route (DESIGN.md deviation 7.2, removed in round 2): the stage-wise
route now replays with the failing iteration as cap, like the fused route."""
import numpy as np
import pytest
import torch

import torchode_b200 as to
from oracle import oracle as orc
from torchode_b200.fields import LotkaVolterra

from helpers import bits_equal

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("fail_at", ["first_iteration", "later"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_abort_end_values_equal_the_oracle_on_both_routes(fail_at, dtype):
    B = 37
    g = torch.Generator().manual_seed(7)
    y0 = (1 + torch.rand(B, 2, generator=g, dtype=dtype))
    max_steps = None
    if fail_at == "first_iteration":
        y0[5, 0] = float("inf")  # INFINITE_NORM in iteration 1
    else:
        max_steps = 6  # REACHED_MAX_STEPS for everybody still running in iteration 6
    t0, t1 = torch.zeros(B, dtype=dtype), torch.full((B,), 10.0, dtype=dtype)
    t1[::3] = 0.05  # a few samples finish before the failure: they keep their own end value
    field = LotkaVolterra()
    sols = []
    for f in (field, lambda t, y: field(t, y)):  # fused kernel, stage-wise kernels
        term = to.ODETerm(f)
        solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term), max_steps=max_steps)
        with torch.no_grad():
            sols.append(solver.solve(to.InitialValueProblem(y0.to(DEV), t0.to(DEV), t1.to(DEV))))
        route = solver.last_run["route"]
        assert route.startswith("fused") if f is field else route.startswith("staged")
    m, c = to.Dopri5(), to.IntegralController(1e-6, 1e-3)
    ref = orc.solve_builtin(field.field_id, field.params(), m.to_cabi(), c.to_cabi(5, dtype, max_steps),
                            y0.numpy(), t0.numpy(), t1.numpy())
    assert (ref["status"] != 0).any() and (ref["status"] == 0).any()
    for sol in sols:
        assert sol.status.cpu().numpy().tolist() == ref["status"].tolist()
        assert sol.stats["n_steps"].cpu().numpy().tolist() == ref["n_steps"].tolist()
        assert sol.stats["n_accepted"].cpu().numpy().tolist() == ref["n_accepted"].tolist()
        assert int(sol.stats["n_f_evals"][0]) == int(ref["n_f_evals"])
        assert bits_equal(sol.ys.cpu().numpy(), ref["ys"]), "end values of the aborted batch differ from the oracle"

"""tcgen05 MLP vector field (BASELINE.json configs[3]): the kernel against a plain PyTorch fp32
reference of the same op, and the stage-wise solve around it against the oracle's loop driven
with the SAME field evaluated on the same GPU (parity for this config must share f: a bf16 GEMM
differs between devices far above the solver tolerance)."""
import numpy as np
import pytest
import torch

import torchode_b200 as to
from oracle import driver
from torchode_b200.fields import TanhMLP256

from helpers import bits_equal

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_field(n_layers=3, gain=3.0, seed=1234):
    torch.manual_seed(seed)
    layers = []
    for i in range(n_layers):
        layers.append(torch.nn.Linear(256, 256))
        if i + 1 < n_layers:
            layers.append(torch.nn.Tanh())
    seq = torch.nn.Sequential(*layers)
    with torch.no_grad():
        for p in seq.parameters():
            p.mul_(gain)
    return TanhMLP256.from_sequential(seq).to(DEV)


@pytest.mark.parametrize("B", [1, 37, 64, 65, 128, 129, 1000, 8192, 19000])  # < 148 x 128 rows: 64-row tiles, else 128-row tiles
@pytest.mark.parametrize("n_layers", [1, 2, 3])
def test_kernel_matches_fp32_reference(B, n_layers):
    field = make_field(n_layers)
    g = torch.Generator().manual_seed(B)
    y = torch.randn(B, 256, generator=g).to(DEV)
    with torch.no_grad():
        got = field(None, y)
        want = field.forward_reference(None, y)
    torch.cuda.synchronize()
    # fp32 accumulation order differs (tensor core vs cuBLAS/fp32 matmul); bf16 re-rounding of the
    # hidden activations can flip a last bf16 bit: tolerance 2e-2 of the output scale, typical 1e-5
    scale = want.abs().max()
    assert torch.isfinite(got).all()
    assert (got - want).abs().max() <= 2e-2 * scale
    assert (got - want).abs().median() <= 1e-4 * scale


def test_solve_around_the_mlp_field_matches_the_oracle_loop():
    field = make_field(3)
    B = 512
    g = torch.Generator().manual_seed(7)
    y0 = torch.randn(B, 256, generator=g)
    t0, t1 = torch.zeros(B), torch.full((B,), 1.0)

    def f_np(t, y):  # the oracle's loop calls the same GPU kernel
        with torch.no_grad():
            return field(None, torch.from_numpy(y).to(DEV)).cpu().numpy()

    method, ctrl = to.Dopri5(), to.IntegralController(1e-6, 1e-3)
    want = driver.solve_opaque(f_np, method.to_cabi(), ctrl.to_cabi(5, torch.float32), y0.numpy(), t0.numpy(),
                               t1.numpy())
    term = to.ODETerm(field)
    solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
    with torch.no_grad():
        sol = solver.solve(to.InitialValueProblem(y0.to(DEV), t0.to(DEV), t1.to(DEV)))
    assert sol.stats["n_steps"].cpu().tolist() == want["n_steps"].tolist()
    assert sol.stats["n_accepted"].cpu().tolist() == want["n_accepted"].tolist()
    assert int(sol.stats["n_f_evals"][0]) == want["n_f_evals"]
    assert bits_equal(sol.ys.cpu().numpy(), want["ys"])
    assert (sol.status == 0).all() and int(sol.stats["n_steps"].max()) >= 5

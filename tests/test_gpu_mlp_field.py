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
    # same bf16 operands, fp32 accumulation on both sides: what differs is the accumulation order (tensor core vs
    # fp32 matmul), tanh.approx (2^-11 relative) and, through them, a last bf16 bit of a hidden activation
    # (2^-8 relative of that activation, times a weight of the next layer).  Measured worst case over these
    # shapes: 1e-3 of the output scale (round 1 asserted 2e-2); bound 4e-3, median 1e-4.
    scale = want.abs().max()
    assert torch.isfinite(got).all()
    err = (got - want).abs()
    print(f"B={B} layers={n_layers}: max err {float(err.max() / scale):.2e} of the output scale, median {float(err.median() / scale):.2e}")
    assert err.max() <= 4e-3 * scale
    assert err.median() <= 1e-4 * scale


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


@pytest.mark.parametrize("B", [37, 8192, 19000])  # 64-row and 128-row tiles
@pytest.mark.parametrize("tdtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("method", [to.Dopri5, to.Tsit5])
def test_stage_fused_evaluation_equals_stage_kernel_then_field(B, tdtype, method):
    """tode_mlp_tanh256_stage_forward (stage combination formed inside the kernel's operand load) against
    tode_erk_stage followed by tode_mlp_tanh256_forward: every bit of f's output and of the stored y_i."""
    import ctypes as C

    from torchode_b200 import _cabi, _launch

    field = make_field(3)
    tab = method().to_cabi()
    g = torch.Generator().manual_seed(B)
    y = torch.randn(B, 256, generator=g).to(DEV)
    ks = [torch.randn(B, 256, generator=g).to(DEV) for _ in range(6)]
    dt = (0.01 + 0.1 * torch.rand(B, generator=g, dtype=tdtype)).to(DEV)
    st = _launch._minimal_state(y, dt)
    lib, stream = _cabi.lib(), _launch.stream_ptr(y.device)
    for stage in range(1, 7):
        y_i = _launch.erk_stage(tab, stage, y, dt, ks[:stage])
        want = field(None, y_i)
        for store in (False, True):
            y_out = torch.full_like(y, float("nan"))
            got = torch.empty_like(y)
            _cabi.check(lib.tode_mlp_tanh256_stage_forward(
                C.byref(tab), stage, C.byref(st), _launch.kptrs(ks[:stage]), y_out.data_ptr() if store else None,
                field.weights.data_ptr(), field.biases.data_ptr(), got.data_ptr(), field.n_layers, stream), "stage mlp")
            torch.cuda.synchronize()
            assert bits_equal(got.cpu().numpy(), want.cpu().numpy()), (stage, store)
            if store:
                assert bits_equal(y_out.cpu().numpy(), y_i.cpu().numpy()), stage
    # a set stop flag turns the launch into a no-op
    ctl = torch.zeros(_cabi.CTL_WORDS, dtype=torch.int32, device=DEV)
    ctl[_cabi.CTL_STOP] = 1
    st.ctl = ctl.data_ptr()
    got = torch.full_like(y, 7.0)
    _cabi.check(lib.tode_mlp_tanh256_stage_forward(
        C.byref(tab), 3, C.byref(st), _launch.kptrs(ks[:3]), None, field.weights.data_ptr(),
        field.biases.data_ptr(), got.data_ptr(), field.n_layers, stream), "stage mlp")
    torch.cuda.synchronize()
    assert (got == 7.0).all()


@pytest.mark.parametrize("B,n_layers", [(700, 3), (8192, 3), (19000, 3), (1000, 1), (1000, 2), (37, 4)])  # 19000: 128-row tiles
@pytest.mark.parametrize("method", [to.Dopri5, to.Tsit5])
def test_all_stages_in_one_launch_equal_one_launch_per_stage(B, n_layers, method):
    """tode_mlp_tanh256_step_forward (round 2: a CTA walks the six stages of its rows; with 64-row tiles the
    operand rows come from partial sums formed under the previous stage's MMAs -- scheduled by layer, hence the
    layer counts) against six tode_mlp_tanh256_stage_forward launches: same bits in every k_i, y1 and the solve."""
    field = make_field(n_layers)
    g = torch.Generator().manual_seed(B)
    problem = to.InitialValueProblem(torch.randn(B, 256, generator=g).to(DEV), torch.zeros(B, device=DEV),
                                     (0.3 + 0.4 * torch.rand(B, generator=g)).to(DEV))
    term = to.ODETerm(field)
    sols = {}
    for mode in (True, "stages"):
        solver = to.AutoDiffAdjoint(method(term), to.IntegralController(1e-6, 1e-3, term=term))
        solver.use_step_fusion = mode
        solver.use_cuda_graph = False
        with torch.no_grad():
            sols[mode] = solver.solve(problem)
        assert solver.last_run["route"] == "stage-fused"
    a, b = sols[True], sols["stages"]
    assert bits_equal(a.ys.cpu().numpy(), b.ys.cpu().numpy())
    for key in ("n_steps", "n_accepted", "n_f_evals"):
        assert a.stats[key].cpu().tolist() == b.stats[key].cpu().tolist()
    assert (a.status == 0).all()


@pytest.mark.parametrize("graph", [False, True])
def test_stage_fused_route_is_bit_identical_to_the_stage_wise_route(graph):
    field = make_field(3)
    B = 700
    g = torch.Generator().manual_seed(11)
    problem = to.InitialValueProblem(torch.randn(B, 256, generator=g).to(DEV), torch.zeros(B, device=DEV),
                                     (0.5 + torch.rand(B, generator=g)).to(DEV))
    term = to.ODETerm(field)
    sols = {}
    for fusion in (True, False):
        solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
        solver.use_step_fusion = fusion
        solver.use_cuda_graph = graph
        with torch.no_grad():
            for _ in range(2 if graph else 1):  # second solve replays the cached plan
                sols[fusion] = solver.solve(problem)
        want_route = ("stage-fused" if fusion else "staged") + ("+graph" if graph else "")
        assert solver.last_run["route"] == want_route
    a, b = sols[True], sols[False]
    assert bits_equal(a.ys.cpu().numpy(), b.ys.cpu().numpy())
    for key in ("n_steps", "n_accepted", "n_f_evals"):
        assert a.stats[key].cpu().tolist() == b.stats[key].cpu().tolist()
    assert (a.status == 0).all() and int(a.stats["n_steps"].max()) >= 5

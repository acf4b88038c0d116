"""GPU kernel-level parity: every C-ABI op against the oracle's restatement of the same
reference lines, bit-exact, over the feature-dimension geometries the kernels special-case
(thread-per-sample, warp-per-sample with register cache, warp-per-sample streaming)."""
import ctypes as C

import numpy as np
import pytest
import torch

import torchode_b200 as to
from oracle import driver
from oracle import oracle as orc
from torchode_b200 import _cabi, _launch
from torchode_b200.single_step_methods import StepResult
from torchode_b200.step_size_controllers import max_norm

from helpers import bits_equal, ulps

pytestmark = pytest.mark.gpu
DEV = "cuda"
NP = {"f32": np.float32, "f64": np.float64}
DTYPES = [("f32", "f32"), ("f64", "f64"), ("f32", "f64"), ("f64", "f32")]
FEATS = [1, 2, 3, 4, 6, 8, 256, 260, 1000]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def rand_step(rng, B, F, D, T, n_stages=7):
    y0 = rng.normal(size=(B, F)).astype(D)
    ks = [rng.normal(size=(B, F)).astype(D) for _ in range(n_stages)]
    dt = (rng.uniform(0.01, 0.2, size=B) * rng.choice([-1, 1], size=B)).astype(T)
    return y0, ks, dt


@pytest.mark.parametrize("dd,td", DTYPES)
@pytest.mark.parametrize("F", FEATS)
@pytest.mark.parametrize("method", [to.Dopri5, to.Tsit5])
def test_stage_kernel(dd, td, F, method):
    rng = np.random.default_rng(F * 7 + len(dd + td))
    B = 37 if F > 8 else 1031
    D, T = NP[dd], NP[td]
    tab = method().to_cabi()
    y0, ks, dt = rand_step(rng, B, F, D, T)
    st = orc.HostState(y0, np.zeros(B, T), np.ones(B, T), None)
    st.dt[:] = dt
    st.running[::5] = 0  # finished rows are neither read nor written
    for stage in range(1, 7):
        want = orc.erk_stage(tab, stage, st, ks[:stage], np.full((B, F), 7.0, D))
        out = torch.full((B, F), 7.0, dtype=cu(y0).dtype, device=DEV)
        g = _launch.StagedState(to.InitialValueProblem(cu(y0), cu(np.zeros(B, T)), cu(np.ones(B, T))), 7, False)
        g.dt.copy_(cu(dt))
        g.running.copy_(cu(st.running))
        kd = [cu(k) for k in ks[:stage]]
        _cabi.check(_cabi.lib().tode_erk_stage(C.byref(tab), stage, C.byref(g.c), _launch.kptrs(kd),
                                               out.data_ptr(), _launch.stream_ptr(out.device)), "stage")
        assert bits_equal(out.cpu().numpy(), want), f"stage {stage}"


@pytest.mark.parametrize("dd,td", DTYPES)
@pytest.mark.parametrize("F", FEATS)
@pytest.mark.parametrize("kind", ["integral", "pid", "pid_max_limits"])
def test_adapt_step_size(dd, td, F, kind):
    rng = np.random.default_rng(F + 100)
    B = 29 if F > 8 else 517
    D, T = NP[dd], NP[td]
    tD = torch.from_numpy(np.zeros(1, D)).dtype
    if kind == "integral":
        ctrl = to.IntegralController(1e-6, 1e-3)
    elif kind == "pid":
        ctrl = to.PIDController(1e-7, 1e-5, 0.2, 0.5, 0.1)
    else:
        ctrl = to.PIDController(1e-4, 1e-3, 0.1, 0.6, 0.05, norm=max_norm, dt_min=0.02, dt_max=0.11,
                                safety=0.8, factor_min=0.3, factor_max=4.0)
    cab = ctrl.to_cabi(5, tD)
    y0 = rng.normal(size=(B, F)).astype(D)
    y1 = (y0 + 0.1 * rng.normal(size=(B, F))).astype(D)
    err = (rng.normal(size=(B, F)) * 10.0 ** rng.uniform(-9, -2, size=(B, 1))).astype(D)
    err[0] = 0.0           # zero error -> almost_zero floor -> factor_max (step_size_controllers_test.py:43-62)
    err[1, 0] = np.inf     # -> INFINITE_NORM (:65-95)
    y1[2, -1] = np.nan     # NaN state -> INFINITE_NORM (:98-122)
    dt = (rng.uniform(0.01, 0.2, size=B) * rng.choice([-1, 1], size=B)).astype(T)
    r1 = rng.uniform(0.1, 2.0, size=B).astype(D)
    r2 = rng.uniform(0.1, 2.0, size=B).astype(D)
    want = orc.adapt_step_size(cab, dt, y0, y1, err, r1, r2)

    class S:
        method_order, dt_min, dt_max = 5, None, None
        prev_error_ratio, prev_prev_error_ratio = cu(r1), cu(r2)

    accept, dt_next, r1o, r2o, status = _launch.adapt_step_size(ctrl, S, cu(dt), cu(y0), cu(y1), cu(err))
    assert accept.cpu().numpy().tolist() == want["accept"].tolist()
    assert bits_equal(dt_next.cpu().numpy(), want["dt_next"])
    assert status.cpu().numpy().tolist() == want["status"].tolist()
    if kind != "pid_max_limits":  # with dt_min, REACHED_DT_MIN may override (:419-422), as in the oracle
        assert status[1].item() == to.Status.INFINITE_NORM.value
        assert status[2].item() == to.Status.INFINITE_NORM.value
    assert not accept[1] and not accept[2] and accept[0]
    if kind != "integral":
        assert bits_equal(r1o.cpu().numpy(), want["r1"]) and bits_equal(r2o.cpu().numpy(), want["r2"])


@pytest.mark.parametrize("dd,td", DTYPES)
@pytest.mark.parametrize("method", [to.Dopri5, to.Tsit5])
def test_weighted_sum_time_nodes_interp(dd, td, method):
    rng = np.random.default_rng(5)
    B, F = 211, 5
    D, T = NP[dd], NP[td]
    m = method()
    tab = m.to_cabi()
    y0, ks, dt = rand_step(rng, B, F, D, T)
    t0 = rng.normal(size=B).astype(T)
    kd = [cu(k) for k in ks]
    for which, name in ((_cabi.W_BERR, "b_err"), (_cabi.W_B, "b")):
        want = orc.erk_weighted_sum(tab, which, dt, ks, y0 if name == "b" else None)
        got = _launch.erk_weighted_sum(tab, name, cu(dt), kd, base=cu(y0) if name == "b" else None)
        assert bits_equal(got.cpu().numpy(), want)
    nodes = _launch.time_nodes(tab, cu(t0), cu(dt)).cpu().numpy()
    c = np.array([tab.c[i] for i in range(7)]).astype(T)
    assert np.abs(nodes - (t0[None] + c[:, None] * dt[None])).max() < 1e-5
    # dense output at random points of random samples
    N = 400
    idx = rng.integers(0, B, size=N)
    tq = (t0[idx] + rng.uniform(0, 1, size=N) * dt[idx]).astype(T)
    y1 = (y0 + rng.normal(size=(B, F)) * 0.01).astype(D)
    want = orc.interp_eval(tab, t0, dt, y0, y1, ks, tq, idx)
    got = _launch.interp_eval(tab, cu(t0), cu(dt), cu(y0), cu(y1), torch.stack(kd), cu(tq), cu(idx))
    assert bits_equal(got.cpu().numpy(), want)


def _gpu_state_like(st, tab, ctrl_pid, t_eval=None, general=False):
    prob = to.InitialValueProblem(cu(st.y), cu(st.t_start), cu(st.t_end), None if t_eval is None else cu(t_eval))
    g = _launch.StagedState(prob, tab.n_stages, ctrl_pid, general=general)
    for name in ("t", "dt", "y", "f0", "running", "n_steps", "n_accepted", "status", "cursor"):
        getattr(g, name).copy_(cu(getattr(st, name)))
    if ctrl_pid:
        g.r1.copy_(cu(st.r1))
        g.r2.copy_(cu(st.r2))
    g.y_eval.copy_(cu(st.y_eval))
    return g


@pytest.mark.parametrize("dd,td", [("f32", "f32"), ("f64", "f64"), ("f32", "f64")])
@pytest.mark.parametrize("F", [1, 2, 4, 12, 128, 256, 520, 3000, 9000])
@pytest.mark.parametrize("with_t_eval", [False, True])
def test_finish_kernel_one_iteration(dd, td, F, with_t_eval):
    """One loop iteration on a random mid-solve state: every state array after tode_erk_finish
    equals the oracle's (covers the thread-per-sample, lane-group, warp-per-sample streaming and
    -- F = 9000 -- the three-launch split-mode variants of the finish kernel)."""
    rng = np.random.default_rng(F)
    B = 300 if F <= 12 else 41
    D, T = NP[dd], NP[td]
    method, ctrl = to.Tsit5(), to.PIDController(1e-5, 1e-4, 0.2, 0.5, 0.05)
    tab = method.to_cabi()
    cab = ctrl.to_cabi(5, torch.from_numpy(np.zeros(1, D)).dtype, 50)
    y0, ks, dt = rand_step(rng, B, F, D, T)
    dt = np.abs(dt)
    t_start = np.zeros(B, T)
    t_end = rng.uniform(0.05, 1.0, size=B).astype(T)
    t_eval = None
    if with_t_eval:
        t_eval = (t_end[:, None] * np.linspace(0, 1, 9)[None]).astype(T)
    st = orc.HostState(y0, t_start, t_end, t_eval, pid=True)
    st.t[:] = rng.uniform(0.0, 0.04, size=B).astype(T)
    st.dt[:] = dt
    st.f0[:] = ks[0]
    st.r1[:] = rng.uniform(0.2, 1.5, size=B).astype(D)
    st.r2[:] = rng.uniform(0.2, 1.5, size=B).astype(D)
    st.n_steps[:] = rng.integers(0, 50, size=B)
    st.n_accepted[:] = st.n_steps // 2
    st.running[::7] = 0
    st.cursor[:] = 1
    st.y_eval[:] = 0
    # small errors so that a good share of the steps is accepted: k ~ O(1), weights sum to 0
    ks = [ks[0]] + [(ks[0] + 1e-3 * rng.normal(size=(B, F))).astype(D) for _ in range(6)]
    y1 = (y0 + dt[:, None].astype(D) * ks[0]).astype(D)
    g = _gpu_state_like(st, tab, True, t_eval)
    orc.erk_finish(tab, cab, st, ks, y1)
    kd = [g.f0] + [cu(k) for k in ks[1:]]
    _cabi.check(_cabi.lib().tode_erk_finish(C.byref(tab), C.byref(cab), C.byref(g.c), _launch.kptrs(kd),
                                            cu(y1).data_ptr(), _launch.stream_ptr(g.device)), "finish")
    torch.cuda.synchronize()
    assert 0.05 < st.n_accepted.sum() / max(1, st.n_steps.sum())  # the scenario exercises the commit
    for name in ("t", "dt", "y", "f0", "r1", "r2", "running", "n_steps", "n_accepted", "status", "y_eval"):
        assert bits_equal(getattr(g, name).cpu().numpy(), getattr(st, name)), name
    if with_t_eval:
        assert g.cursor.cpu().numpy().tolist() == st.cursor.tolist()
    run = st.running.astype(bool)
    assert bits_equal(g.t_nodes.cpu().numpy()[1:, run], st.t_nodes[1:, run])
    ctl = g.ctl.cpu().numpy()
    assert ctl[_cabi.CTL_ITERS] == 1 and ctl[_cabi.CTL_STOP] == st.ctl[_cabi.CTL_STOP]
    assert ctl[_cabi.CTL_RUNNING] == 0 and ctl[_cabi.CTL_TICKET] == 0  # scratch words reset


@pytest.mark.parametrize("dd,td", [("f32", "f32"), ("f64", "f64"), ("f32", "f64")])
@pytest.mark.parametrize("F", [2, 12, 256, 1000, 9000, 8193, 16384 + 4])
@pytest.mark.parametrize("mode", ["hairer", "hairer+t_eval", "dt0"])
def test_init_kernels(dd, td, F, mode):
    """tode_init_step_a / _b / tode_init_with_dt0 against the oracle: thread-per-sample, warp-per-sample and
    -- rows of >= 2048 vectors -- the split variants (one warp per chunk, erk_init_split.cuh)."""
    rng = np.random.default_rng(F + len(mode))
    B = 200 if F <= 12 else 23
    D, T = NP[dd], NP[td]
    method, ctrl = to.Dopri5(), to.PIDController(1e-5, 1e-4, 0.2, 0.5, 0.05)
    tab = method.to_cabi()
    cab = ctrl.to_cabi(5, torch.from_numpy(np.zeros(1, D)).dtype)
    y0 = rng.normal(size=(B, F)).astype(D)
    y0[3] *= 1e-7  # d0 < 1e-5 branch
    f0 = rng.normal(size=(B, F)).astype(D)
    f0[5] = 0      # d1 < 1e-5 branch
    t_start = rng.uniform(-1, 1, size=B).astype(T)
    t_end = (t_start + rng.uniform(0.05, 1.0, size=B) * rng.choice([-1, 1], size=B)).astype(T)
    t_eval = None
    if mode == "hairer+t_eval":
        t_eval = (t_start[:, None] + (t_end - t_start)[:, None] * np.linspace(0, 1, 9)[None]).astype(T)
        t_eval[::3, 0] += (t_end - t_start)[::3] * T(0.01)  # rows whose first point is not t_start
        t_eval[7, [2, 5]] = t_eval[7, [5, 2]]                # one non-monotone row
    st = orc.HostState(y0, t_start, t_end, t_eval, pid=True)
    st.f0[:] = f0
    st.y_eval[:] = 0
    g = _gpu_state_like(st, tab, True, t_eval)
    lib, stream = _cabi.lib(), _launch.stream_ptr(g.device)
    if mode == "dt0":
        dt0 = (rng.uniform(1e-3, 2.0, size=B) * np.sign(t_end - t_start)).astype(T)
        orc.init_with_dt0(tab, cab, st, dt0)
        _cabi.check(lib.tode_init_with_dt0(C.byref(tab), C.byref(cab), C.byref(g.c), cu(dt0).data_ptr(), stream), "init")
    else:
        y1, t1 = np.zeros((B, F), D), np.zeros(B, T)
        orc.init_step_a(tab, cab, st, y1, t1)
        y1g, t1g = torch.zeros((B, F), dtype=g.y.dtype, device=DEV), torch.zeros(B, dtype=g.t.dtype, device=DEV)
        _cabi.check(lib.tode_init_step_a(C.byref(tab), C.byref(cab), C.byref(g.c), y1g.data_ptr(), t1g.data_ptr(),
                                         stream), "init a")
        assert bits_equal(y1g.cpu().numpy(), y1) and bits_equal(t1g.cpu().numpy(), t1)
        f1 = (f0 + 0.1 * rng.normal(size=(B, F))).astype(D)
        f1[9] = f0[9]  # max(d1, d2) tiny branch needs d1 tiny as well: sample 5 has f0 = 0
        f1[5] = 0
        orc.init_step_b(tab, cab, st, f1)
        _cabi.check(lib.tode_init_step_b(C.byref(tab), C.byref(cab), C.byref(g.c), cu(f1).data_ptr(), stream), "init b")
    torch.cuda.synchronize()
    for name in ("t", "dt", "r1", "r2", "running", "n_steps", "n_accepted", "status", "y_eval"):
        assert bits_equal(getattr(g, name).cpu().numpy(), getattr(st, name)), name
    if t_eval is not None:
        assert g.cursor.cpu().numpy().tolist() == st.cursor.tolist()
    assert bits_equal(g.t_nodes.cpu().numpy()[1:], st.t_nodes[1:])
    ctl = g.ctl.cpu().numpy()
    assert ctl[_cabi.CTL_NONMONO] == st.ctl[_cabi.CTL_NONMONO] == (1 if t_eval is not None else 0)
    assert ctl[_cabi.CTL_STOP] == 0 and ctl[_cabi.CTL_ITERS] == 0


def heat_rhs_np(kappa):
    def f(t, y):
        out = np.zeros_like(y)
        out[:, 1:-1] = y.dtype.type(kappa) * ((y[:, 2:] - y.dtype.type(2) * y[:, 1:-1]) + y[:, :-2])
        return out
    return f


def heat_rhs_torch(kappa):
    def f(t, y):
        out = torch.zeros_like(y)
        out[:, 1:-1] = kappa * ((y[:, 2:] - 2 * y[:, 1:-1]) + y[:, :-2])
        return out
    return f


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("shape", [(1, 4), (3, 8), (5, 1000), (2, 65536)])
def test_heat1d_field_kernel_is_bit_identical_to_the_torch_expression(shape, dtype):
    from torchode_b200.fields import Heat1D

    g = torch.Generator().manual_seed(shape[1])
    y = torch.randn(*shape, generator=g, dtype=dtype).cuda()
    field = Heat1D(25.0)
    got, want = field(None, y), field.forward_reference(None, y)
    assert bits_equal(got.cpu().numpy(), want.cpu().numpy())


@pytest.mark.parametrize("field_kind", ["torch", "kernel"])
@pytest.mark.parametrize("N", [64, 1024, 4100, 16384])
def test_staged_opaque_heat_equation_matches_oracle(N, field_kind):
    """Config C5 in miniature: method-of-lines heat equation (opaque stencil f -- the PyTorch
    expression or the one-pass Heat1D kernel), Tsit5 + I."""
    rng = np.random.default_rng(N)
    B = 6
    x = np.linspace(0, 1, N, dtype=np.float32)
    amp = rng.uniform(size=(B, 3)).astype(np.float32)
    y0 = sum(amp[:, k - 1:k] * np.sin(np.float32(k * np.pi) * x)[None] for k in (1, 2, 3)).astype(np.float32)
    kappa = 20.0  # stencil without the 1/dx^2 factor: spectral radius 4*kappa = 80, non-stiff
    t0, t1 = np.zeros(B, np.float32), np.full(B, 0.5, np.float32)
    method = to.Tsit5()
    ctrl = to.IntegralController(1e-6, 1e-3)
    want = driver.solve_opaque(heat_rhs_np(kappa), method.to_cabi(), ctrl.to_cabi(5, torch.float32), y0, t0, t1)
    from torchode_b200.fields import Heat1D

    term = to.ODETerm(heat_rhs_torch(kappa) if field_kind == "torch" else Heat1D(kappa))
    solver = to.AutoDiffAdjoint(to.Tsit5(term), to.IntegralController(1e-6, 1e-3, term=term))
    with torch.no_grad():
        sol = solver.solve(to.InitialValueProblem(cu(y0), cu(t0), cu(t1)))
    assert sol.stats["n_steps"].cpu().tolist() == want["n_steps"].tolist()
    assert sol.stats["n_accepted"].cpu().tolist() == want["n_accepted"].tolist()
    assert int(sol.stats["n_f_evals"][0]) == want["n_f_evals"]
    assert sol.status.cpu().tolist() == want["status"].tolist()
    assert bits_equal(sol.ys.cpu().numpy(), want["ys"])


def _heat_problem(B, N, dtype, tdtype, seed, reverse=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.linspace(0, 1, N, dtype=dtype)
    amp = torch.rand(B, 3, generator=g, dtype=dtype)
    y0 = sum(amp[:, k - 1:k] * torch.sin(k * torch.pi * x)[None] for k in (1, 2, 3))
    y0 = y0 + 0.01 * torch.randn(B, N, generator=g, dtype=dtype)  # rough rows: every chunk contributes to the norm
    t_a = torch.zeros(B, dtype=tdtype)
    t_b = 0.5 + 0.5 * torch.rand(B, generator=g, dtype=tdtype)  # samples finish at different iterations
    if reverse:
        t_a, t_b = t_b, t_a
    return to.InitialValueProblem(y0.to(DEV), t_a.to(DEV), t_b.to(DEV))


def _solve_both_heat_routes(problem, make_solver, dt0=None):
    out = []
    for fusion in (True, False):
        solver = make_solver()
        solver.use_step_fusion = fusion
        with torch.no_grad():
            sol = solver.solve(problem, dt0=dt0)
        out.append((sol, dict(solver.last_run)))
    return out


def _assert_same_solution(a, b):
    assert bits_equal(a.ys.cpu().numpy(), b.ys.cpu().numpy())
    for key in ("n_steps", "n_accepted", "n_initialized", "n_f_evals"):
        assert a.stats[key].cpu().tolist() == b.stats[key].cpu().tolist(), key
    assert a.status.cpu().tolist() == b.status.cpu().tolist()


HEAT_CASES = [
    # B, N, data dtype, time dtype, method, controller kind
    (3, 8, torch.float32, torch.float32, "tsit5", "i"),
    (5, 1000, torch.float32, torch.float32, "dopri5", "i"),
    (4, 4100, torch.float32, torch.float32, "tsit5", "pid"),      # 1025 vectors: a one-vector tail chunk
    (2, 3 * 4096 + 1024 + 4, torch.float32, torch.float32, "tsit5", "i"),  # four chunks, partial last tile
    (3, 2048 + 6, torch.float64, torch.float64, "tsit5", "i"),    # fp64: 2-element vectors, 2 chunks
    (3, 5000, torch.float64, torch.float64, "dopri5", "pid"),
    (2, 4100, torch.float32, torch.float64, "dopri5", "max"),     # mixed dtypes, max norm
    (2, 2050, torch.float64, torch.float32, "tsit5", "i"),
    (70, 16384, torch.float32, torch.float32, "tsit5", "i"),      # the split-finish shape of the stage-wise route
]


@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("B,N,dtype,tdtype,method,kind", HEAT_CASES)
def test_step_fused_heat_route_is_bit_identical_to_the_stage_wise_route(B, N, dtype, tdtype, method, kind, reverse):
    """tode_heat_step (whole iteration in one pass, stage values on chip) against 6 x (stage kernel,
    Heat1D kernel) + finish kernel -- which test_staged_opaque_heat_equation_matches_oracle pins on the
    oracle: every bit of ys and every counter."""
    from torchode_b200.fields import Heat1D

    problem = _heat_problem(B, N, dtype, tdtype, seed=N + B, reverse=reverse)

    def make_solver():
        term = to.ODETerm(Heat1D(20.0))
        step = (to.Tsit5 if method == "tsit5" else to.Dopri5)(term)
        if kind == "pid":
            ctrl = to.PIDController(1e-6, 1e-4, 0.2, 0.5, 0.1, term=term)
        elif kind == "max":
            ctrl = to.IntegralController(1e-6, 1e-3, term=term, norm=max_norm)
        else:
            ctrl = to.IntegralController(1e-6, 1e-3, term=term)
        return to.AutoDiffAdjoint(step, ctrl)

    (fused, run_f), (staged, run_s) = _solve_both_heat_routes(problem, make_solver)
    assert run_f["route"] == "step-fused" and run_s["route"] == "staged"
    assert run_f["iterations"] == run_s["iterations"] > 3
    assert (fused.status == 0).all()
    _assert_same_solution(fused, staged)


@pytest.mark.parametrize("layout", ["rows", "broadcast", "reverse", "non-monotone"])
@pytest.mark.parametrize("N,dtype,method", [(4100, torch.float32, "tsit5"), (2054, torch.float64, "dopri5"),
                                            (3 * 4096 + 1024 + 4, torch.float32, "dopri5")])
def test_step_fused_heat_route_dense_output(N, dtype, method, layout):
    """t_eval with the step-fused kernels: the points a step covers are evaluated before its accept
    decision is known (rows of a rejected step are rewritten) -- same bits, same n_initialized as the
    stage-wise finish kernel; rows that are not monotone in time hand over to its scan-all mode."""
    from torchode_b200.fields import Heat1D

    B = 5
    base = _heat_problem(B, N, dtype, dtype, seed=N, reverse=layout == "reverse")
    g = torch.Generator().manual_seed(N + 1)
    lo, hi = base.t_start.cpu(), base.t_end.cpu()
    frac = torch.sort(torch.rand(B, 17, generator=g, dtype=dtype), dim=1).values
    frac[::2, 0] = 0.0  # rows whose first point is t_start
    frac[1, -1] = 1.0   # a row whose last point is t_end
    t_eval = lo[:, None] + (hi - lo)[:, None] * frac
    if layout == "broadcast":
        problem = to.InitialValueProblem(base.y0, base.t_start, torch.full_like(base.t_end, 0.5),
                                         torch.linspace(0, 0.5, 23, dtype=dtype, device=DEV).expand(B, -1))
    else:
        if layout == "non-monotone":
            t_eval[2, [3, 9]] = t_eval[2, [9, 3]]
        problem = to.InitialValueProblem(base.y0, base.t_start, base.t_end, t_eval.to(DEV))

    def make_solver():
        term = to.ODETerm(Heat1D(20.0))
        step = (to.Tsit5 if method == "tsit5" else to.Dopri5)(term)
        return to.AutoDiffAdjoint(step, to.IntegralController(1e-6, 1e-3, term=term))

    (fused, run_f), (staged, run_s) = _solve_both_heat_routes(problem, make_solver)
    if layout == "non-monotone":
        assert run_f["route"] == "staged" and run_f["general"]
    else:
        assert run_f["route"] == "step-fused" and run_s["route"] == "staged"
        assert int(fused.stats["n_steps"].sum()) > int(fused.stats["n_accepted"].sum())  # some steps were rejected
    assert (fused.status == 0).all()
    assert fused.stats["n_initialized"].cpu().tolist() == [problem.n_evaluation_points] * B
    _assert_same_solution(fused, staged)


def test_step_fused_heat_route_with_dt0_and_graph_replay():
    from torchode_b200.fields import Heat1D

    problem = _heat_problem(4, 8192, torch.float32, torch.float32, seed=5)
    dt0 = torch.full((4,), 1e-4, device=DEV)
    term = to.ODETerm(Heat1D(20.0))

    def make_solver():
        return to.AutoDiffAdjoint(to.Tsit5(term), to.IntegralController(1e-6, 1e-3, term=term))

    (fused, run_f), (staged, _) = _solve_both_heat_routes(problem, make_solver, dt0=dt0)
    assert run_f["route"] == "step-fused"
    _assert_same_solution(fused, staged)
    solver = make_solver()
    solver.use_cuda_graph = True
    with torch.no_grad():
        for _ in range(2):  # second solve replays the cached plan
            sol = solver.solve(problem, dt0=dt0)
            assert solver.last_run["route"] == "step-fused+graph"
            _assert_same_solution(sol, staged)


@pytest.mark.parametrize("max_steps,dt_min", [(5, None), (None, 0.05)])
def test_step_fused_heat_route_hands_failing_problems_to_the_stage_wise_route(max_steps, dt_min):
    """A step that ends with status != SUCCESS has no end-point value in the step-fused kernels: the
    solve is redone on the stage-wise kernels, so failures look exactly as they do there."""
    from torchode_b200.fields import Heat1D

    problem = _heat_problem(3, 4100, torch.float32, torch.float32, seed=9)
    term = to.ODETerm(Heat1D(20.0))

    def make_solver():
        kw = {} if dt_min is None else {"dt_min": dt_min}
        return to.AutoDiffAdjoint(to.Tsit5(term), to.IntegralController(1e-6, 1e-3, term=term, **kw),
                                  max_steps=max_steps)

    (fused, run_f), (staged, run_s) = _solve_both_heat_routes(problem, make_solver)
    assert run_f["route"] == "staged" == run_s["route"]
    assert (staged.status != 0).any()
    _assert_same_solution(fused, staged)


def test_general_mode_for_non_monotone_t_eval():
    rng = np.random.default_rng(3)
    B, F = 33, 2
    y0 = (1 + rng.uniform(size=(B, F))).astype(np.float32)
    t_eval = np.tile(np.linspace(0, 2, 9, dtype=np.float32), (B, 1))
    t_eval[:, [2, 5]] = t_eval[:, [5, 2]]
    t0, t1 = np.zeros(B, np.float32), np.full(B, 2.0, np.float32)
    field = to.fields.LotkaVolterra()

    def f_np(t, y):
        x, z = y[:, 0], y[:, 1]
        xz = x * z
        return np.stack((np.float32(1.5) * x - np.float32(1.0) * xz, np.float32(1.0) * xz - np.float32(3.0) * z), 1)

    want = driver.solve_opaque(f_np, to.Dopri5().to_cabi(), to.IntegralController(1e-6, 1e-3).to_cabi(5, torch.float32),
                               y0, t0, t1, t_eval)
    for f in (field, lambda t, y: field(t, y)):  # the fused route must hand over to the staged general mode
        term = to.ODETerm(f)
        solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
        with torch.no_grad():
            sol = solver.solve(to.InitialValueProblem(cu(y0), cu(t0), cu(t1), cu(t_eval)))
        assert sol.stats["n_steps"].cpu().tolist() == want["n_steps"].tolist()
        assert sol.stats["n_initialized"].cpu().tolist() == want["n_initialized"].tolist()
        assert bits_equal(sol.ys.cpu().numpy(), want["ys"])


def test_protocol_step_and_custom_controller_route():
    """Built-in Dopri5 under a foreign controller object: the generic loop calls Dopri5.step /
    build_interpolation, which run the same CUDA kernels (stage, weighted sum, interp)."""
    from torchode_b200.step_size_controllers import StepSizeController

    class HalvingController(StepSizeController):
        def init(self, term, problem, order, dt0, *, stats, args):
            return torch.full_like(problem.t_start, 0.125), None, None

        def adapt_step_size(self, t0, dt, y0, step_result, state, stats):
            assert step_result.error_estimate is not None
            return torch.ones_like(dt, dtype=torch.bool), dt, state, None

        def merge_states(self, running, current, previous):
            return current

    B = 16
    y0 = torch.linspace(1, 2, B, device=DEV)[:, None].repeat(1, 3)
    t_eval = torch.linspace(0, 1, 5, device=DEV).repeat(B, 1)
    term = to.ODETerm(lambda t, y: -0.5 * y)
    solver = to.AutoDiffAdjoint(to.Dopri5(term), HalvingController())
    with torch.no_grad():
        sol = solver.solve(to.InitialValueProblem(y0, t_eval=t_eval))
    assert sol.stats["n_steps"].tolist() == [8] * B and sol.stats["n_f_evals"].tolist() == [1 + 6 * 8] * B
    want = y0[:, None, :] * torch.exp(-0.5 * t_eval)[:, :, None]
    assert torch.allclose(sol.ys, want, rtol=1e-5)
    # stand-alone protocol call: one step equals the oracle's ops
    method = to.Dopri5(term)
    prob = to.InitialValueProblem(y0, t_eval=t_eval)
    stats = {}
    term.init(prob, stats)
    state = method.init(None, prob, None, stats=stats, args=None)
    dt = torch.full((B,), 0.1, device=DEV)
    res, interp_data, state2, status = method.step(None, None, y0, prob.t_start, dt, state, stats=stats, args=None)
    assert status is None and isinstance(res, StepResult)
    tab = method.to_cabi()
    ks = [k.cpu().numpy() for k in interp_data.k]
    err = orc.erk_error_estimate(tab, dt.cpu().numpy(), ks)
    assert bits_equal(res.error_estimate.cpu().numpy(), err)
    assert torch.equal(state2.prev_vf1, interp_data.k[-1])


def test_fixed_step_controller_with_builtin_methods():
    """fixed_step_controller_test.py:10-19: fixed steps vs the analytic solution (rel 1e-2)."""
    B = 4
    y0 = torch.ones(B, 2, device=DEV, dtype=torch.float64)
    t_eval = torch.linspace(0, 2, 6, device=DEV, dtype=torch.float64).repeat(B, 1)
    for method_cls in (to.Dopri5, to.Tsit5):
        term = to.ODETerm(lambda t, y: -y)
        solver = to.AutoDiffAdjoint(method_cls(term), to.FixedStepController())
        with torch.no_grad():
            sol = solver.solve(to.InitialValueProblem(y0, t_eval=t_eval),
                               dt0=torch.full((B,), 0.05, device=DEV, dtype=torch.float64))
        assert torch.allclose(sol.ys[:, :, 0], torch.exp(-t_eval), rtol=1e-2)
        assert (sol.status == 0).all()


def test_solve_ivp_default_matches_analytic_solution():
    """interface_test.py:7-12: tsit5 + PID(1e-7) on y' = 2y/t + t^4 sin 2t - t^2 + 4t^3, rel 1e-5."""
    def sol_fn(t):
        return (-0.5 * t**4 * torch.cos(2 * t) + 0.5 * t**3 * torch.sin(2 * t) + 0.25 * t**2 * torch.cos(2 * t)
                - t**3 + 2 * t**4 + (torch.pi - 0.25) * t**2)[..., None]

    def dyn(t, y):
        return (2 * y[:, 0] / t + t**4 * torch.sin(2 * t) - t**2 + 4 * t**3)[..., None]

    t_eval = torch.tensor([[1.0, 1.5, 2.0, 3.0], [0.5, 1.0, 1.5, 2.5]], device=DEV, dtype=torch.float64)
    with torch.no_grad():
        sol = to.solve_ivp(dyn, sol_fn(t_eval[:, 0]), t_eval)
    assert (sol.status == 0).all()
    assert torch.allclose(sol.ys, sol_fn(t_eval), rtol=1e-5)


def test_extra_args_reach_f_by_identity():
    """extra_args_test.py:10-22."""
    marker, seen = object(), []
    term = to.ODETerm(lambda t, y, a: seen.append(a) or -y, with_args=True)
    solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
    y0 = torch.ones(3, 2, device=DEV)
    with torch.no_grad():
        solver.solve(to.InitialValueProblem(y0, torch.zeros(3, device=DEV), torch.ones(3, device=DEV)), args=marker)
    assert len(seen) >= 8 and all(a is marker for a in seen)


def test_cuda_graph_replay_of_the_staged_iteration_is_bit_identical():
    rng = np.random.default_rng(11)
    B = 257
    y0 = cu((1 + rng.uniform(size=(B, 2))).astype(np.float32))
    t_eval = torch.linspace(0, 5, 20, device=DEV).expand(B, 20)
    field = to.fields.LotkaVolterra()
    sols = []
    for use_graph in (False, True):
        term = to.ODETerm(lambda t, y: field(t, y))
        solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
        solver.use_cuda_graph = use_graph
        with torch.no_grad():
            sols.append(solver.solve(to.InitialValueProblem(y0, t_eval=t_eval)))
    a, b = sols
    assert torch.equal(a.ys, b.ys) and torch.equal(a.stats["n_steps"], b.stats["n_steps"])
    assert torch.equal(a.stats["n_accepted"], b.stats["n_accepted"]) and torch.equal(a.status, b.status)
    assert a.stats["n_f_evals"].tolist() == b.stats["n_f_evals"].tolist()


def test_time_inversion_round_trip():
    """adjoint_test.py:322-347: integrate forwards, then backwards from the end state."""
    B = 64
    g = torch.Generator().manual_seed(5)
    y0 = (1 + torch.rand(B, 2, generator=g, dtype=torch.float64)).to(DEV)
    t0 = torch.zeros(B, device=DEV, dtype=torch.float64)
    t1 = torch.full((B,), 3.0, device=DEV, dtype=torch.float64)
    field = to.fields.LotkaVolterra()
    for ctrl_cls, extra in ((to.IntegralController, ()), (to.PIDController, (0.2, 0.5, 0.0))):
        term = to.ODETerm(field)
        solver = to.AutoDiffAdjoint(to.Tsit5(term), ctrl_cls(1e-10, 1e-10, *extra, term=term))
        with torch.no_grad():
            fwd = solver.solve(to.InitialValueProblem(y0, t0, t1))
            bwd = solver.solve(to.InitialValueProblem(fwd.ys[:, 0], t1, t0))
        assert (fwd.status == 0).all() and (bwd.status == 0).all()
        assert torch.allclose(bwd.ys[:, 0], y0, rtol=1e-6)


def test_tsit5_dense_output_recovers_a_quartic():
    """interpolation_test.py:105-127: y' = 4t^3 - 2t has the quartic solution t^4 - t^2 + c."""
    B = 8
    c = torch.linspace(0.5, 2.0, B, device=DEV, dtype=torch.float64)
    t_eval = torch.linspace(0.0, 1.5, 31, device=DEV, dtype=torch.float64).repeat(B, 1)
    term = to.ODETerm(lambda t, y: (4 * t**3 - 2 * t)[:, None].expand_as(y))
    solver = to.AutoDiffAdjoint(to.Tsit5(term), to.IntegralController(1e-3, 1e-3, term=term))
    with torch.no_grad():
        sol = solver.solve(to.InitialValueProblem(c[:, None].clone(), t_eval=t_eval))
    want = (t_eval**4 - t_eval**2 + c[:, None])[:, :, None]
    assert (sol.stats["n_initialized"] == 31).all()
    assert torch.allclose(sol.ys, want, atol=5e-4)


def test_cuda_graph_plan_is_reused_across_solves_with_new_inputs():
    B = 300
    field = to.fields.LotkaVolterra()
    term = to.ODETerm(lambda t, y: field(t, y))
    graph_solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
    graph_solver.use_cuda_graph = True
    plain_solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
    t_eval = torch.linspace(0, 4, 12, device=DEV).expand(B, 12)
    previous = None
    for seed in range(3):
        g = torch.Generator().manual_seed(seed)
        y0 = (1 + torch.rand(B, 2, generator=g)).to(DEV)
        with torch.no_grad():
            a = graph_solver.solve(to.InitialValueProblem(y0, t_eval=t_eval))
            b = plain_solver.solve(to.InitialValueProblem(y0, t_eval=t_eval))
        assert torch.equal(a.ys, b.ys) and torch.equal(a.stats["n_steps"], b.stats["n_steps"])
        assert a.stats["n_f_evals"].tolist() == b.stats["n_f_evals"].tolist()
        if previous is not None:  # earlier solutions are not clobbered by the reused buffers
            assert torch.equal(previous[0].ys, previous[1])
        previous = (a, a.ys.clone())
    assert len(graph_solver._plans) == 1 and graph_solver.last_run["route"] == "staged+graph"


@pytest.mark.parametrize("method_cls", [to.Heun, to.Euler, to.Dopri5, to.Tsit5])
def test_fixed_step_methods_follow_the_analytic_solution(method_cls):
    """fixed_step_controller_test.py:10-19 incl. Heun; Euler gets a finer step."""
    B = 3
    y0 = torch.full((B, 2), 2.0, device=DEV, dtype=torch.float64)
    t_eval = torch.linspace(0, 1.5, 7, device=DEV, dtype=torch.float64).repeat(B, 1)
    term = to.ODETerm(lambda t, y: -y)
    solver = to.AutoDiffAdjoint(method_cls(term), to.FixedStepController())
    dt = 0.002 if method_cls is to.Euler else 0.05
    with torch.no_grad():
        sol = solver.solve(to.InitialValueProblem(y0, t_eval=t_eval),
                           dt0=torch.full((B,), dt, device=DEV, dtype=torch.float64))
    assert (sol.status == 0).all() and (sol.stats["n_initialized"] == 7).all()
    assert torch.allclose(sol.ys[:, :, 0], 2.0 * torch.exp(-t_eval), rtol=1e-2)


@pytest.mark.parametrize("time_dtype,data_dtype", [(torch.float32, torch.float64), (torch.float64, torch.float32)])
@pytest.mark.parametrize("method_cls", [to.Heun, to.Dopri5, to.Tsit5])
def test_time_and_data_dtypes_never_mix(method_cls, time_dtype, data_dtype):
    """dtype_stability_test.py:15-42: adaptive Heun / Dopri5 / Tsit5 + PID with different dtypes."""
    B = 4
    y0 = torch.ones(B, 3, device=DEV, dtype=data_dtype)
    t_eval = torch.linspace(0, 1, 5, device=DEV, dtype=time_dtype).repeat(B, 1)

    def f(t, y):
        assert t.dtype == time_dtype and y.dtype == data_dtype
        return -y * t[:, None].to(data_dtype)

    term = to.ODETerm(f)
    solver = to.AutoDiffAdjoint(method_cls(term), to.PIDController(1e-5, 1e-5, 0.2, 0.5, 0.0, term=term))
    with torch.no_grad():
        sol = solver.solve(to.InitialValueProblem(y0, t_eval=t_eval))
    assert sol.ys.dtype == data_dtype and sol.ts.dtype == time_dtype and (sol.status == 0).all()
    want = torch.exp(-0.5 * t_eval.to(data_dtype) ** 2)[:, :, None].expand(-1, -1, 3)
    assert torch.allclose(sol.ys, want, rtol=2e-3)


def test_solve_ivp_with_registered_method_names():
    y0 = torch.ones(2, 1, device=DEV, dtype=torch.float64)
    t_eval = torch.linspace(0, 1, 4, device=DEV, dtype=torch.float64)
    for name in ("heun", "dopri5", "tsit5"):
        with torch.no_grad():
            sol = to.solve_ivp(lambda t, y: -y, y0, t_eval, method=name)
        assert torch.allclose(sol.ys[:, :, 0], torch.exp(-t_eval).expand(2, -1), rtol=1e-4), name


def _lv_solver(staged):
    field = to.fields.LotkaVolterra()
    term = to.ODETerm((lambda t, y: field(t, y)) if staged else field)
    return to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))


@pytest.mark.parametrize("staged", [False, True])
def test_edge_cases_empty_batch_zero_span_and_strided_inputs(staged):
    solver = _lv_solver(staged)
    with torch.no_grad():
        # empty batch
        sol = solver.solve(to.InitialValueProblem(torch.empty(0, 2, device=DEV), torch.empty(0, device=DEV),
                                                  torch.empty(0, device=DEV)))
        assert sol.ys.shape == (0, 1, 2) and sol.stats["n_steps"].shape == (0,)
        # t_start == t_end: one masked step with dt = 0 (time_direction = -1)
        y0 = torch.tensor([[1.0, 2.0], [1.5, 0.5]], device=DEV)
        t = torch.tensor([0.0, 1.0], device=DEV)
        sol = solver.solve(to.InitialValueProblem(y0, t, t.clone()))
        assert sol.stats["n_steps"].tolist() == [1, 1] and sol.stats["n_accepted"].tolist() == [1, 1]
        assert torch.equal(sol.ys[:, 0], y0) and (sol.status == 0).all()
        # non-contiguous y0 / t_eval views, a single sample, an expanded scalar time
        big = (1 + torch.rand(2, 7, device=DEV))
        y0_view = big.T[:5]                       # (5, 2) with strides (1, 7)
        te_big = torch.linspace(0, 2, 12, device=DEV).repeat(5, 1)
        te_view = te_big[:, ::2]                  # every other column
        a = solver.solve(to.InitialValueProblem(y0_view, t_eval=te_view))
        b = solver.solve(to.InitialValueProblem(y0_view.contiguous(), t_eval=te_view.contiguous()))
        assert torch.equal(a.ys, b.ys) and torch.equal(a.stats["n_steps"], b.stats["n_steps"])
        one = solver.solve(to.InitialValueProblem(y0_view[:1], torch.zeros(1, device=DEV),
                                                  torch.tensor(2.0, device=DEV).expand(1)))
        assert one.ys.shape == (1, 1, 2) and (one.status == 0).all() and torch.isfinite(one.ys).all()


def test_fast_scalar_math_is_bit_identical():
    """The fused kernel's branch-free division / log2 / exp2 / controller must return the bits
    of the checked functions wherever they leave their range flag set (erk_math.cuh)."""
    lib = _cabi.lib()
    ctrl = to.PIDController(1e-8, 1e-8, 0.2, 0.5, 0.1).to_cabi(5, torch.float64)
    counts = torch.zeros(8, dtype=torch.int64, device="cuda")
    n = 1 << 26
    _cabi.check(lib.tode_selftest_fast_math(n, 20241017, C.byref(ctrl), counts.data_ptr(),
                                            _launch.stream_ptr(counts.device)), "selftest")
    got = counts.tolist()
    assert got[:4] == [0, 0, 0, 0], f"mismatches (div, log2, exp2, controller): {got[:4]}"
    # the fast path must actually have been exercised: most moderate-range operands keep the flag
    assert got[4] > n // 4 and got[5] > n // 4 and got[6] > n // 8 and got[7] > n // 16, got


@pytest.mark.parametrize("kind", ["fused", "fused_teval", "opaque"])
def test_solve_from_host_equals_the_device_solve(kind):
    """Chunked, stream-pipelined solve of host-resident problems: same bits as one device solve
    (no sample fails here, so the per-chunk "failure stops the batch" scope does not show)."""
    from torchode_b200.fields import LotkaVolterra, VanDerPol

    B = 1000
    g = torch.Generator().manual_seed(5)
    if kind == "fused":
        y0 = (torch.rand(B, 2, generator=g, dtype=torch.float64) * 4 - 2).pin_memory()
        host = to.InitialValueProblem(y0, torch.zeros(B, dtype=torch.float64).pin_memory(),
                                      torch.full((B,), 2.0, dtype=torch.float64).pin_memory())
        term = to.ODETerm(VanDerPol(10.0))
        solver = to.AutoDiffAdjoint(to.Tsit5(term), to.PIDController(1e-8, 1e-8, 0.2, 0.5, 0.0, term=term))
    else:
        y0 = (1 + torch.rand(B, 2, generator=g)).pin_memory()
        t_eval = torch.linspace(0, 3, 17).expand(B, -1)
        host = to.InitialValueProblem(y0, t_eval=t_eval)
        lv = LotkaVolterra()
        term = to.ODETerm(lv if kind == "fused_teval" else (lambda t, y: lv(t, y)))
        solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
    dev_problem = to.InitialValueProblem(host.y0.cuda(), host.t_start.cuda(), host.t_end.cuda(),
                                         None if host.t_eval is None else host.t_eval.cuda())
    want = solver.solve(dev_problem)
    got = to.solve_from_host(solver, host, "cuda", chunks=3, min_chunk=1)
    again = to.solve_from_host(solver, host, "cuda", chunks=7, min_chunk=1, out=got)  # buffers reused
    assert again.ys.data_ptr() == got.ys.data_ptr()
    for sol in (got, again):
        assert sol.ys.device.type == "cpu" and sol.ys.is_pinned()
        assert bits_equal(sol.ys.numpy(), want.ys.cpu().numpy())
        assert sol.status.tolist() == want.status.tolist()
        for k in ("n_steps", "n_accepted", "n_initialized", "n_f_evals"):
            assert sol.stats[k].tolist() == want.stats[k].tolist(), k


@pytest.mark.parametrize("with_t_eval", [False, True])
def test_solve_from_host_cuts_few_wide_rows_by_bytes(with_t_eval):
    """configs[4] in miniature: 12 rows of a method-of-lines grid (kernel-backed field: the solve drives its loop from
    the host).  By sample count that is one chunk; ``min_chunk_bytes`` cuts it so that the copies of the other
    chunks run under a chunk's solve -- all copy-ins are queued first, each chunk's results leave on its own
    stream.  Same bits as one device solve."""
    from torchode_b200.fields import Heat1D

    B, F = 12, 4096
    g = torch.Generator().manual_seed(3)
    x = torch.linspace(0, 1, F)
    y0 = (torch.sin(torch.pi * x)[None] * (1 + torch.rand(B, 1, generator=g))).pin_memory()
    t_eval = torch.linspace(0, 2e-6, 5).expand(B, -1) if with_t_eval else None
    host = to.InitialValueProblem(y0, torch.zeros(B).pin_memory(), torch.full((B,), 2e-6).pin_memory(), t_eval)
    term = to.ODETerm(Heat1D(1.0 / (x[1] - x[0]).item() ** 2))
    solver = to.AutoDiffAdjoint(to.Tsit5(term), to.IntegralController(1e-6, 1e-4, term=term))
    want = solver.solve(to.InitialValueProblem(y0.cuda(), host.t_start.cuda(), host.t_end.cuda(),
                                               None if t_eval is None else t_eval.cuda()))
    one = to.solve_from_host(solver, host, "cuda")  # 12 samples, 0.2-1 MB moved: not worth a second stream
    assert solver.last_run["chunks"] == 1
    got = to.solve_from_host(solver, host, "cuda", chunks=5, min_chunk_bytes=64 << 10, out=one)
    assert solver.last_run["chunks"] == 5
    for sol in (one, got):
        assert sol.ys.is_pinned() and bits_equal(sol.ys.numpy(), want.ys.cpu().numpy())
        assert sol.status.tolist() == want.status.tolist() and (sol.status == 0).all()
        for k in ("n_steps", "n_accepted", "n_initialized"):
            assert sol.stats[k].tolist() == want.stats[k].tolist(), k


@pytest.mark.parametrize("with_t_eval", [False, True])
def test_fused_kernel_writes_replicas_of_the_gathered_buffers(with_t_eval):
    """tode_solution.peer_*: the multi-GPU "write the all-gather while solving" path, exercised on
    one GPU with two replicas that both live here (on a box they are peer mappings over NVLink)."""
    from torchode_b200.fields import LotkaVolterra

    B, F, T, world, rank = 300, 2, 9, 2, 1
    g = torch.Generator().manual_seed(11)
    y0 = (1 + torch.rand(B, F, generator=g)).cuda()
    t_eval = torch.linspace(0, 2, T).cuda().expand(B, -1) if with_t_eval else None
    prob = to.InitialValueProblem(y0, torch.zeros(B, device="cuda"), torch.full((B,), 2.0, device="cuda"), t_eval)
    term = to.ODETerm(LotkaVolterra())
    solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
    want = solver.solve(prob)
    Tn = T if with_t_eval else 1

    class Replicas:  # stands in for distributed.SymmetricWorkspace
        def __init__(self):
            G = world * B
            self.ys = [torch.full((G, Tn, F), -7.0, device="cuda") for _ in range(world)]
            self.stats = [torch.full((4, G), -7, dtype=torch.long, device="cuda") for _ in range(world)]
            self.glob = [torch.zeros(4, dtype=torch.int32, device="cuda") for _ in range(world)]

        def own_rows(self, b, n_points, f, dtype, rows=None):  # replica `rank` is this rank's own gathered buffer
            lo, hi = rank * B, (rank + 1) * B
            return (self.ys[rank][lo:hi],) + tuple(self.stats[rank][k][lo:hi] for k in range(4))

        def fill(self, sol, b, n_points, f, dtype, rows=None):
            sol.n_peers, sol.peer_row0 = world, rank * B
            for p in range(world):
                remote = p != rank
                sol.peer_ys[p] = self.ys[p].data_ptr() if remote else None
                for k, name in enumerate(("peer_n_steps", "peer_n_accepted", "peer_n_initialized", "peer_status")):
                    getattr(sol, name)[p] = self.stats[p][k].data_ptr() if remote else None
                sol.peer_global[p] = self.glob[p].data_ptr()

    reps = Replicas()
    ctx = solver._fused_launch(prob, term, term.f, None, peers=reps)
    got = solver._fused_finish(ctx)
    assert bits_equal(got.ys.cpu().numpy(), want.ys.cpu().numpy())
    iters = (int(want.stats["n_f_evals"][0]) - 2) // 6
    lo, hi = rank * B, (rank + 1) * B
    for p in range(world):
        assert bits_equal(reps.ys[p][lo:hi].cpu().numpy(), want.ys.cpu().numpy())
        assert bool((reps.ys[p][:lo] == -7).all())  # other shards' rows untouched
        for k, ref in enumerate((want.stats["n_steps"], want.stats["n_accepted"], want.stats["n_initialized"],
                                 want.status)):
            assert reps.stats[p][k][lo:hi].tolist() == ref.tolist()
        assert reps.glob[p].tolist() == [iters, 0, 0, 0]


def test_solve_from_host_edge_cases():
    """Chunks are independent solves: a failing sample stops its own chunk only (the documented
    per-chunk scope), non-monotone t_eval rows fall back to the stage-wise route inside their chunk,
    user dt0 and an empty batch pass through."""
    from torchode_b200.fields import LotkaVolterra

    B, chunks = 96, 3
    g = torch.Generator().manual_seed(21)
    y0 = 1 + torch.rand(B, 2, generator=g)
    y0[70, 0] = float("inf")  # third chunk (rows 64..95) fails at its first step
    t_eval = torch.linspace(0, 3, 11).repeat(B, 1)
    t_eval[5] = t_eval[5].flip(0) * 0 + torch.tensor([0, 2, 1, 3, 4, 5, 6, 7, 8, 9, 10.0]) * 0.3  # row 5 not monotone
    dt0 = torch.full((B,), 1e-3)
    term = to.ODETerm(LotkaVolterra())
    solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
    host = to.InitialValueProblem(y0.pin_memory(), t_eval=t_eval.pin_memory())
    got = to.solve_from_host(solver, host, "cuda", chunks=chunks, min_chunk=1, dt0=dt0)
    n_f = 0
    for c in range(chunks):
        lo, hi = c * 32, (c + 1) * 32
        want = solver.solve(to.InitialValueProblem(y0[lo:hi].cuda(), t_eval=t_eval[lo:hi].cuda()), dt0=dt0[lo:hi].cuda())
        n_init = want.stats["n_initialized"].cpu()
        for b in range(32):  # only initialised evaluation points are defined
            k = int(n_init[b])
            assert bits_equal(got.ys[lo + b, :k].numpy(), want.ys[b, :k].cpu().numpy()), (c, b)
        assert got.status[lo:hi].tolist() == want.status.tolist()
        for key in ("n_steps", "n_accepted", "n_initialized"):
            assert got.stats[key][lo:hi].tolist() == want.stats[key].tolist(), (c, key)
        n_f = max(n_f, int(want.stats["n_f_evals"][0]))
    assert int(got.stats["n_f_evals"][0]) == n_f
    assert int(got.status[70]) != 0 and int((got.status[:64] != 0).sum()) == 0
    empty = to.solve_from_host(solver, to.InitialValueProblem(torch.empty(0, 2), t_eval=torch.empty(0, 11)), "cuda")
    assert empty.ys.shape == (0, 11, 2) and empty.status.shape == (0,)

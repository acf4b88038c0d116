"""The packed-fp32 persistent whole-solve kernel (csrc/erk_fused_f2.cuh: fp32 state and time, two features,
t_eval rows -- BASELINE configs[2]) against the general fused kernel and the oracle, bit for bit.

``TODE_NO_F2=1`` makes ``tode_solve_fused`` take the general kernel (read at every launch), so both kernels run
in one process on the same inputs.  Sizes cover: fewer samples than a warp, ragged tails, and batches several
times the resident grid (every lane is refilled from the queue many times)."""
import os

import numpy as np
import pytest
import torch

import torchode_b200 as to
from oracle import oracle as orc
from torchode_b200 import _cabi
from torchode_b200.fields import LinearDecay, LotkaVolterra, VanDerPol
from torchode_b200.step_size_controllers import max_norm

from helpers import bits_equal

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _solve(field, method_cls, ctrl_fn, y0, t_start, t_end, t_eval, dt0=None, max_steps=None, general=False):
    term = to.ODETerm(field)
    solver = to.AutoDiffAdjoint(method_cls(term), ctrl_fn(term), max_steps=max_steps)
    if general:
        os.environ["TODE_NO_F2"] = "1"
    try:
        with torch.no_grad():
            sol = solver.solve(to.InitialValueProblem(y0, t_start, t_end, t_eval), dt0=dt0)
        torch.cuda.synchronize()
    finally:
        os.environ.pop("TODE_NO_F2", None)
    return sol, solver.last_run


def _same(a, b, what):
    for k in ("n_steps", "n_accepted", "n_initialized"):
        assert torch.equal(a.stats[k], b.stats[k]), f"{what}: {k}"
    assert torch.equal(a.status, b.status), what
    assert torch.equal(a.stats["n_f_evals"], b.stats["n_f_evals"]), what
    n_init = a.stats["n_initialized"].cpu().numpy()
    ya, yb = a.ys.cpu().numpy(), b.ys.cpu().numpy()
    valid = np.arange(ya.shape[1])[None, :, None] < n_init[:, None, None]
    assert bits_equal(np.where(valid, ya, 0), np.where(valid, yb, 0)), f"{what}: ys differ"


def _inputs(field_name, B, T, seed, bidir=False, per_sample_rows=False):
    g = torch.Generator().manual_seed(seed)
    if field_name == "lv":
        field, y0, t1 = LotkaVolterra(), 1 + torch.rand(B, 2, generator=g), 10.0
    elif field_name == "vdp":
        field, y0, t1 = VanDerPol(2.0), torch.rand(B, 2, generator=g) * 4 - 2, 6.0
    else:
        field, y0, t1 = LinearDecay(-0.7), torch.rand(B, 2, generator=g) * 3 + 0.1, 4.0
    t_start = torch.zeros(B)
    t_end = torch.full((B,), t1)
    row = torch.linspace(0, t1, T)
    if per_sample_rows:
        t_eval = (row[None] * (0.5 + 0.5 * torch.rand(B, 1, generator=g))).contiguous()
        t_end = t_eval[:, -1].clone()
    else:
        t_eval = row.expand(B, -1)
    if bidir:  # odd samples run backwards in time over the same points
        t_eval = t_eval.clone()
        t_eval[1::2] = t_eval[1::2].flip(1)
        t_start = t_eval[:, 0].clone()
        t_end = t_eval[:, -1].clone()
    return field, y0.to(DEV), t_start.to(DEV), t_end.to(DEV), t_eval.to(DEV)


CTRLS = {
    "integral": lambda term: to.IntegralController(1e-6, 1e-3, term=term),
    "integral_tight": lambda term: to.IntegralController(1e-7, 1e-6, term=term),
    "pid": lambda term: to.PIDController(1e-6, 1e-4, 0.2, 0.5, 0.0, term=term),
    "pid_d": lambda term: to.PIDController(1e-6, 1e-4, 0.3, 0.4, 0.1, term=term),
    "integral_max": lambda term: to.IntegralController(1e-6, 1e-3, term=term, norm=max_norm),
    "integral_dtlim": lambda term: to.IntegralController(1e-6, 1e-3, term=term, dt_min=1e-4, dt_max=0.3),
}


@pytest.mark.parametrize("ctrl", sorted(CTRLS))
@pytest.mark.parametrize("method", ["dopri5", "tsit5"])
@pytest.mark.parametrize("field_name", ["lv", "vdp", "linear"])
def test_f2_kernel_equals_general_kernel(field_name, method, ctrl):
    B = 4099  # ragged: 128 full warps + 3 lanes
    field, y0, t0, t1, te = _inputs(field_name, B, 37, seed=11)
    cls = {"dopri5": to.Dopri5, "tsit5": to.Tsit5}[method]
    a, run_a = _solve(field, cls, CTRLS[ctrl], y0, t0, t1, te)
    b, run_b = _solve(field, cls, CTRLS[ctrl], y0, t0, t1, te, general=True)
    assert run_a["route"].startswith("fused") and run_b["route"].startswith("fused")
    _same(a, b, f"{field_name} {method} {ctrl}")


@pytest.mark.parametrize("B", [1, 5, 31, 33, 200_003, 1 << 20])
def test_f2_kernel_sizes_and_refill(B):
    field, y0, t0, t1, te = _inputs("lv", B, 100, seed=B)
    a, _ = _solve(field, to.Dopri5, CTRLS["integral"], y0, t0, t1, te)
    b, _ = _solve(field, to.Dopri5, CTRLS["integral"], y0, t0, t1, te, general=True)
    _same(a, b, f"B={B}")


@pytest.mark.parametrize("variant", ["bidir", "rows", "dt0", "max_steps", "t1", "long_row"])
def test_f2_kernel_variants(variant):
    B, T = 3000, 64
    kw = {}
    if variant == "t1":
        T = 1
    if variant == "long_row":
        T = 1500  # more points than the shared-memory copy of a broadcast row holds
    field, y0, t0, t1, te = _inputs("lv", B, T, seed=5, bidir=variant == "bidir", per_sample_rows=variant == "rows")
    if variant == "dt0":
        kw["dt0"] = torch.full((B,), 0.05, device=DEV)
    if variant == "max_steps":
        kw["max_steps"] = 9
    a, _ = _solve(field, to.Tsit5, CTRLS["pid"], y0, t0, t1, te, **kw)
    b, _ = _solve(field, to.Tsit5, CTRLS["pid"], y0, t0, t1, te, general=True, **kw)
    _same(a, b, variant)
    if variant == "max_steps":
        assert int((a.status == 3).sum()) > 0


def test_f2_kernel_failure_replay_and_nonfinite():
    """A sample with a non-finite state fails at iteration 1: the batch is cut there (adjoints.py:186-190)."""
    field, y0, t0, t1, te = _inputs("lv", 2048, 20, seed=3)
    y0[77, 0] = float("inf")
    a, run_a = _solve(field, to.Dopri5, CTRLS["integral"], y0, t0, t1, te)
    b, run_b = _solve(field, to.Dopri5, CTRLS["integral"], y0, t0, t1, te, general=True)
    assert run_a["route"] == run_b["route"] == "fused+replay"
    _same(a, b, "failure replay")
    assert int(a.status[77]) != 0


def test_f2_kernel_nonmonotone_rows_fall_back():
    field, y0, t0, t1, te = _inputs("lv", 512, 16, seed=9)
    te = te.clone()
    te[:, [3, 4]] = te[:, [4, 3]]
    a, run_a = _solve(field, to.Dopri5, CTRLS["integral"], y0, t0, t1, te)
    b, run_b = _solve(field, to.Dopri5, CTRLS["integral"], y0, t0, t1, te, general=True)
    assert not run_a["route"].startswith("fused") and not run_b["route"].startswith("fused")
    _same(a, b, "non-monotone rows")


@pytest.mark.parametrize("field_name,method,ctrl", [("lv", "dopri5", "integral"), ("vdp", "tsit5", "pid"),
                                                    ("linear", "dopri5", "integral_max")])
def test_f2_kernel_equals_oracle(field_name, method, ctrl):
    B = 1031
    field, y0, t0, t1, te = _inputs(field_name, B, 50, seed=21)
    cls = {"dopri5": to.Dopri5, "tsit5": to.Tsit5}[method]
    sol, _ = _solve(field, cls, CTRLS[ctrl], y0, t0, t1, te)
    term = to.ODETerm(field)
    m, c = cls(term), CTRLS[ctrl](term)
    ref = orc.solve_builtin(field.field_id, field.params(), m.to_cabi(), c.to_cabi(5, torch.float32),
                            y0.cpu().numpy(), t0.cpu().numpy(), t1.cpu().numpy(), te.cpu().numpy())
    assert sol.stats["n_steps"].cpu().numpy().tolist() == ref["n_steps"].tolist()
    assert sol.stats["n_accepted"].cpu().numpy().tolist() == ref["n_accepted"].tolist()
    assert sol.stats["n_initialized"].cpu().numpy().tolist() == ref["n_initialized"].tolist()
    assert int(sol.stats["n_f_evals"][0]) == int(ref["n_f_evals"])
    assert bits_equal(sol.ys.cpu().numpy(), ref["ys"])


def test_f32_fast_division_and_sqrt_are_bit_identical_to_the_ieee_instructions():
    """div_fast / div_fast2 / rcp_refined2 / sqrt_fast (erk_fused_f2.cuh, also used by heat_step.cu) against
    div.rn.f32 / sqrt.rn.f32 wherever their range flag is set: 2^28 operand triples -- a third any bit pattern,
    a third in the solver's range, a third around the limits of the flags (2^-60, 2^60, 2^-101, zeros,
    subnormals, near-overflow)."""
    import ctypes as C

    from torchode_b200 import _launch

    counts = torch.zeros(6, dtype=torch.int64, device=DEV)
    n = 1 << 28
    _cabi.check(_cabi.lib().tode_selftest_fast_math_f32(n, 20261017, counts.data_ptr(),
                                                        _launch.stream_ptr(counts.device)), "selftest f32")
    got = counts.tolist()
    assert got[:3] == [0, 0, 0], f"mismatches (division, division by sqrt(2), square root): {got[:3]}"
    # the flags must leave the fast path open for the solver's range (a third of the operands) and more
    assert got[3] > n // 4 and got[4] > n // 3 and got[5] > n // 2, got

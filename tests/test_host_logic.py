"""Host-side logic on CPU: the generic plug-in route of AutoDiffAdjoint.solve (the reference's
operator API: arbitrary SingleStepMethod / StepSizeController objects), parameter packing,
input validation, and the loud refusal to run built-in components without CUDA.

The scenarios mirror what the reference's own loop tests pin (tests/adjoint_test.py:26-319):
exact n_steps / n_accepted / n_initialized lists, rejected steps retrying from the same t,
several evaluation points inside one step, max_steps, status codes stopping the batch."""
import math

import pytest
import torch

import torchode_b200 as to
from torchode_b200 import _cabi
from torchode_b200.single_step_methods import SingleStepMethod, StepResult
from torchode_b200.step_size_controllers import StepSizeController, max_norm, rms_norm


def exact(t):
    """closed-form solution used by the scripted step method: y(t) = sin(t) + t^2"""
    return (torch.sin(t) + t * t)[..., None]


class ExactInterp:
    def evaluate(self, t, idx):
        return exact(t)


class ExactStep(SingleStepMethod):
    """'Steps' by evaluating the closed form at t + dt; error and status are scripted."""

    def __init__(self, status=0, error=None):
        super().__init__()
        self.term = to.ODETerm(lambda t, y: (_ for _ in ()).throw(AssertionError("f must not be called")))
        self.status, self.error = status, error
        self.calls = []

    def init(self, term, problem, f0, *, stats, args):
        return None

    def step(self, term, running, y, t, dt, state, *, stats, args):
        self.calls.append((t.clone(), dt.clone(), running.clone()))
        y1 = exact(t + dt)
        err = torch.zeros_like(y1) if self.error is None else self.error(t)
        status = self.status(t + dt) if callable(self.status) else torch.full_like(t, self.status, dtype=torch.long)
        return StepResult(y1, err), None, state, status

    def merge_states(self, accept, current, previous):
        return current

    def build_interpolation(self, data):
        return ExactInterp()

    def convergence_order(self):
        return 3


class Scripted(StepSizeController):
    """dt0, then per-iteration scripted (dt_next, accept, status): constants or lists consumed in order."""

    def __init__(self, dt0, dt, accept=True, status=0):
        super().__init__()
        self.dt0, self.script = dt0, dict(dt=dt, accept=accept, status=status)
        self.i = 0

    def _get(self, key, shape, dtype):
        v = self.script[key]
        if isinstance(v, list) and len(v) and isinstance(v[0], (list, tuple)):
            v = v[min(self.i, len(v) - 1)]
        return torch.as_tensor(v, dtype=dtype).expand(shape)

    def init(self, term, problem, order, dt0, *, stats, args):
        shape = (problem.batch_size,)
        return torch.as_tensor(self.dt0, dtype=problem.time_dtype).expand(shape), {}, None

    def adapt_step_size(self, t0, dt, y0, step_result, state, stats):
        out = (self._get("accept", dt.shape, torch.bool), self._get("dt", dt.shape, dt.dtype), state,
               self._get("status", dt.shape, torch.long))
        self.i += 1
        return out

    def merge_states(self, running, current, previous):
        return current


def problem_of(t_eval=None, t_start=None, t_end=None):
    if t_eval is not None:
        t_eval = torch.tensor(t_eval)
        t_start = t_eval[:, 0] if t_start is None else torch.tensor(t_start)
        t_end = t_eval[:, -1] if t_end is None else torch.tensor(t_end)
    else:
        t_start, t_end = torch.tensor(t_start), torch.tensor(t_end)
    return to.InitialValueProblem(exact(t_start), t_start, t_end, t_eval)


def test_evaluates_at_every_evaluation_point():
    problem = problem_of([[0.0, 2.0], [2.0, 4.0], [0.5, 2.5]])
    sol = to.AutoDiffAdjoint(ExactStep(), Scripted(0.3, 0.3)).solve(problem)
    assert sol.status.tolist() == [0, 0, 0]
    assert sol.stats["n_steps"].tolist() == [7, 7, 7]
    assert sol.stats["n_accepted"].tolist() == [7, 7, 7]
    assert sol.stats["n_initialized"].tolist() == [2, 2, 2]
    assert sol.ts is problem.t_eval
    assert torch.equal(sol.ys, exact(problem.t_eval))


def test_samples_step_independently():
    problem = problem_of([[0.0, 0.15, 1.0], [1.0, 1.9, 2.0]])
    sol = to.AutoDiffAdjoint(ExactStep(), Scripted([0.1, 0.3], [0.5, 0.125])).solve(problem)
    assert sol.stats["n_steps"].tolist() == [3, 7]
    assert sol.stats["n_accepted"].tolist() == [3, 7]
    assert sol.stats["n_initialized"].tolist() == [3, 3]
    assert torch.equal(sol.ys, exact(problem.t_eval))


def test_several_evaluation_points_inside_one_step():
    problem = problem_of([[0.0, 0.25, 0.33, 1.0]])
    sol = to.AutoDiffAdjoint(ExactStep(), Scripted(0.1, 0.5)).solve(problem)
    assert sol.stats["n_steps"].tolist() == [3]
    assert sol.stats["n_initialized"].tolist() == [4]
    assert torch.equal(sol.ys, exact(problem.t_eval))


def test_no_t_eval_reports_t_end():
    problem = problem_of(t_start=[0.0, 5.0, 2.0], t_end=[10.0, 9.0, 4.5])
    sol = to.AutoDiffAdjoint(ExactStep(), Scripted([0.5001, 0.2501, 1.0], [0.5001, 0.2501, 1.0])).solve(problem)
    assert sol.stats["n_steps"].tolist() == [20, 16, 3]
    assert sol.stats["n_initialized"].tolist() == [1, 1, 1]
    assert sol.ts[:, 0].tolist() == [10.0, 9.0, 4.5]
    assert sol.ys.shape == (3, 1, 1)
    assert torch.allclose(sol.ys[:, 0], exact(problem.t_end))


def test_rejected_steps_retry_from_the_same_time():
    problem = problem_of([[0.0, 1.0], [1.0, 2.0]])
    method = ExactStep()
    accept = [[True, True], [False, True], [True, False], [True, True]]
    sol = to.AutoDiffAdjoint(method, Scripted(0.4, 0.4, accept=accept)).solve(problem)
    t_seen = torch.stack([c[0] for c in method.calls])
    assert torch.allclose(t_seen[:4], torch.tensor([[0.0, 1.0], [0.4, 1.4], [0.4, 1.8], [0.8, 1.8]]))
    assert sol.stats["n_steps"].tolist() == [4, 4]
    assert sol.stats["n_accepted"].tolist() == [3, 3]
    assert torch.equal(sol.ys, exact(problem.t_eval))


def test_f_is_never_evaluated_outside_the_time_domain():
    problem = problem_of([[0.0, 1.0], [3.0, 2.0]])  # second sample runs backwards
    method = ExactStep()
    to.AutoDiffAdjoint(method, Scripted([0.7, -0.7], [0.7, -0.7])).solve(problem)
    for t, dt, _ in method.calls:
        t1 = t + dt
        assert (t1[0] <= 1.0 + 1e-6) and (t1[1] >= 2.0 - 1e-6)


def test_max_steps_sets_status_for_unfinished_samples_only():
    t_eval = torch.tensor([[1.0, 4.9], [2.0, 13.0]])
    problem = to.InitialValueProblem(torch.zeros(2, 1), t_eval[:, 0], t_eval[:, -1], t_eval)
    solver = to.AutoDiffAdjoint(ExactStep(), to.FixedStepController(), max_steps=7)
    sol = solver.solve(problem, dt0=torch.ones(2))
    assert sol.status.tolist() == [0, to.Status.REACHED_MAX_STEPS.value]
    assert sol.stats["n_steps"].tolist() == [4, 7]
    assert sol.stats["n_initialized"].tolist() == [2, 1]


@pytest.mark.parametrize("who", ["method", "controller"])
def test_non_success_status_stops_the_whole_batch(who):
    problem = problem_of([[0.0, 1.0], [0.0, 1.0]])
    bad = [[0, 0], [0, to.Status.GENERAL_ERROR.value]]
    if who == "method":
        n = {"i": 0}

        def status(t):
            s = torch.tensor(bad[min(n["i"], 1)])
            n["i"] += 1
            return s
        solver = to.AutoDiffAdjoint(ExactStep(status=status), Scripted(0.1, 0.1))
    else:
        solver = to.AutoDiffAdjoint(ExactStep(), Scripted(0.1, 0.1, status=bad))
    sol = solver.solve(problem)
    assert sol.status.tolist() == [0, 1]
    assert sol.stats["n_steps"].tolist() == [2, 2]
    assert sol.stats["n_initialized"].tolist() == [1, 1]


def test_finished_samples_keep_their_step_size_clamped_to_zero():
    problem = problem_of([[0.0, 0.2], [0.0, 1.0]])
    method = ExactStep()
    to.AutoDiffAdjoint(method, Scripted(0.2, 0.2)).solve(problem)
    dts = torch.stack([c[1] for c in method.calls])
    assert dts[0].tolist() == pytest.approx([0.2, 0.2])
    assert (dts[1:, 0] == 0).all()  # the finished sample is stepped with dt == 0 (masked)


def test_builtin_components_refuse_cpu_tensors_loudly():
    term = to.ODETerm(lambda t, y: -y)
    solver = to.AutoDiffAdjoint(to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term))
    with pytest.raises(RuntimeError, match="no CPU"):
        solver.solve(to.InitialValueProblem(torch.ones(2, 1), torch.zeros(2), torch.ones(2)))
    with pytest.raises(RuntimeError, match="no CPU"):
        to.solve_ivp(lambda t, y: -y, torch.ones(2, 1), torch.linspace(0, 1, 3))


def test_problem_validation_and_properties():
    with pytest.raises(AssertionError):
        to.InitialValueProblem(torch.ones(3), torch.zeros(3), torch.ones(3))  # y0 must be 2-d
    with pytest.raises(AssertionError):
        to.InitialValueProblem(torch.ones(3, 1), torch.zeros(3), torch.ones(3, dtype=torch.float64))
    p = to.InitialValueProblem(torch.ones(3, 2), t_eval=torch.tensor([[0.0, 1.0], [1.0, 0.0], [2.0, 2.0]]))
    assert p.time_direction.tolist() == [1, -1, -1]  # equal times count as backwards
    assert (p.batch_size, p.n_features, p.n_evaluation_points) == (3, 2, 2)
    assert p.data_dtype == torch.float32 and p.time_dtype == torch.float32


def test_controller_packing_matches_the_reference_conventions():
    c = to.PIDController(1e-6, 1e-3, 0.2, 0.5, 0.1, dt_min=1e-4, norm=max_norm).to_cabi(5, torch.float64, 11)
    assert c.pid == 1 and c.norm == _cabi.NORM_MAX and c.has_dt_min == 1 and c.has_dt_max == 0
    assert c.atol == float(torch.tensor(1e-6)) and c.rtol == float(torch.tensor(1e-3))  # fp32-rounded
    assert c.exp_ratio == -(0.5 / 5 + 0.2 / 5 + 0.1 / 5) and c.exp_prev == 0.2 / 5 + 2 * (0.1 / 5)
    assert c.exp_prev2 == -(0.1 / 5) and c.max_steps == 11 and c.almost_zero == 1e-38
    i = to.IntegralController(1e-6, 1e-3).to_cabi(5, torch.float32)
    assert i.pid == 0 and i.norm == _cabi.NORM_RMS and i.exp_ratio == -(1.0 / 5) and i.max_steps == -1
    assert not to.IntegralController(1e-6, 1e-3, norm=lambda y: y.abs().sum(1)).fusable()


def test_term_counts_evaluations_and_passes_args():
    seen = []
    term = to.ODETerm(lambda t, y, a: seen.append(a) or y, with_args=True)
    stats = {}
    p = to.InitialValueProblem(torch.ones(4, 1), torch.zeros(4), torch.ones(4))
    term.init(p, stats)
    marker = object()
    term.vf(p.t_start, p.y0, stats, marker)
    term.vf(p.t_start, p.y0, stats, marker)
    assert stats["n_f_evals"].tolist() == [2, 2, 2, 2] and stats["n_f_evals"].device.type == "cpu"
    assert seen == [marker, marker]
    quiet = to.ODETerm(lambda t, y: y, with_stats=False)
    stats = {}
    quiet.init(p, stats)
    assert "n_f_evals" not in stats


def test_pow_tables_are_the_generators_output_and_identical_in_both_trees(tmp_path):
    """csrc/pow_tables.h (kernels) and oracle/pow_tables.h (CPU oracle) must hold the same numbers,
    and those must be what scripts/gen_pow_tables.py computes (200-bit mpmath, rounded to nearest)."""
    import importlib.util
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def numbers(path):
        text = open(path).read()
        return re.findall(r"0x[0-9a-f]{16}ULL", text)

    a = numbers(os.path.join(root, "torchode_b200", "csrc", "pow_tables.h"))
    b = numbers(os.path.join(root, "oracle", "pow_tables.h"))
    assert a == b and len(a) == 1 + 6 + 5 + 4 * 128
    mpmath = pytest.importorskip("mpmath")  # noqa: F841
    spec = importlib.util.spec_from_file_location("gen_pow_tables", os.path.join(root, "scripts", "gen_pow_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    out = tmp_path / "pow_tables.h"
    gen.emit(str(out), "X", *gen.tables())
    assert numbers(str(out)) == a


def test_route_predicates_of_the_fused_step_and_stage_routes():
    """Which problems the step-fused (fields.Heat1D) route takes -- host logic only, no kernel runs."""
    from torchode_b200.adjoints import plain_term_of
    from torchode_b200.fields import Heat1D

    term = to.ODETerm(Heat1D(25.0))
    solver = to.AutoDiffAdjoint(to.Tsit5(term), to.IntegralController(1e-6, 1e-3, term=term))
    y0 = torch.zeros(3, 4096)
    t0, t1 = torch.zeros(3), torch.ones(3)
    plain = to.InitialValueProblem(y0, t0, t1)
    assert solver._step_fusable(plain, term, None, None)
    assert not solver._step_fusable(plain, term, ("args",), None)                 # f(t, y, args)
    assert not solver._step_fusable(plain, term, None, object())                  # recorded (autograd) solve
    with_t_eval = to.InitialValueProblem(y0, t0, t1, torch.linspace(0, 1, 5).expand(3, -1))
    assert solver._step_fusable(with_t_eval, term, None, None)                    # dense output in cursor mode
    assert not solver._step_fusable(with_t_eval, term, None, None, general=True)  # scan-all mask mode: stage-wise
    ragged = to.InitialValueProblem(torch.zeros(3, 4098 + 1), t0, t1)
    assert not solver._step_fusable(ragged, term, None, None)                     # rows of whole 16-byte vectors only
    tiny = to.InitialValueProblem(torch.zeros(3, 4), t0, t1)
    assert not solver._step_fusable(tiny, term, None, None)
    f64 = to.InitialValueProblem(torch.zeros(3, 6, dtype=torch.float64), t0.double(), t1.double())
    assert solver._step_fusable(f64, term, None, None)                            # 2-element vectors in fp64
    class CountingTerm(to.ODETerm):  # a term with its own protocol must see every call of f
        pass

    assert not solver._step_fusable(plain, CountingTerm(Heat1D(25.0)), None, None)
    other = to.ODETerm(lambda t, y: -y)
    assert not solver._step_fusable(plain, other, None, None)
    solver.use_step_fusion = False
    assert not solver._step_fusable(plain, term, None, None)
    assert plain_term_of(to.ODETerm(lambda t, y: y)) and not plain_term_of(to.ODETerm(lambda t, y, a: y, with_args=True))


def test_new_entry_points_report_argument_errors_without_a_gpu():
    import ctypes

    from torchode_b200 import _cabi

    lib = _cabi.lib()
    tab, ctrl, st = _cabi.Tableau(), _cabi.Controller(), _cabi.State()
    tab.n_stages = 7
    assert lib.tode_heat_step(ctypes.byref(tab), ctypes.byref(ctrl), ctypes.byref(st), 1.0, None, None, None, None) == -1
    assert lib.tode_mlp_tanh256_stage_forward(ctypes.byref(tab), 1, ctypes.byref(st), _cabi.KPtrs(), None, None, None,
                                              None, 3, None) == -1
    # split-mode scratch: two partials per (sample, chunk) for the initial step + the finish's records
    B, F = 64, 1 << 20
    chunks_f32, chunks_f64 = F // 4096, F // 2048
    assert lib.tode_scratch_elems(B, F) >= 2 * B + 2 * B * max(chunks_f32, chunks_f64) + 8 * B


def test_solve_from_host_chunk_count():
    """host_pipeline.chunk_count: chunks are worth a stream by samples OR by bytes moved, never more than asked for
    or than there are samples."""
    from torchode_b200.host_pipeline import chunk_count

    mb = 1 << 20
    # configs[1]: 2^20 samples x 2 fp64, no t_eval: 32 MB moved, 256 chunks' worth of samples -> the 8 asked for
    assert chunk_count(1 << 20, 2, 0, 8, 8, 4096, 128 * mb) == 8
    # a small batch of narrow states is launch-bound: one chunk
    assert chunk_count(1000, 2, 17, 4, 8, 4096, 128 * mb) == 1
    # configs[4]: 64 rows x 2^20 fp32, 536 MB moved -> cut by bytes into 4
    assert chunk_count(64, 1 << 20, 0, 4, 8, 4096, 128 * mb) == 4
    assert chunk_count(64, 1 << 20, 0, 4, 2, 4096, 128 * mb) == 2  # never more than asked for
    assert chunk_count(3, 1 << 26, 0, 4, 8, 4096, 128 * mb) == 3   # never more than samples
    assert chunk_count(0, 2, 0, 8, 8, 4096, 128 * mb) == 1
    assert chunk_count(12, 4096, 5, 4, 5, 4096, 64 << 10) == 5      # the explicit threshold of the GPU test

"""Step-size controllers (API of torchode/step_size_controllers.py).

``IntegralController`` and ``PIDController`` keep the reference's constructor signature
and plug-in protocol (``init`` / ``adapt_step_size`` / ``merge_states``).  Inside
``AutoDiffAdjoint.solve`` they are not called per iteration: their parameters are packed
into a ``tode_controller`` and the error-norm / accept / dt-update arithmetic runs fused
in the finish kernel.  The protocol methods remain for stand-alone use and are thin
wrappers over the same CUDA code (``tode_adapt_step_size``, ``tode_init_step_a/_b``).
"""
import ctypes as C
from math import sqrt
from typing import Any, Callable, Dict, Generic, NamedTuple, Optional, Tuple, TypeVar

import torch
import torch.nn as nn

from . import _cabi, _launch, status_codes
from .problems import InitialValueProblem
from .single_step_methods import StepResult
from .terms import ODETerm

ControllerState = TypeVar("ControllerState")


class StepSizeController(nn.Module, Generic[ControllerState]):
    """Plug-in protocol the solve loop drives (step_size_controllers.py:16-113).

    ``init(term, problem, method_order, dt0, *, stats, args) -> (dt, state, f0 | None)``;
    ``adapt_step_size(t0, dt, y0, step_result, state, stats) -> (accept, dt_next, state,
    status | None)``; ``merge_states(running, current, previous) -> state``.
    """

    def init(self, term, problem, method_order, dt0, *, stats, args):
        raise NotImplementedError()

    def adapt_step_size(self, t0, dt, y0, step_result, state, stats):
        raise NotImplementedError()

    def merge_states(self, running, current, previous):
        raise NotImplementedError()


class FixedStepState(NamedTuple):
    accept_all: torch.Tensor
    dt0: torch.Tensor


class FixedStepController(StepSizeController[FixedStepState]):
    """Accept everything, keep ``dt0`` (step_size_controllers.py:121-167)."""

    def init(self, term, problem, method_order, dt0, *, stats, args):
        assert dt0 is not None, "Fixed step size solving requires you to configure dt0"
        everyone = torch.ones(problem.batch_size, device=problem.device, dtype=torch.bool)
        return dt0, FixedStepState(everyone, dt0), None

    def adapt_step_size(self, t0, dt, y0, step_result, state, stats):
        return state.accept_all, state.dt0, state, None

    def merge_states(self, running, current, previous):
        return current


def rms_norm(y: torch.Tensor) -> torch.Tensor:
    """Root-mean-square norm over features (Hairer I, eq. II.4.11)."""
    return torch.linalg.vector_norm(y / sqrt(y.shape[1]), ord=2, dim=1)


def max_norm(y: torch.Tensor) -> torch.Tensor:
    return torch.linalg.vector_norm(y, ord=torch.inf, dim=1)


_NORM_IDS = {rms_norm: _cabi.NORM_RMS, max_norm: _cabi.NORM_MAX}


def _almost_zero(dtype: torch.dtype) -> float:
    # lower bound of the error ratio (step_size_controllers.py:251-256)
    return 1e-5 if dtype == torch.float16 else 1e-38


class _AdaptiveState:
    """Controller state: static limits plus (PID only) the last two accepted error ratios."""

    def __init__(self, method_order, almost_zero, dt_min=None, dt_max=None,
                 prev_error_ratio=None, prev_prev_error_ratio=None):
        self.method_order = method_order
        self.almost_zero = almost_zero
        self.dt_min, self.dt_max = dt_min, dt_max
        self.prev_error_ratio = prev_error_ratio
        self.prev_prev_error_ratio = prev_prev_error_ratio

    def update_error_ratios(self, prev_error_ratio, prev_prev_error_ratio):
        if self.prev_error_ratio is None:
            return self
        return type(self)(self.method_order, self.almost_zero, self.dt_min, self.dt_max,
                          prev_error_ratio, prev_prev_error_ratio)

    def __repr__(self):
        return (f"{type(self).__name__}(method_order={self.method_order}, "
                f"prev_error_ratio={self.prev_error_ratio}, "
                f"prev_prev_error_ratio={self.prev_prev_error_ratio}, "
                f"almost_zero={self.almost_zero}, dt_min={self.dt_min}, dt_max={self.dt_max})")


class IntegralState(_AdaptiveState):
    pass


class PIDState(_AdaptiveState):
    pass


class _AdaptiveController(StepSizeController[_AdaptiveState]):
    """Shared machinery of the I and PID controllers."""

    _pid = False

    def __init__(self, atol: float, rtol: float, *, term: Optional[ODETerm] = None,
                 norm: Callable[[torch.Tensor], torch.Tensor] = rms_norm,
                 dt_min: Optional[float] = None, dt_max: Optional[float] = None,
                 safety: float = 0.9, factor_min: float = 0.2, factor_max: float = 10.0):
        super().__init__()
        # fp32 buffers, like the reference (its effective tolerances are the fp32-rounded
        # values even in fp64 solves, step_size_controllers.py:278-279)
        self.register_buffer("atol", torch.tensor(atol))
        self.register_buffer("rtol", torch.tensor(rtol))
        self.term = term
        self.norm = norm
        self.dt_min, self.dt_max = dt_min, dt_max
        self.safety, self.factor_min, self.factor_max = safety, factor_min, factor_max

    # ---- packing for the kernels -------------------------------------------------
    def _exponents(self, order: int) -> Tuple[float, float, float]:
        raise NotImplementedError()

    def fusable(self) -> bool:
        """Can the finish kernel evaluate this controller (built-in norm)?"""
        return self.norm in _NORM_IDS

    def to_cabi(self, order: int, data_dtype: torch.dtype, max_steps: Optional[int] = None):
        c = _cabi.Controller()
        c.norm = _NORM_IDS[self.norm]
        c.pid = int(self._pid)
        c.has_dt_min, c.has_dt_max = int(self.dt_min is not None), int(self.dt_max is not None)
        c.atol, c.rtol = float(self.atol), float(self.rtol)  # fp32-rounded values
        c.safety, c.factor_min, c.factor_max = self.safety, self.factor_min, self.factor_max
        c.exp_ratio, c.exp_prev, c.exp_prev2 = self._exponents(order)
        c.dt_min = 0.0 if self.dt_min is None else float(self.dt_min)
        c.dt_max = 0.0 if self.dt_max is None else float(self.dt_max)
        c.almost_zero = _almost_zero(data_dtype)
        c.max_steps = -1 if max_steps is None else int(max_steps)
        c.iter_cap = 0
        return c

    # ---- plug-in protocol ---------------------------------------------------------
    def _new_state(self, method_order, problem, dt_min, dt_max):
        az = torch.tensor(_almost_zero(problem.data_dtype), dtype=problem.data_dtype,
                          device=problem.device)
        if self._pid:
            ones = torch.ones(problem.batch_size, dtype=problem.data_dtype, device=problem.device)
            return PIDState(method_order, az, dt_min, dt_max, ones, ones)
        return IntegralState(method_order, az, dt_min, dt_max)

    def initial_state(self, method_order, problem, dt_min, dt_max):
        return self._new_state(method_order, problem, dt_min, dt_max)

    def init(self, term, problem: InitialValueProblem, method_order: int, dt0, *,
             stats: Dict[str, Any], args: Any):
        f0 = None
        if dt0 is None:
            term_ = self.term if term is None else term
            assert term_ is not None
            dt0, f0 = _launch.select_initial_step(self, term_, problem, method_order, stats, args)

        def lim(v):
            return None if v is None else torch.tensor(v, dtype=problem.time_dtype,
                                                       device=problem.device)

        return dt0, self._new_state(method_order, problem, lim(self.dt_min), lim(self.dt_max)), f0

    def adapt_step_size(self, t0, dt, y0, step_result: StepResult, state, stats):
        y1, err = step_result.y, step_result.error_estimate
        if err is None:
            # no error estimate: accept, keep dt (step_size_controllers.py:382-391)
            new = state
            if self._pid:
                new = state.update_error_ratios(y0.new_ones(dt.shape), state.prev_error_ratio)
            return torch.ones_like(dt, dtype=torch.bool), dt, new, None
        out = _launch.adapt_step_size(self, state, dt, y0, y1, err)
        accept, dt_next, r1, r2, status = out
        return accept, dt_next, state.update_error_ratios(r1, r2), status

    def merge_states(self, running, current, previous):
        if not self._pid:
            return current
        return current.update_error_ratios(
            torch.where(running, current.prev_error_ratio, previous.prev_error_ratio),
            torch.where(running, current.prev_prev_error_ratio, previous.prev_prev_error_ratio))


class IntegralController(_AdaptiveController):
    """dt *= clamp(safety * ratio^(-1/order)) (step_size_controllers.py:260-294)."""

    def _exponents(self, order):
        k_i = 1.0 / order
        return -k_i, 0.0, 0.0

    def __repr__(self):
        return (f"IntegralController(atol={float(self.atol)}, rtol={float(self.rtol)}, "
                f"dt_min={self.dt_min}, dt_max={self.dt_max})")


class PIDController(_AdaptiveController):
    """Soederlind's PID step-size filter; coefficients are divided by the method order
    and use the last two *accepted* error ratios (step_size_controllers.py:554-671)."""

    _pid = True

    def __init__(self, atol: float, rtol: float, pcoeff: float, icoeff: float, dcoeff: float,
                 **kwargs):
        super().__init__(atol, rtol, **kwargs)
        self.pcoeff, self.icoeff, self.dcoeff = pcoeff, icoeff, dcoeff

    def _exponents(self, order):
        k_i, k_p, k_d = self.icoeff / order, self.pcoeff / order, self.dcoeff / order
        return -(k_i + k_p + k_d), k_p + 2 * k_d, -k_d

    def __repr__(self):
        return (f"PIDController(atol={float(self.atol)}, rtol={float(self.rtol)}, "
                f"pcoeff={self.pcoeff}, icoeff={self.icoeff}, dcoeff={self.dcoeff}, "
                f"dt_min={self.dt_min}, dt_max={self.dt_max})")

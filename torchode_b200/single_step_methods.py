"""Single-step methods (API of torchode/single_step_methods/).

``Dopri5`` and ``Tsit5`` keep the reference's constructor and plug-in protocol
(``init`` / ``step`` / ``merge_states`` / ``build_interpolation`` /
``convergence_order``).  Inside ``AutoDiffAdjoint.solve`` the per-stage arithmetic runs
in the CUDA stage / finish kernels; the protocol methods here are thin wrappers over
the same kernels for stand-alone use (custom controllers, unit tests).
"""
from typing import Any, Dict, Generic, NamedTuple, Optional, Tuple, TypeVar

import torch
import torch.nn as nn

from . import _cabi, _launch
from .interpolation import (FourthOrderPolynomialInterpolation, LinearInterpolation, LocalInterpolation,
                            ThirdOrderPolynomialInterpolation)
from .problems import InitialValueProblem
from .tableaus import DOPRI5, EULER, HEUN, TSIT5, ButcherTableau
from .terms import ODETerm


class StepResult(NamedTuple):
    y: torch.Tensor
    error_estimate: Optional[torch.Tensor]


MethodState = TypeVar("MethodState")
InterpolationData = TypeVar("InterpolationData")


class SingleStepMethod(nn.Module, Generic[MethodState, InterpolationData]):
    """Plug-in protocol of a stepping method (single_step_methods/base.py:20-89).

    ``init(term, problem, f0, *, stats, args) -> state``;
    ``step(term, running, y0, t0, dt, state, *, stats, args) -> (StepResult, interp_data,
    state, status | None)`` advances every sample from ``t0`` to ``t0 + dt``;
    ``merge_states(accept, current, previous)``; ``build_interpolation(interp_data)``;
    ``convergence_order()``.
    """

    def init(self, term, problem, f0, *, stats, args):
        raise NotImplementedError()

    def step(self, term, running, y0, t0, dt, state, *, stats, args):
        raise NotImplementedError()

    def merge_states(self, accept, current, previous):
        raise NotImplementedError()

    def build_interpolation(self, data) -> LocalInterpolation:
        raise NotImplementedError()

    def convergence_order(self) -> int:
        raise NotImplementedError()


class ERKInterpolationData(NamedTuple):
    tableau: ButcherTableau
    t0: torch.Tensor
    dt: torch.Tensor
    y0: torch.Tensor
    y1: torch.Tensor
    k: torch.Tensor  # (stages, batch, features)


class ERKState(NamedTuple):
    tableau: ButcherTableau
    prev_vf1: Optional[torch.Tensor]


class ExplicitRungeKutta(SingleStepMethod[ERKState, ERKInterpolationData]):
    """Generic explicit Runge-Kutta step over a Butcher tableau (runge_kutta.py:175-282)."""

    # dense-output recipe understood by the kernels; None = not fusable
    INTERP_ID: Optional[int] = None

    def __init__(self, term: Optional[ODETerm], tableau: ButcherTableau):
        super().__init__()
        self.term = term
        self.tableau = tableau

    def _term(self, term):
        term_ = self.term if term is None else term
        assert term_ is not None, "no ODE term: pass one to the method or to solve()"
        return term_

    def fusable(self) -> bool:
        """Can the fused kernels run this method (FSAL + SSAL tableau, known interpolant)?"""
        tb = self.tableau
        return self.INTERP_ID is not None and tb.fsal and tb.ssal and tb.n_stages <= _cabi.MAX_STAGES

    def to_cabi(self) -> _cabi.Tableau:
        interp = _cabi.INTERP_DOPRI5 if self.INTERP_ID is None else self.INTERP_ID
        try:
            order = self.convergence_order()
        except NotImplementedError:  # a custom tableau stepped on its own (the order only matters to controllers)
            order = 0
        return self.tableau.to_cabi(interp, order)

    def init(self, term, problem: InitialValueProblem, f0, *, stats: Dict[str, Any], args: Any):
        prev = None
        if self.tableau.fsal:
            prev = f0 if f0 is not None else self._term(term).vf(problem.t_start, problem.y0, stats, args)
        tb = self.tableau.to(device=problem.device, time_dtype=problem.time_dtype,
                             data_dtype=problem.data_dtype)
        return ERKState(tb, prev)

    def merge_states(self, accept, current: ERKState, previous: ERKState):
        if current.prev_vf1 is None or previous.prev_vf1 is None:
            return current
        return ERKState(current.tableau,
                        torch.where(accept[:, None], current.prev_vf1, previous.prev_vf1))

    def step(self, term, running, y0, t0, dt, state: ERKState, *, stats: Dict[str, Any], args: Any):
        """One step for the whole batch: stage kernels around ``term.vf`` (runge_kutta.py:227-279)."""
        term_ = self._term(term)
        tb = self.tableau
        S = tb.n_stages
        reuse = tb.fsal and state.prev_vf1 is not None
        k0 = state.prev_vf1 if reuse else term_.vf(t0, y0, stats, args)
        cab = self.to_cabi()
        t_nodes = _launch.time_nodes(cab, t0, dt)
        ks = [k0.contiguous()]
        y_i = y0
        for i in range(1, S):
            y_i = _launch.erk_stage(cab, i, y0, dt, ks)
            ks.append(term_.vf(t_nodes[i], y_i, stats, args).contiguous())
        y1 = y_i if tb.ssal else _launch.erk_weighted_sum(cab, "b", dt, ks, base=y0)
        err = _launch.erk_weighted_sum(cab, "b_err", dt, ks)
        k = torch.stack(ks)
        new_state = ERKState(state.tableau, k[-1]) if tb.fsal else state
        return StepResult(y1, err), ERKInterpolationData(state.tableau, t0, dt, y0, y1, k), new_state, None


class _KernelQuartic(FourthOrderPolynomialInterpolation):
    """Quartic dense output whose coefficients are never materialised: ``evaluate`` runs
    ``tode_interp_eval`` straight from the step data (dopri5.py:54-60, tsit5.py:124-139)."""

    def __init__(self, interp_id: int, data: ERKInterpolationData):
        self.interp_id, self.data = interp_id, data
        self.t0, self.t1 = data.t0, data.t0 + data.dt
        self._cab = self._coefficients = None

    @property
    def cab(self):
        if self._cab is None:
            self._cab = self.data.tableau.to_cabi(self.interp_id, 5)
        return self._cab

    @property
    def coefficients(self):
        """The quartic in increasing powers of the unit coordinate, materialised on request (the reference's
        interpolation objects expose them; the solve loop never asks)."""
        if self._coefficients is None:
            d, w = self.data, self.data.tableau.b_other
            h = d.dt.to(dtype=d.y0.dtype)[:, None]
            if self.interp_id == _cabi.INTERP_DOPRI5:
                self._coefficients = FourthOrderPolynomialInterpolation.from_k(
                    d.t0, d.dt, d.y0, d.y1, d.k, w[0]).coefficients
            else:
                c2, c3, c4 = (h * torch.einsum("s, sbf -> bf", w[r], d.k) for r in range(3))
                self._coefficients = (d.y0, h * d.k[0], c2, c3, c4)
        return self._coefficients

    def evaluate(self, t, idx):
        d = self.data
        if not d.y0.is_cuda:  # stand-alone use on host tensors (plug-in authors, the reference's unit tests)
            return FourthOrderPolynomialInterpolation.evaluate(self, t, idx)
        return _launch.interp_eval(self.cab, d.t0, d.dt, d.y0, d.y1, d.k, t, idx)


class Dopri5(ExplicitRungeKutta):
    """Dormand-Prince 5(4) with the 4th-order dense output through the step midpoint."""

    TABLEAU = DOPRI5
    INTERP_ID = _cabi.INTERP_DOPRI5

    def __init__(self, term: Optional[ODETerm] = None):
        super().__init__(term, Dopri5.TABLEAU)

    def convergence_order(self):
        return 5

    def build_interpolation(self, data: ERKInterpolationData):
        return _KernelQuartic(_cabi.INTERP_DOPRI5, data)  # (no use of self: the reference's tests call it unbound)


class Tsit5(ExplicitRungeKutta):
    """Tsitouras 5(4) (Comput. Math. Appl. 62 (2011) 770-775) with its free interpolant."""

    TABLEAU = TSIT5
    INTERP_ID = _cabi.INTERP_TSIT5

    def __init__(self, term: Optional[ODETerm] = None):
        super().__init__(term, Tsit5.TABLEAU)

    def convergence_order(self):
        return 5

    def build_interpolation(self, data: ERKInterpolationData):
        return _KernelQuartic(_cabi.INTERP_TSIT5, data)


class Heun(ExplicitRungeKutta):
    """Heun's 2nd-order method with cubic Hermite dense output (single_step_methods/heun.py).
    Runs through the generic plug-in route: its stage combination, solution and error estimate
    are the same CUDA ops (``tode_erk_stage`` / ``tode_erk_weighted_sum``)."""

    TABLEAU = HEUN

    def __init__(self, term: Optional[ODETerm] = None):
        super().__init__(term, Heun.TABLEAU)

    def convergence_order(self):
        return 2

    def build_interpolation(self, data: ERKInterpolationData):
        return ThirdOrderPolynomialInterpolation.from_k(data.t0, data.dt, data.y0, data.y1, data.k)


class LinearInterpolationData(NamedTuple):
    t0: torch.Tensor
    dt: torch.Tensor
    y0: torch.Tensor
    y1: torch.Tensor


class Euler(SingleStepMethod[None, LinearInterpolationData]):
    """Forward Euler, no error estimate, linear dense output (single_step_methods/euler.py:22-84)."""

    def __init__(self, term: Optional[ODETerm]):
        super().__init__()
        self.term = term
        self._cab = EULER.to_cabi(_cabi.INTERP_DOPRI5, 1)

    def init(self, term, problem, f0, *, stats, args):
        return None

    def step(self, term, running, y0, t0, dt, state, *, stats, args):
        term_ = self.term if term is None else term
        assert term_ is not None
        k0 = term_.vf(t0, y0, stats, args)
        y1 = _launch.erk_stage(self._cab, 1, y0, dt, [k0])  # y0 + dt * k0 (euler.py:62)
        return StepResult(y1, None), LinearInterpolationData(t0, dt, y0, y1), state, None

    def merge_states(self, accept, current, previous):
        return None

    def convergence_order(self):
        return 1

    def build_interpolation(self, data: LinearInterpolationData):
        return LinearInterpolation(data.t0, data.dt, data.y0, data.y1)

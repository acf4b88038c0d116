"""scipy-style convenience entry point (API of torchode/interface.py:17-97)."""
from typing import Any, Callable, Dict, Optional, Tuple, Union

import torch

from .adjoints import AutoDiffAdjoint
from .problems import InitialValueProblem
from .single_step_methods import SingleStepMethod
from .solution import Solution
from .step_size_controllers import PIDController, StepSizeController
from .terms import ODETerm

METHODS: Dict[str, Callable[..., SingleStepMethod]] = {}


def register_method(name: str, constructor: Callable[..., SingleStepMethod]):
    METHODS[name] = constructor


def solve_ivp(f: Union[ODETerm, Callable], y0: torch.Tensor, t_eval: Optional[torch.Tensor], *,
              t_span: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
              method: Union[str, SingleStepMethod] = "tsit5", max_steps: Optional[int] = None,
              controller: Optional[StepSizeController] = None, dt0: Optional[torch.Tensor] = None,
              args: Any = None) -> Solution:
    """Solve ``y' = f(t, y)`` from ``y0``; defaults: Tsit5 + PID(atol=rtol=1e-7, 0.2/0.5/0.0)."""
    term = f if isinstance(f, ODETerm) else ODETerm(f, with_args=args is not None)
    if not isinstance(method, SingleStepMethod):
        method = METHODS[method](term=term)
    if controller is None:
        controller = PIDController(term=term, atol=1e-7, rtol=1e-7, pcoeff=0.2, icoeff=0.5, dcoeff=0.0)
    batch = y0.shape[0]
    if t_eval is not None and t_eval.ndim == 1:
        t_eval = t_eval.expand((batch, -1))
    t_start, t_end = t_span if t_span is not None else (t_eval[:, 0], t_eval[:, -1])
    if t_start.ndim == 0:
        t_start = t_start.expand(batch)
    if t_end.ndim == 0:
        t_end = t_end.expand(batch)
    problem = InitialValueProblem(y0, t_start, t_end, t_eval)
    return AutoDiffAdjoint(method, controller, max_steps=max_steps).solve(problem, term, dt0=dt0, args=args)

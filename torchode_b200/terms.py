"""The ODE term ``dy/dt = f(t, y)`` (mirrors torchode/terms.py:9-63)."""
from typing import Any, Callable, Dict

import torch
import torch.nn as nn

from .problems import InitialValueProblem


class ODETerm(nn.Module):
    def __init__(self, f: Callable, *, with_stats: bool = True, with_args: bool = False):
        """``f(t, y)`` (or ``f(t, y, args)`` if ``with_args``) returns dy/dt with y's shape/dtype.

        ``with_stats`` tracks ``stats["n_f_evals"]`` (a CPU int64 tensor that counts one
        evaluation for every sample per call, terms.py:42-58).
        """
        super().__init__()
        self.f = f
        self.with_stats = with_stats
        self.with_args = with_args

    def init(self, problem: InitialValueProblem, stats: Dict[str, Any]):
        if self.with_stats:
            stats["n_f_evals"] = torch.zeros(problem.batch_size, device="cpu", dtype=torch.long)

    def vf(self, t: torch.Tensor, y: torch.Tensor, stats: Dict[str, Any], args: Any) -> torch.Tensor:
        if self.with_stats:
            stats["n_f_evals"].add_(1)
        return self.f(t, y, args) if self.with_args else self.f(t, y)

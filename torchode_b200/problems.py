"""Initial value problem description (mirrors torchode/problems.py:6-90)."""
from typing import Optional

import torch


class InitialValueProblem:
    """``y0`` is ``(batch, features)``; times are per sample.

    Time and data may use different floating dtypes (all time arithmetic runs in
    the dtype of ``t_start``, all state arithmetic in the dtype of ``y0``).  Without
    ``t_eval`` the solution is reported at ``t_end`` only.
    """

    def __init__(
        self,
        y0: torch.Tensor,
        t_start: Optional[torch.Tensor] = None,
        t_end: Optional[torch.Tensor] = None,
        t_eval: Optional[torch.Tensor] = None,
    ):
        if t_start is None:
            assert t_eval is not None, "t_start or t_eval is required"
            t_start = t_eval[:, 0]
        if t_end is None:
            assert t_eval is not None, "t_end or t_eval is required"
            t_end = t_eval[:, -1]
        self.y0, self.t_start, self.t_end, self.t_eval = y0, t_start, t_end, t_eval
        self._time_direction = None

        assert y0.ndim == 2, "y0 must be (batch, features)"
        assert t_start.ndim == 1 and t_end.ndim == 1
        assert t_start.dtype == t_end.dtype
        assert y0.shape[0] == t_start.shape[0] == t_end.shape[0]
        assert y0.device == t_start.device == t_end.device
        if t_eval is not None:
            assert t_eval.ndim == 2
            assert t_eval.dtype == t_start.dtype
            assert t_eval.shape[0] == t_start.shape[0]
            assert t_eval.device == t_start.device

    @property
    def time_direction(self) -> torch.Tensor:
        """+1 forward in time, -1 backward (t_start == t_end counts as backward, problems.py:42).
        Computed on first use: the kernel routes derive the direction on the device themselves, and
        three tiny torch launches per problem are a third of a small solve."""
        if self._time_direction is None:
            self._time_direction = torch.where(self.t_end > self.t_start, 1, -1)
        return self._time_direction

    data_dtype = property(lambda self: self.y0.dtype)
    time_dtype = property(lambda self: self.t_start.dtype)
    device = property(lambda self: self.y0.device)
    batch_size = property(lambda self: self.y0.shape[0])
    n_features = property(lambda self: self.y0.shape[1])

    @property
    def n_evaluation_points(self):
        return 0 if self.t_eval is None else self.t_eval.shape[1]

    def __repr__(self):
        return (
            f"InitialValueProblem(y0={self.y0}, t_start={self.t_start}, "
            f"t_end={self.t_end}, t_eval={self.t_eval})"
        )

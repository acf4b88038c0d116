"""Built-in analytic vector fields.

Each field is an ordinary ``nn.Module`` computing ``f(t, y)`` with plain PyTorch ops, so
it works as the ``f`` of an ``ODETerm`` anywhere (including the reference torchode).
When such a field is the term of a solve on a B200, ``AutoDiffAdjoint.solve`` recognises
it and runs the fully fused whole-solve kernel (``tode_solve_fused``), whose in-register
evaluation performs exactly the arithmetic of ``forward`` below: one IEEE rounding per
PyTorch op, in the same order, no FMA contraction.
"""
import torch
import torch.nn as nn

from . import _cabi


class BuiltinField(nn.Module):
    """Marker base class: ``field_id`` / ``params`` select the fused kernel's field."""

    field_id: int = -1
    n_features = None  # required feature count, None = any (<= 4 on the fused path)

    def params(self):
        raise NotImplementedError


class LinearDecay(BuiltinField):
    """``y' = rate * y`` (README example: ``rate = -0.5``)."""

    field_id = _cabi.FIELD_LINEAR

    def __init__(self, rate: float):
        super().__init__()
        self.rate = float(rate)

    def params(self):
        return [self.rate]

    def forward(self, t, y):
        return self.rate * y


class VanDerPol(BuiltinField):
    """``x' = v``, ``v' = mu (1 - x^2) v - x``."""

    field_id = _cabi.FIELD_VAN_DER_POL
    n_features = 2

    def __init__(self, mu: float):
        super().__init__()
        self.mu = float(mu)

    def params(self):
        return [self.mu]

    def forward(self, t, y):
        x, v = y[:, 0], y[:, 1]
        dv = self.mu * (1 - x * x) * v - x
        return torch.stack((v, dv), dim=1)


class LotkaVolterra(BuiltinField):
    """``x' = alpha x - beta x z``, ``z' = delta x z - gamma z``."""

    field_id = _cabi.FIELD_LOTKA_VOLTERRA
    n_features = 2

    def __init__(self, alpha: float = 1.5, beta: float = 1.0, delta: float = 1.0, gamma: float = 3.0):
        super().__init__()
        self.alpha, self.beta, self.delta, self.gamma = map(float, (alpha, beta, delta, gamma))

    def params(self):
        return [self.alpha, self.beta, self.delta, self.gamma]

    def forward(self, t, y):
        x, z = y[:, 0], y[:, 1]
        xz = x * z
        dx = self.alpha * x - self.beta * xz
        dz = self.delta * xz - self.gamma * z
        return torch.stack((dx, dz), dim=1)

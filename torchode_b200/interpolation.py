"""Local (per-step) dense-output interpolants (API of torchode/interpolation.py).

On the built-in path nothing here runs: Dopri5 / Tsit5 dense output is evaluated inside
the finish kernel (or ``tode_interp_eval``) directly from the stage values.  The classes
below exist for plug-in authors whose own step methods hand explicit polynomial
coefficients to the generic loop; they only do a gather + Horner evaluation.
"""
from typing import Protocol, Sequence

import torch


class LocalInterpolation(Protocol):
    def evaluate(self, t: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        """Values at times ``t[n]`` of samples ``idx[n]`` -> ``(n, features)``."""
        raise NotImplementedError()


def _unit_coordinate(t, t0, t1, dtype):
    h = t1 - t0
    h = torch.where(h.abs() > 0.0, h, 1.0)  # finished samples step with dt == 0
    return ((t - t0) / h)[:, None].to(dtype=dtype)


def _horner(coefficients: Sequence[torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """``coefficients`` in increasing power order."""
    acc = coefficients[-1]
    for c in reversed(coefficients[:-1]):
        acc = torch.addcmul(c, acc, x)
    return acc


class _PolynomialInterpolation:
    DEGREE = 0

    def __init__(self, t0, t1, coefficients):
        assert len(coefficients) == self.DEGREE + 1
        self.t0, self.t1, self.coefficients = t0, t1, tuple(coefficients)

    def evaluate(self, t, idx):
        x = _unit_coordinate(t, self.t0[idx], self.t1[idx], self.coefficients[0].dtype)
        return _horner([c[idx] for c in self.coefficients], x)

    def __repr__(self):
        return f"{type(self).__name__}(t0={self.t0}, t1={self.t1}, coefficients={self.coefficients})"


class LinearInterpolation(_PolynomialInterpolation):
    DEGREE = 1

    def __init__(self, t0, dt, y0, y1):
        super().__init__(t0, t0 + dt, (y0, y1 - y0))


class ThirdOrderPolynomialInterpolation(_PolynomialInterpolation):
    """Cubic on [t0, t1]; ``coefficients[i]`` multiplies ``x**i`` on the unit interval."""

    DEGREE = 3

    @staticmethod
    def from_k(t0, dt, y0, y1, k):
        """Cubic Hermite through (y0, k[0]) and (y1, k[-1]) in monomial form."""
        h = dt.to(dtype=y0.dtype)[:, None]
        m0, m1 = h * k[0], h * k[-1]
        d = y0 - y1
        return ThirdOrderPolynomialInterpolation(
            t0, t0 + dt, (y0, m0, -3 * d - 2 * m0 - m1, 2 * d + m0 + m1))


class FourthOrderPolynomialInterpolation(_PolynomialInterpolation):
    """Quartic on [t0, t1]; ``coefficients[i]`` multiplies ``x**i`` on the unit interval."""

    DEGREE = 4

    @staticmethod
    def from_k(t0, dt, y0, y1, k, b_mid):
        """Quartic through y0, y1, the midpoint value ``y0 + dt * b_mid . k`` and the end slopes."""
        h = dt.to(dtype=y0.dtype)[:, None]
        m0, m1 = h * k[0], h * k[-1]
        ymid = y0 + h * torch.einsum("s, sbf -> bf", b_mid, k)
        a = 2 * (m1 - m0) - 8 * (y1 + y0) + 16 * ymid
        b = 5 * m0 - 3 * m1 + 18 * y0 + 14 * y1 - 32 * ymid
        c = m1 - 4 * m0 - 11 * y0 - 5 * y1 + 16 * ymid
        return FourthOrderPolynomialInterpolation(t0, t0 + dt, (y0, m0, c, b, a))

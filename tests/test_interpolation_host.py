"""Host-side interpolation utilities for plug-in authors (API of torchode/interpolation.py);
mirrors what the reference's tests/interpolation_test.py pins."""
import numpy as np
import torch

from torchode_b200.interpolation import (FourthOrderPolynomialInterpolation, LinearInterpolation,
                                         ThirdOrderPolynomialInterpolation)
from torchode_b200.tableaus import DOPRI5


def test_polynomials_evaluate_like_numpy():
    rng = np.random.default_rng(0)
    for cls, deg in ((ThirdOrderPolynomialInterpolation, 3), (FourthOrderPolynomialInterpolation, 4)):
        co = rng.normal(size=(deg + 1, 5, 2))
        t0, t1 = torch.tensor(rng.uniform(0, 1, 5)), torch.tensor(rng.uniform(2, 3, 5))
        interp = cls(t0, t1, tuple(torch.tensor(c) for c in co))
        t = t0 + (t1 - t0) * torch.tensor(rng.uniform(0, 1, 5))
        got = interp.evaluate(t, torch.arange(5))
        x = ((t - t0) / (t1 - t0)).numpy()
        want = np.stack([[np.polynomial.polynomial.polyval(x[b], co[:, b, f]) for f in range(2)] for b in range(5)])
        assert np.allclose(got.numpy(), want, rtol=1e-12)


def test_zero_length_step_gives_no_nan():
    z = torch.zeros(3)
    y = torch.ones(3, 2)
    for interp in (LinearInterpolation(z, z, y, y),
                   ThirdOrderPolynomialInterpolation(z, z, (y, y, y, y)),
                   FourthOrderPolynomialInterpolation(z, z, (y, y, y, y, y))):
        out = interp.evaluate(z, torch.arange(3))
        assert torch.isfinite(out).all() and torch.equal(out, y)


def test_from_k_recovers_polynomial_solutions():
    # y(t) = t^3 - 2t on [t0, t0 + dt]: the cubic Hermite interpolant is exact
    t0, dt = torch.tensor([0.3, 1.0], dtype=torch.float64), torch.tensor([0.7, 0.25], dtype=torch.float64)
    y = lambda t: (t**3 - 2 * t)[:, None]
    dy = lambda t: (3 * t**2 - 2)[:, None]
    k = torch.stack((dy(t0), dy(t0 + dt)))
    cubic = ThirdOrderPolynomialInterpolation.from_k(t0, dt, y(t0), y(t0 + dt), k)
    tq = t0 + 0.4 * dt
    assert torch.allclose(cubic.evaluate(tq, torch.arange(2)), y(tq), rtol=1e-12)
    # quartic through the Dopri5 midpoint weights: exact for a quartic right-hand side in t
    y4 = lambda t: (t**4 - t**2 + 1)[:, None]
    dy4 = lambda t: (4 * t**3 - 2 * t)[:, None]
    c = DOPRI5.c
    ks = torch.stack([dy4(t0 + ci * dt) for ci in c])
    y1 = y4(t0 + dt)
    quartic = FourthOrderPolynomialInterpolation.from_k(t0, dt, y4(t0), y1, ks, DOPRI5.b_other[0])
    assert torch.allclose(quartic.evaluate(tq, torch.arange(2)), y4(tq), rtol=1e-6)

"""Butcher tableaux of the embedded explicit Runge-Kutta pairs.

Restates ``ButcherTableau`` (torchode/single_step_methods/runge_kutta.py:31-158) and
the coefficient tables of Dopri5 (dopri5.py:11-44) and Tsit5 (tsit5.py:74-114).
Coefficients are held in float64 exactly like the reference and are rounded to the
data / time dtype only inside the kernels (runge_kutta.py:107-121).
"""
from fractions import Fraction as Fr
from typing import List, Optional, Sequence

import torch

from . import _cabi


class ButcherTableau:
    """c: nodes, a: Runge-Kutta matrix, b: solution weights, b_err: error weights,
    b_other: extra weight rows used by the dense output."""

    def __init__(self, c, a, b, b_err, b_other=None, fsal: Optional[bool] = None,
                 ssal: Optional[bool] = None):
        self.c, self.a, self.b, self.b_err, self.b_other = c, a, b, b_err, b_other
        self.fsal = self.is_fsal() if fsal is None else fsal
        self.ssal = self.is_ssal() if ssal is None else ssal

    @staticmethod
    def from_lists(*, c: Sequence[float], a: Sequence[Sequence[float]], b: Sequence[float],
                   b_err: Optional[Sequence[float]] = None,
                   b_low_order: Optional[Sequence[float]] = None,
                   b_other: Optional[Sequence[Sequence[float]]] = None) -> "ButcherTableau":
        assert b_err is not None or b_low_order is not None, (
            "either the error weights or the weights of the embedded lower-order method are needed"
        )
        n = len(c)
        assert len(b) == n and len(a) == n
        f64 = torch.float64
        a_sq = [list(row) + [0.0] * (n - len(row)) for row in a]
        b_t = torch.tensor(list(b), dtype=f64)
        if b_err is None:
            assert len(b_low_order) == n
            # float64 subtraction, as runge_kutta.py:84-88 does it
            b_err_t = b_t - torch.tensor(list(b_low_order), dtype=f64)
        else:
            b_err_t = torch.tensor(list(b_err), dtype=f64)
        other = None
        if b_other is not None:
            other = torch.tensor([list(r) for r in b_other], dtype=f64)
            assert other.ndim == 2 and other.shape[1] == n
        return ButcherTableau(torch.tensor(list(c), dtype=f64), torch.tensor(a_sq, dtype=f64),
                              b_t, b_err_t, other)

    def to(self, device, time_dtype, data_dtype) -> "ButcherTableau":
        other = None if self.b_other is None else self.b_other.to(device, data_dtype)
        return ButcherTableau(self.c.to(device, time_dtype), self.a.to(device, data_dtype),
                              self.b.to(device, data_dtype), self.b_err.to(device, data_dtype),
                              other, fsal=self.fsal, ssal=self.ssal)

    @property
    def n_stages(self) -> int:
        return self.c.shape[0]

    def _explicit(self) -> bool:
        return bool((torch.triu(self.a, diagonal=1) == 0).all())

    def is_fsal(self) -> bool:
        """First stage of the next step == last stage of this one (runge_kutta.py:127-142)."""
        return (self._explicit() and bool((self.b == self.a[-1]).all()) and float(self.c[0]) == 0.0
                and float(self.c[-1]) == 1.0 and float(self.a[0, 0]) == 0.0)

    def is_ssal(self) -> bool:
        """Solution == last stage input (runge_kutta.py:144-158)."""
        return (self._explicit() and bool((self.b == self.a[-1]).all())
                and float(self.c[-1]) == 1.0 and float(self.a[-1, -1]) == 0.0)

    def to_cabi(self, interp: int, order: int) -> _cabi.Tableau:
        """Pack into the C-ABI struct (float64; kernels round per dtype)."""
        n = self.n_stages
        if n > _cabi.MAX_STAGES:
            raise ValueError(f"at most {_cabi.MAX_STAGES} stages are supported, got {n}")
        t = _cabi.Tableau()
        t.n_stages, t.interp, t.order = n, interp, order
        c, a, b, be = (x.double().cpu().tolist() for x in (self.c, self.a, self.b, self.b_err))
        for i in range(n):
            t.c[i], t.b[i], t.b_err[i] = c[i], b[i], be[i]
            for j in range(n):
                t.a[i][j] = a[i][j]
        if self.b_other is not None:
            w = self.b_other.double().cpu().tolist()
            for r in range(min(3, len(w))):
                for s in range(n):
                    t.w[r][s] = w[r][s]
        return t


def _f(*fracs) -> List[float]:
    return [float(x) for x in fracs]


# Dormand-Prince 5(4), 7 stages, FSAL + SSAL.  b_other[0] = weights of y(t + dt/2).
DOPRI5 = ButcherTableau.from_lists(
    c=_f(0, Fr(1, 5), Fr(3, 10), Fr(4, 5), Fr(8, 9), 1, 1),
    a=[
        [],
        _f(Fr(1, 5)),
        _f(Fr(3, 40), Fr(9, 40)),
        _f(Fr(44, 45), Fr(-56, 15), Fr(32, 9)),
        _f(Fr(19372, 6561), Fr(-25360, 2187), Fr(64448, 6561), Fr(-212, 729)),
        _f(Fr(9017, 3168), Fr(-355, 33), Fr(46732, 5247), Fr(49, 176), Fr(-5103, 18656)),
        _f(Fr(35, 384), 0, Fr(500, 1113), Fr(125, 192), Fr(-2187, 6784), Fr(11, 84)),
    ],
    b=_f(Fr(35, 384), 0, Fr(500, 1113), Fr(125, 192), Fr(-2187, 6784), Fr(11, 84), 0),
    b_low_order=_f(Fr(1951, 21600), 0, Fr(22642, 50085), Fr(451, 720), Fr(-12231, 42400),
                   Fr(649, 6300), Fr(1, 60)),
    b_other=[
        _f(Fr(6025192743, 2 * 30085553152), 0, Fr(51252292925, 2 * 65400821598),
           Fr(-2691868925, 2 * 45128329728), Fr(187940372067, 2 * 1594534317056),
           Fr(-1776094331, 2 * 19743644256), Fr(11237099, 2 * 235043384)),
    ],
)

# Tsitouras 5(4), 7 stages, FSAL + SSAL.  b_err is given directly (with the corrected
# sign of the last entry, tsit5.py:101-112).  b_other are the coefficients of x^2, x^3, x^4
# of the dense output in monomial form; the reference derives them with sympy at import
# time (tsit5.py:11-50) -- these are the float64 values it produces (sympy 1.14).
TSIT5 = ButcherTableau.from_lists(
    c=[0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0],
    a=[
        [],
        [0.161],
        [-0.008480655492356989, 0.335480655492357],
        [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
        [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
        [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401,
         -0.02826905039406838],
        [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
         2.324710524099774],
    ],
    b=[0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081,
       2.324710524099774, 0.0],
    b_err=[0.00178001105222577714, 0.0008164344596567469, -0.007880878010261995,
           0.1447110071732629, -0.5823571654525552, 0.45808210592918697, -1 / 66],
    b_other=[
        [-2.76370619727482580, 0.13169999999999998, 3.93029623689475116, -12.41107716693367635,
         37.50931341651104134, -27.89652628919728627, 1.5],
        [2.91325546182191264, -0.22339999999999999, -5.94103387213150480, 30.33818863028231760,
         -88.17890489476640425, 65.09189467479367863, -4.0],
        [-1.05308849772902158, 0.10170000000000000, 2.49062728565125280, -16.54810288924490180,
         47.37952196281928252, -34.87065786149661051, 2.5],
    ],
)

# Heun's method (explicit trapezoidal rule) with the embedded Euler step as error estimate.
HEUN = ButcherTableau.from_lists(c=[0.0, 1.0], a=[[], [1.0]], b=[0.5, 0.5], b_low_order=[1.0, 0.0])

# Forward Euler written as a 2-node tableau: its single stage combination y0 + dt * k0 is the step.
EULER = ButcherTableau.from_lists(c=[0.0, 1.0], a=[[], [1.0]], b=[1.0, 0.0], b_err=[0.0, 0.0])

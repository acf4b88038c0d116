"""B200-native batch-parallel adaptive explicit Runge-Kutta solve loop behind torchode's API.

Drop-in for the ``to.ODETerm / Dopri5 / Tsit5 / IntegralController / PIDController /
AutoDiffAdjoint.solve(InitialValueProblem)`` path of martenlienen/torchode: same names,
arguments, Solution and status codes; the arithmetic runs in hand-written sm_100a CUDA
kernels behind the C-ABI of ``include/torchode_b200.h`` (no CPU fallback).
"""

__version__ = "0.1.0"

from . import fields
from .adjoints import AutoDiffAdjoint
from .backsolve import BacksolveAdjoint, JointBacksolveAdjoint
from .host_pipeline import solve_from_host
from .interface import register_method, solve_ivp
from .problems import InitialValueProblem
from .single_step_methods import Dopri5, Euler, Heun, Tsit5
from .solution import Solution
from .status_codes import Status
from .step_size_controllers import FixedStepController, IntegralController, PIDController
from .terms import ODETerm

register_method("heun", Heun)
register_method("dopri5", Dopri5)
register_method("tsit5", Tsit5)

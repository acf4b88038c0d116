"""ctypes mirror of ``include/torchode_b200.h`` and the loader of the CUDA library.

The product path has no CPU fallback: if ``libtorchode_b200.so`` (hand-written
sm_100a kernels behind the C-ABI) is missing, :func:`lib` raises.
"""
import ctypes as C
import os
from pathlib import Path

ABI_VERSION = 3
SUMMARY_WORDS = 8
MAX_STAGES = 7
MAX_FIELD_PARAMS = 8
MAX_PEERS = 8

# enum tode_dtype
F32, F64 = 0, 1
# enum tode_norm
NORM_RMS, NORM_MAX = 0, 1
# enum tode_interp
INTERP_DOPRI5, INTERP_TSIT5 = 0, 1
# enum tode_field
FIELD_LINEAR, FIELD_VAN_DER_POL, FIELD_LOTKA_VOLTERRA = 0, 1, 2
# enum tode_weights
W_B, W_BERR = 0, 1
# enum tode_ctl_word
CTL_STOP, CTL_ITERS, CTL_RUNNING, CTL_FAILED, CTL_TICKET, CTL_NONMONO = range(6)
CTL_WORDS = 8

_vp = C.c_void_p


class Tableau(C.Structure):
    _fields_ = [
        ("n_stages", C.c_int32),
        ("interp", C.c_int32),
        ("order", C.c_int32),
        ("reserved", C.c_int32),
        ("c", C.c_double * MAX_STAGES),
        ("a", (C.c_double * MAX_STAGES) * MAX_STAGES),
        ("b", C.c_double * MAX_STAGES),
        ("b_err", C.c_double * MAX_STAGES),
        ("w", (C.c_double * MAX_STAGES) * 3),
    ]


class Controller(C.Structure):
    _fields_ = [
        ("norm", C.c_int32),
        ("pid", C.c_int32),
        ("has_dt_min", C.c_int32),
        ("has_dt_max", C.c_int32),
        ("atol", C.c_double),
        ("rtol", C.c_double),
        ("safety", C.c_double),
        ("factor_min", C.c_double),
        ("factor_max", C.c_double),
        ("exp_ratio", C.c_double),
        ("exp_prev", C.c_double),
        ("exp_prev2", C.c_double),
        ("dt_min", C.c_double),
        ("dt_max", C.c_double),
        ("almost_zero", C.c_double),
        ("max_steps", C.c_int64),
        ("iter_cap", C.c_int64),
    ]


class State(C.Structure):
    _fields_ = [
        ("B", C.c_int64),
        ("F", C.c_int64),
        ("T", C.c_int64),
        ("data_dtype", C.c_int32),
        ("time_dtype", C.c_int32),
        ("t_start", _vp),
        ("t_end", _vp),
        ("t_eval", _vp),
        ("t_eval_stride_b", C.c_int64),
        ("t", _vp),
        ("dt", _vp),
        ("y", _vp),
        ("f0", _vp),
        ("r1", _vp),
        ("r2", _vp),
        ("running", _vp),
        ("n_steps", _vp),
        ("n_accepted", _vp),
        ("status", _vp),
        ("cursor", _vp),
        ("not_yet", _vp),
        ("y_eval", _vp),
        ("t_nodes", _vp),
        ("ctl", _vp),
        ("scratch", _vp),
        ("scratch_elems", C.c_int64),
    ]


class Problem(C.Structure):
    _fields_ = [
        ("B", C.c_int64),
        ("F", C.c_int64),
        ("T", C.c_int64),
        ("data_dtype", C.c_int32),
        ("time_dtype", C.c_int32),
        ("y0", _vp),
        ("t_start", _vp),
        ("t_end", _vp),
        ("t_eval", _vp),
        ("t_eval_stride_b", C.c_int64),
        ("dt0", _vp),
    ]


class SolutionOut(C.Structure):
    _fields_ = [
        ("ys", _vp),
        ("n_steps", _vp),
        ("n_accepted", _vp),
        ("n_initialized", _vp),
        ("status", _vp),
        ("t_final", _vp),
        ("dt_final", _vp),
        ("summary", _vp),
        ("n_peers", C.c_int32),
        ("reserved", C.c_int32),
        ("peer_row0", C.c_int64),
        ("peer_ys", _vp * MAX_PEERS),
        ("peer_n_steps", _vp * MAX_PEERS),
        ("peer_n_accepted", _vp * MAX_PEERS),
        ("peer_n_initialized", _vp * MAX_PEERS),
        ("peer_status", _vp * MAX_PEERS),
        ("peer_global", _vp * MAX_PEERS),
    ]


KPtrs = _vp * MAX_STAGES

# name -> (restype, argtypes); every symbol include/torchode_b200.h declares
_P = C.POINTER
PROTOTYPES = {
    "tode_abi_version": (C.c_int, []),
    "tode_error_string": (C.c_char_p, [C.c_int]),
    "tode_scratch_elems": (C.c_int64, [C.c_int64, C.c_int64]),
    "tode_erk_stage": (C.c_int, [_P(Tableau), C.c_int, _P(State), _P(_vp), _vp, _vp]),
    "tode_erk_finish": (C.c_int, [_P(Tableau), _P(Controller), _P(State), _P(_vp), _vp, _vp]),
    "tode_init_step_a": (C.c_int, [_P(Tableau), _P(Controller), _P(State), _vp, _vp, _vp]),
    "tode_init_step_b": (C.c_int, [_P(Tableau), _P(Controller), _P(State), _vp, _vp]),
    "tode_init_with_dt0": (C.c_int, [_P(Tableau), _P(Controller), _P(State), _vp, _vp]),
    "tode_solve_fused": (
        C.c_int,
        [C.c_int, _P(C.c_double), _P(Tableau), _P(Controller), _P(Problem), _P(SolutionOut),
         C.c_int64, _vp],
    ),
    "tode_erk_weighted_sum": (
        C.c_int,
        [_P(Tableau), C.c_int, C.c_int32, C.c_int32, C.c_int64, C.c_int64, _vp, _P(_vp), _vp, _vp,
         _vp],
    ),
    "tode_time_nodes": (C.c_int, [_P(Tableau), C.c_int32, C.c_int64, _vp, _vp, _vp, _vp]),
    "tode_mlp_tanh256_forward": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.c_int32, _vp]),
    "tode_mlp_tanh256_stage_forward": (C.c_int, [_P(Tableau), C.c_int, _P(State), _P(_vp), _vp, _vp, _vp, _vp,
                                                 C.c_int32, _vp]),
    "tode_mlp_tanh256_step_forward": (C.c_int, [_P(Tableau), _P(State), _P(_vp), _vp, _vp, _vp, C.c_int32, _vp]),
    "tode_peer_push": (C.c_int, [_vp, _P(_vp), C.c_int32, C.c_int64, _vp]),
    "tode_heat1d_forward": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, C.c_double, C.c_int32, _vp]),
    "tode_heat_step": (C.c_int, [_P(Tableau), _P(Controller), _P(State), C.c_double, _vp, _vp, _vp, _vp]),
    "tode_bench_fp64_fma_threads": (C.c_int64, []),
    "tode_bench_fp64_fma": (C.c_int, [C.c_int64, _vp, _P(C.c_int64), _vp]),
    "tode_selftest_fast_math": (C.c_int, [C.c_int64, C.c_uint64, _P(Controller), _vp, _vp]),
    "tode_selftest_fast_math_f32": (C.c_int, [C.c_int64, C.c_uint64, _vp, _vp]),
    "tode_adapt_step_size": (
        C.c_int,
        [_P(Controller), C.c_int32, C.c_int32, C.c_int64, C.c_int64] + [_vp] * 12 + [_vp],
    ),
    "tode_interp_eval": (
        C.c_int,
        [_P(Tableau), C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int64] + [_vp] * 4
        + [_P(_vp), _vp, _vp, _vp, _vp],
    ),
}

LIB_NAME = "libtorchode_b200.so"
_lib = None


def lib_path() -> Path:
    return Path(os.environ.get("TORCHODE_B200_LIB", Path(__file__).resolve().parent / LIB_NAME))


def lib():
    """Load (once) the CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not path.exists():
            raise RuntimeError(
                f"{path} not found: torchode_b200 has no CPU / PyTorch fallback. Build the "
                "sm_100a kernels first (python -c 'import __graft_entry__ as g; g.build()' "
                "or make -C torchode_b200/csrc)."
            )
        handle = C.CDLL(str(path))
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        got = handle.tode_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"{path}: ABI version {got}, expected {ABI_VERSION}")
        _lib = handle
    return _lib


def check(code: int, what: str):
    if code != 0:
        msg = lib().tode_error_string(code)
        raise RuntimeError(f"{what} failed with code {code}: {msg.decode() if msg else '?'}")

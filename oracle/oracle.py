"""ORACLE -- TEST INFRASTRUCTURE ONLY.

numpy front-end of ``liberk_oracle.so`` (the plain-C CPU restatement of the reference's
solve loop, see erk_oracle.c).  May only be imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / ``--impl reference`` legs -- never by the product package.

Struct arguments are the ctypes mirrors of include/torchode_b200.h (torchode_b200._cabi),
i.e. exactly what the product passes to the CUDA library, but all pointers are HOST
pointers here.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from torchode_b200 import _cabi

HERE = Path(__file__).resolve().parent
LIB = HERE / "liberk_oracle.so"
_lib = None

_NP_DT = {np.dtype(np.float32): _cabi.F32, np.dtype(np.float64): _cabi.F64}


def build(force: bool = False):
    if force or not LIB.exists():
        subprocess.run(["make", "-C", str(HERE)] + (["-B"] if force else []), check=True,
                       capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.orc_det_pow_f64.restype = C.c_double
        _lib.orc_det_pow_f64.argtypes = [C.c_double, C.c_double]
        _lib.orc_det_pow_f32.restype = C.c_float
        _lib.orc_det_pow_f32.argtypes = [C.c_float, C.c_double]
        _lib.orc_det_log2.restype = C.c_double
        _lib.orc_det_log2.argtypes = [C.c_double]
        _lib.orc_det_exp2.restype = C.c_double
        _lib.orc_det_exp2.argtypes = [C.c_double]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _kptrs(k_list):
    arr = _cabi.KPtrs()
    for i, k in enumerate(k_list):
        arr[i] = k.ctypes.data
    return arr


def _chk(code, what):
    if code != 0:
        raise RuntimeError(f"oracle {what} failed: {code}")


class HostState:
    """Owns the numpy arrays of one path-A solve and the tode_state pointing at them."""

    def __init__(self, y0, t_start, t_end, t_eval=None, *, n_stages=7, pid=True, general=False):
        B, F = y0.shape
        D, T = y0.dtype, t_start.dtype
        self.B, self.F = B, F
        self.T = 0 if t_eval is None else t_eval.shape[1]
        self.t_start = np.ascontiguousarray(t_start)
        self.t_end = np.ascontiguousarray(t_end)
        self.t_eval = None if t_eval is None else np.ascontiguousarray(t_eval)
        self.t = self.t_start.copy()
        self.dt = np.zeros(B, T)
        self.y = np.ascontiguousarray(y0).copy()
        self.f0 = np.zeros((B, F), D)
        self.r1 = np.ones(B, D)
        self.r2 = np.ones(B, D)
        self.running = np.ones(B, np.uint8)
        self.n_steps = np.zeros(B, np.int32)
        self.n_accepted = np.zeros(B, np.int32)
        self.status = np.zeros(B, np.int32)
        self.cursor = np.zeros(B, np.int32)
        self.not_yet = np.ones((B, max(self.T, 1)), np.uint8) if general else None
        self.y_eval = np.full((B, max(self.T, 1), F), np.nan, D)
        self.t_nodes = np.zeros((n_stages, B), T)
        self.ctl = np.zeros(_cabi.CTL_WORDS, np.int32)
        self.scratch = np.zeros(2 * B + 16, D)
        s = _cabi.State()
        s.B, s.F, s.T = B, F, self.T
        s.data_dtype, s.time_dtype = _NP_DT[np.dtype(D)], _NP_DT[np.dtype(T)]
        s.t_start, s.t_end, s.t_eval = _p(self.t_start), _p(self.t_end), _p(self.t_eval)
        s.t_eval_stride_b = 0 if self.t_eval is None else self.T
        s.t, s.dt, s.y, s.f0 = _p(self.t), _p(self.dt), _p(self.y), _p(self.f0)
        s.r1, s.r2 = (_p(self.r1), _p(self.r2)) if pid else (None, None)
        s.running, s.n_steps, s.n_accepted = _p(self.running), _p(self.n_steps), _p(self.n_accepted)
        s.status, s.cursor, s.not_yet = _p(self.status), _p(self.cursor), _p(self.not_yet)
        s.y_eval, s.t_nodes, s.ctl = _p(self.y_eval), _p(self.t_nodes), _p(self.ctl)
        s.scratch, s.scratch_elems = _p(self.scratch), self.scratch.size
        self.c = s


def erk_stage(tab, stage, st: HostState, k_list, y_out):
    _chk(lib().orc_erk_stage(C.byref(tab), C.c_int(stage), C.byref(st.c), _kptrs(k_list), _p(y_out)),
         "erk_stage")
    return y_out


def erk_finish(tab, ctrl, st: HostState, k_list, y1):
    _chk(lib().orc_erk_finish(C.byref(tab), C.byref(ctrl), C.byref(st.c), _kptrs(k_list), _p(y1)),
         "erk_finish")


def init_step_a(tab, ctrl, st: HostState, y1_out, t1_out):
    _chk(lib().orc_init_step_a(C.byref(tab), C.byref(ctrl), C.byref(st.c), _p(y1_out), _p(t1_out)),
         "init_step_a")


def init_step_b(tab, ctrl, st: HostState, f1):
    _chk(lib().orc_init_step_b(C.byref(tab), C.byref(ctrl), C.byref(st.c), _p(f1)), "init_step_b")


def init_with_dt0(tab, ctrl, st: HostState, dt0):
    _chk(lib().orc_init_with_dt0(C.byref(tab), C.byref(ctrl), C.byref(st.c), _p(dt0)), "init_with_dt0")


def erk_weighted_sum(tab, which, dt, k_list, base=None):
    B, F = k_list[0].shape
    out = np.empty((B, F), k_list[0].dtype)
    _chk(lib().orc_erk_weighted_sum(C.byref(tab), C.c_int(which), _NP_DT[out.dtype], _NP_DT[dt.dtype],
                                    C.c_int64(B), C.c_int64(F), _p(dt), _kptrs(k_list), _p(base),
                                    _p(out)), "weighted_sum")
    return out


def erk_error_estimate(tab, dt, k_list):
    return erk_weighted_sum(tab, _cabi.W_BERR, dt, k_list)


def adapt_step_size(ctrl, dt, y0, y1, err, r1=None, r2=None):
    B, F = y0.shape
    D, T = y0.dtype, dt.dtype
    accept = np.zeros(B, np.uint8)
    dt_next = np.zeros(B, T)
    ratio, r1o, r2o = np.zeros(B, D), np.zeros(B, D), np.zeros(B, D)
    status = np.zeros(B, np.int64)
    _chk(lib().orc_adapt_step_size(C.byref(ctrl), _NP_DT[np.dtype(D)], _NP_DT[np.dtype(T)], C.c_int64(B),
                                   C.c_int64(F), _p(dt), _p(y0), _p(y1), _p(err), _p(r1), _p(r2),
                                   _p(accept), _p(dt_next), _p(ratio), _p(r1o), _p(r2o), _p(status)),
         "adapt_step_size")
    return dict(accept=accept.astype(bool), dt_next=dt_next, ratio=ratio, r1=r1o, r2=r2o, status=status)


def interp_eval(tab, t0, dt, y0, y1, k_list, t, idx):
    B, F = y0.shape
    N = t.shape[0]
    out = np.empty((N, F), y0.dtype)
    idx = np.ascontiguousarray(idx, np.int64)
    _chk(lib().orc_interp_eval(C.byref(tab), _NP_DT[y0.dtype], _NP_DT[t0.dtype], C.c_int64(B), C.c_int64(F),
                               C.c_int64(N), _p(t0), _p(dt), _p(y0), _p(y1), _kptrs(k_list), _p(t),
                               _p(idx), _p(out)), "interp_eval")
    return out


def solve_builtin(field_id, params, tab, ctrl, y0, t_start, t_end, t_eval=None, dt0=None, iter_cap=0):
    """Whole solve of a built-in field; returns a dict shaped like the reference's Solution."""
    y0 = np.ascontiguousarray(y0)
    B, F = y0.shape
    D, T = y0.dtype, t_start.dtype
    Tn = 0 if t_eval is None else t_eval.shape[1]
    t_start = np.ascontiguousarray(t_start)
    t_end = np.ascontiguousarray(t_end)
    stride_b = 0
    if t_eval is not None:
        if t_eval.strides[0] == 0:
            te = np.ascontiguousarray(t_eval[0])
        else:
            te = np.ascontiguousarray(t_eval)
            stride_b = Tn
    prob = _cabi.Problem()
    prob.B, prob.F, prob.T = B, F, Tn
    prob.data_dtype, prob.time_dtype = _NP_DT[np.dtype(D)], _NP_DT[np.dtype(T)]
    prob.y0, prob.t_start, prob.t_end = _p(y0), _p(t_start), _p(t_end)
    prob.t_eval = _p(te) if t_eval is not None else None
    prob.t_eval_stride_b = stride_b
    dt0 = None if dt0 is None else np.ascontiguousarray(dt0)
    prob.dt0 = _p(dt0)
    ys = np.full((B, max(Tn, 1), F), np.nan, D)
    n_steps, n_acc = np.zeros(B, np.int64), np.zeros(B, np.int64)
    n_init, status = np.zeros(B, np.int64), np.zeros(B, np.int64)
    t_final, dt_final = np.zeros(B, T), np.zeros(B, T)
    summary = np.zeros(4, np.int32)
    sol = _cabi.SolutionOut()
    sol.ys, sol.n_steps, sol.n_accepted = _p(ys), _p(n_steps), _p(n_acc)
    sol.n_initialized, sol.status = _p(n_init), _p(status)
    sol.t_final, sol.dt_final, sol.summary = _p(t_final), _p(dt_final), _p(summary)
    fp = (C.c_double * _cabi.MAX_FIELD_PARAMS)(*list(params))
    _chk(lib().orc_solve_builtin(C.c_int(field_id), fp, C.byref(tab), C.byref(ctrl), C.byref(prob),
                                 C.byref(sol), C.c_int64(iter_cap)), "solve_builtin")
    iters = int(summary[0])
    n_stages = tab.n_stages
    n_f_evals = (2 if dt0 is None else 1) + (n_stages - 1) * iters
    return dict(ys=ys, n_steps=n_steps, n_accepted=n_acc, n_initialized=n_init, status=status,
                t_final=t_final, dt_final=dt_final, iters=iters, n_f_evals=n_f_evals,
                first_fail=int(summary[1]), nonmono=int(summary[2]))


def det_pow(x, e, dtype=np.float64):
    if np.dtype(dtype) == np.float32:
        return np.float32(lib().orc_det_pow_f32(C.c_float(float(x)), C.c_double(e)))
    return lib().orc_det_pow_f64(C.c_double(float(x)), C.c_double(e))

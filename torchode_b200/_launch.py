"""ctypes glue between torch CUDA tensors and the C-ABI of ``libtorchode_b200.so``.

Everything here is plumbing: pointer extraction, stream handles, output allocation.  No
arithmetic of the solve loop is done with PyTorch ops, and nothing here runs on CPU
tensors -- a non-CUDA tensor raises.
"""
import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _cabi

_DTYPE_ID = {torch.float32: _cabi.F32, torch.float64: _cabi.F64}


def dtype_id(dtype: torch.dtype) -> int:
    try:
        return _DTYPE_ID[dtype]
    except KeyError:
        raise TypeError(f"torchode_b200 kernels support float32 and float64, got {dtype}") from None


def require_cuda(*tensors: torch.Tensor):
    for t in tensors:
        if t is not None and t.device.type != "cuda":
            raise RuntimeError(
                "torchode_b200 runs its solver arithmetic in sm_100a CUDA kernels and has no CPU / "
                f"PyTorch fallback; got a tensor on '{t.device}'. Move the problem to a CUDA device."
            )


def stream_ptr(device: torch.device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dense16(t: torch.Tensor) -> torch.Tensor:
    """Contiguous and 16-byte aligned (the layout contract of every (B,F) operand)."""
    t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone(memory_format=torch.contiguous_format)
    return t


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def kptrs(ks: Sequence[torch.Tensor]) -> "C.Array":
    arr = _cabi.KPtrs()
    for i, k in enumerate(ks):
        arr[i] = k.data_ptr()
    return arr


def _minimal_state(y: torch.Tensor, dt: torch.Tensor) -> _cabi.State:
    st = _cabi.State()
    st.B, st.F = y.shape
    st.data_dtype, st.time_dtype = dtype_id(y.dtype), dtype_id(dt.dtype)
    st.y, st.dt = y.data_ptr(), dt.data_ptr()
    return st


# ---- stand-alone protocol ops -------------------------------------------------------------
def time_nodes(cab: _cabi.Tableau, t0: torch.Tensor, dt: torch.Tensor) -> torch.Tensor:
    require_cuda(t0, dt)
    t0, dt = t0.contiguous(), dt.contiguous()
    out = torch.empty((cab.n_stages, t0.shape[0]), dtype=t0.dtype, device=t0.device)
    with torch.cuda.device(t0.device):
        _cabi.check(_cabi.lib().tode_time_nodes(C.byref(cab), dtype_id(t0.dtype), t0.shape[0],
                                                t0.data_ptr(), dt.data_ptr(), out.data_ptr(),
                                                stream_ptr(t0.device)), "tode_time_nodes")
    return out


def erk_stage(cab: _cabi.Tableau, stage: int, y0: torch.Tensor, dt: torch.Tensor,
              ks: List[torch.Tensor]) -> torch.Tensor:
    require_cuda(y0, dt, *ks)
    y0, dt = dense16(y0), dt.contiguous()
    ks = [dense16(k) for k in ks]
    out = torch.empty_like(y0)
    st = _minimal_state(y0, dt)
    with torch.cuda.device(y0.device):
        _cabi.check(_cabi.lib().tode_erk_stage(C.byref(cab), stage, C.byref(st), kptrs(ks),
                                               out.data_ptr(), stream_ptr(y0.device)), "tode_erk_stage")
    return out


def erk_weighted_sum(cab: _cabi.Tableau, which: str, dt: torch.Tensor, ks: List[torch.Tensor],
                     base: Optional[torch.Tensor] = None) -> torch.Tensor:
    require_cuda(dt, *ks)
    ks = [k.contiguous() for k in ks]
    dt = dt.contiguous()
    base = None if base is None else base.contiguous()
    out = torch.empty_like(ks[0])
    B, F = out.shape
    w = _cabi.W_B if which == "b" else _cabi.W_BERR
    with torch.cuda.device(out.device):
        _cabi.check(_cabi.lib().tode_erk_weighted_sum(
            C.byref(cab), w, dtype_id(out.dtype), dtype_id(dt.dtype), B, F, dt.data_ptr(), kptrs(ks),
            ptr(base), out.data_ptr(), stream_ptr(out.device)), "tode_erk_weighted_sum")
    return out


def interp_eval(cab: _cabi.Tableau, t0, dt, y0, y1, k, t, idx) -> torch.Tensor:
    require_cuda(t0, dt, y0, y1, k, t, idx)
    t0, dt, y0, y1, t = (x.contiguous() for x in (t0, dt, y0, y1, t))
    ks = [k[s].contiguous() for s in range(k.shape[0])]
    idx = idx.to(torch.int64).contiguous()
    B, F = y0.shape
    N = t.shape[0]
    out = torch.empty((N, F), dtype=y0.dtype, device=y0.device)
    with torch.cuda.device(y0.device):
        _cabi.check(_cabi.lib().tode_interp_eval(
            C.byref(cab), dtype_id(y0.dtype), dtype_id(t0.dtype), B, F, N, t0.data_ptr(), dt.data_ptr(),
            y0.data_ptr(), y1.data_ptr(), kptrs(ks), t.data_ptr(), idx.data_ptr(), out.data_ptr(),
            stream_ptr(y0.device)), "tode_interp_eval")
    return out


def adapt_step_size(controller, state, dt, y0, y1, err):
    """(accept, dt_next, prev_ratio, prev_prev_ratio, status) via tode_adapt_step_size."""
    require_cuda(dt, y0, y1, err)
    if not controller.fusable():
        raise NotImplementedError("custom norm functions are not supported by the CUDA controller")
    dt = dt.contiguous()
    y0, y1, err = dense16(y0), dense16(y1), dense16(err)
    B, F = y0.shape
    cab = controller.to_cabi(state.method_order, y0.dtype)
    if state.dt_min is not None:
        cab.has_dt_min, cab.dt_min = 1, float(state.dt_min)
    if state.dt_max is not None:
        cab.has_dt_max, cab.dt_max = 1, float(state.dt_max)
    dev = y0.device
    accept = torch.empty(B, dtype=torch.bool, device=dev)
    dt_next = torch.empty_like(dt)
    status = torch.empty(B, dtype=torch.long, device=dev)
    r1 = r2 = r1o = r2o = None
    if cab.pid:
        r1, r2 = state.prev_error_ratio.contiguous(), state.prev_prev_error_ratio.contiguous()
        r1o, r2o = torch.empty_like(r1), torch.empty_like(r2)
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().tode_adapt_step_size(
            C.byref(cab), dtype_id(y0.dtype), dtype_id(dt.dtype), B, F, dt.data_ptr(), y0.data_ptr(),
            y1.data_ptr(), err.data_ptr(), ptr(r1), ptr(r2), accept.data_ptr(), dt_next.data_ptr(), None,
            ptr(r1o), ptr(r2o), status.data_ptr(), stream_ptr(dev)), "tode_adapt_step_size")
    return accept, dt_next, r1o, r2o, status


class StagedState:
    """Device buffers of one stage-wise ("path A") solve and the ``tode_state`` over them."""

    def __init__(self, problem, n_stages: int, pid: bool, general: bool = False, persistent: bool = False):
        """``persistent``: own copies of the problem tensors, so that the same buffers (and a CUDA
        graph recorded over them) can be reused for later problems of the same shape (`rebind`)."""
        y0 = dense16(problem.y0)
        dev, D, Tt = y0.device, y0.dtype, problem.time_dtype
        B, F = y0.shape
        Tn = problem.n_evaluation_points
        self.B, self.F, self.T, self.device = B, F, Tn, dev
        self.persistent = persistent
        self.t_start, self.t_end = problem.t_start.contiguous(), problem.t_end.contiguous()
        if persistent:
            self.t_start, self.t_end = self.t_start.clone(), self.t_end.clone()
        self.t_eval, stride_b = None, 0
        self.t_eval_broadcast = False
        if problem.t_eval is not None:
            te = problem.t_eval
            if te.stride(0) == 0 and (Tn <= 1 or te.stride(1) == 1):
                self.t_eval = te  # broadcast row (e.g. solve_ivp's expand): no 6.4 GB copy
                self.t_eval_broadcast = True
                if persistent:
                    self.t_eval = te[0].clone()
            else:
                self.t_eval, stride_b = te.contiguous(), Tn
                if persistent and self.t_eval.data_ptr() == te.data_ptr():
                    self.t_eval = self.t_eval.clone()
        self.t = torch.empty(B, dtype=Tt, device=dev)
        self.dt = torch.empty(B, dtype=Tt, device=dev)
        self.y = y0.clone()
        self.f0 = torch.empty_like(y0)
        self.r1 = torch.empty(B, dtype=D, device=dev) if pid else None
        self.r2 = torch.empty(B, dtype=D, device=dev) if pid else None
        self.running = torch.empty(B, dtype=torch.uint8, device=dev)
        self.n_steps = torch.empty(B, dtype=torch.int32, device=dev)
        self.n_accepted = torch.empty(B, dtype=torch.int32, device=dev)
        self.status = torch.empty(B, dtype=torch.int32, device=dev)
        self.cursor = torch.zeros(B, dtype=torch.int32, device=dev)
        self.not_yet = torch.ones((B, Tn), dtype=torch.uint8, device=dev) if general and Tn else None
        self.y_eval = torch.empty((B, max(Tn, 1), F), dtype=D, device=dev)
        self.t_nodes = torch.empty((n_stages, B), dtype=Tt, device=dev)
        self.ctl = torch.zeros(_cabi.CTL_WORDS, dtype=torch.int32, device=dev)
        n_scr = int(_cabi.lib().tode_scratch_elems(B, F))
        self.scratch = torch.empty(n_scr, dtype=D, device=dev)
        # stage outputs handed to f: one buffer per stage (f may return a view of its input)
        self.y_stage = [torch.empty_like(y0) for _ in range(n_stages - 1)]
        st = _cabi.State()
        st.B, st.F, st.T = B, F, Tn
        st.data_dtype, st.time_dtype = dtype_id(D), dtype_id(Tt)
        st.t_start, st.t_end = self.t_start.data_ptr(), self.t_end.data_ptr()
        st.t_eval, st.t_eval_stride_b = ptr(self.t_eval), stride_b
        st.t, st.dt, st.y, st.f0 = (x.data_ptr() for x in (self.t, self.dt, self.y, self.f0))
        st.r1, st.r2 = ptr(self.r1), ptr(self.r2)
        st.running, st.n_steps = self.running.data_ptr(), self.n_steps.data_ptr()
        st.n_accepted, st.status = self.n_accepted.data_ptr(), self.status.data_ptr()
        st.cursor, st.not_yet = self.cursor.data_ptr(), ptr(self.not_yet)
        st.y_eval, st.t_nodes, st.ctl = self.y_eval.data_ptr(), self.t_nodes.data_ptr(), self.ctl.data_ptr()
        st.scratch, st.scratch_elems = self.scratch.data_ptr(), n_scr
        self.c = st

    def rebind(self, problem):
        """Load another problem of the same shape / layout into the persistent buffers."""
        assert self.persistent
        self.y.copy_(problem.y0)
        self.t_start.copy_(problem.t_start)
        self.t_end.copy_(problem.t_end)
        if self.t_eval is not None:
            self.t_eval.copy_(problem.t_eval[0] if self.t_eval_broadcast else problem.t_eval)
        if self.not_yet is not None:
            self.not_yet.fill_(1)


def select_initial_step(controller, term, problem, method_order, stats, args):
    """Hairer initial-step heuristic through tode_init_step_a/_b (returns dt0, f0)."""
    require_cuda(problem.y0, problem.t_start, problem.t_end)
    if not controller.fusable():
        raise NotImplementedError("custom norm functions are not supported by the CUDA controller")
    from .tableaus import DOPRI5  # only the order matters for the heuristic

    cab_t = DOPRI5.to_cabi(_cabi.INTERP_DOPRI5, method_order)
    cab_c = controller.to_cabi(method_order, problem.data_dtype)
    st = StagedState(problem, cab_t.n_stages, bool(cab_c.pid))
    dev = st.device
    lib = _cabi.lib()
    f0 = term.vf(problem.t_start, problem.y0, stats, args)
    st.f0.copy_(f0)
    y1 = torch.empty_like(st.y)
    t1 = torch.empty_like(st.t)
    with torch.cuda.device(dev):
        _cabi.check(lib.tode_init_step_a(C.byref(cab_t), C.byref(cab_c), C.byref(st.c), y1.data_ptr(),
                                         t1.data_ptr(), stream_ptr(dev)), "tode_init_step_a")
        f1 = dense16(term.vf(t1, y1, stats, args))
        _cabi.check(lib.tode_init_step_b(C.byref(cab_t), C.byref(cab_c), C.byref(st.c), f1.data_ptr(),
                                         stream_ptr(dev)), "tode_init_step_b")
    # init_step_b clamps to the time domain (adjoints.py:109), which solve() repeats; the
    # clamp is idempotent, so handing out the clamped dt changes nothing downstream
    return st.dt, f0

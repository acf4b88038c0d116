"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Stage-wise solve loop around an OPAQUE vector field, driven from Python with the oracle's
C ops (orc_erk_stage / orc_erk_finish / orc_init_*): the CPU restatement of
adjoints.py:43-311 + runge_kutta.py:227-279 for a user-supplied ``f(t, y) -> dy`` working on
numpy arrays.  Used by the parity tests of the product's staged route (path A).
"""
import numpy as np

from torchode_b200 import _cabi

from . import oracle as orc


def solve_opaque(f, tab, ctrl, y0, t_start, t_end, t_eval=None, dt0=None, general=False,
                 max_iters=1_000_000):
    S = tab.n_stages
    st = orc.HostState(y0, t_start, t_end, t_eval, n_stages=S, pid=bool(ctrl.pid), general=general)
    D = st.y.dtype

    def vf(t, y):
        out = np.ascontiguousarray(f(t.copy(), y.copy()), dtype=D)
        assert out.shape == y.shape
        return out

    st.f0[:] = vf(st.t_start, st.y)
    n_init = 1
    if dt0 is None:
        y1, t1 = np.zeros_like(st.y), np.zeros_like(st.t)
        orc.init_step_a(tab, ctrl, st, y1, t1)
        f1 = vf(t1, y1)
        n_init = 2
        orc.init_step_b(tab, ctrl, st, f1)
    else:
        orc.init_with_dt0(tab, ctrl, st, np.ascontiguousarray(dt0, dtype=st.t.dtype))
    if st.ctl[_cabi.CTL_NONMONO] and not general and st.T:
        return solve_opaque(f, tab, ctrl, y0, t_start, t_end, t_eval, dt0, True, max_iters)
    if general and st.T:
        st.not_yet[:, 0] = (st.cursor == 0)
    ks = [st.f0] + [None] * (S - 1)
    y_stage = [np.zeros_like(st.y) for _ in range(S - 1)]
    while not st.ctl[_cabi.CTL_STOP] and st.ctl[_cabi.CTL_ITERS] < max_iters:
        for s in range(1, S):
            orc.erk_stage(tab, s, st, ks[:s], y_stage[s - 1])
            ks[s] = vf(st.t_nodes[s], y_stage[s - 1])
        orc.erk_finish(tab, ctrl, st, ks, y_stage[S - 2])
    iters = int(st.ctl[_cabi.CTL_ITERS])
    if st.T == 0:
        n_initialized = np.ones(st.B, np.int64)
    elif general:
        n_initialized = np.array([np.searchsorted(row.astype(np.int32), 1, side="left") for row in st.not_yet],
                                 dtype=np.int64)
    else:
        n_initialized = st.cursor.astype(np.int64)
    return dict(ys=st.y_eval, n_steps=st.n_steps.astype(np.int64), n_accepted=st.n_accepted.astype(np.int64),
                n_initialized=n_initialized, status=st.status.astype(np.int64), iters=iters,
                n_f_evals=n_init + (S - 1) * iters, t_final=st.t, dt_final=st.dt, state=st)

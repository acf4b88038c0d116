"""The batch-parallel adaptive solve loop (API of torchode/adjoints.py:26-319).

``AutoDiffAdjoint.solve`` keeps the reference's signature and Solution, and picks one of
three routes:

* **fused** (path B) -- built-in step method + built-in controller + built-in analytic
  field: the whole solve is ONE kernel launch (``tode_solve_fused``).
* **staged** (path A) -- built-in step method + controller around an opaque user ``f``:
  per iteration 6 stage kernels interleaved with the 6 calls of ``f`` and one finish
  kernel; the host never synchronises inside an iteration, it polls a device control
  block a few iterations late.
* **generic** -- any foreign ``SingleStepMethod`` / ``StepSizeController`` plug-in object
  (the reference's operator API, e.g. the stubs of its test-suite): the loop bookkeeping
  is orchestrated from Python exactly like adjoints.py:135-260 and calls the plug-ins'
  protocol methods.

There is no CPU route for the built-in components: CPU tensors raise.
"""
import ctypes as C
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from . import _cabi, _launch, status_codes
from .fields import BuiltinField, Heat1D, TanhMLP256
from .problems import InitialValueProblem
from .single_step_methods import Dopri5, SingleStepMethod, Tsit5
from .solution import Solution
from .step_size_controllers import FixedStepController, IntegralController, PIDController, StepSizeController
from .terms import ODETerm

_INT32_MAX = 2**31 - 1


def _uniform_stats(term_, problem, stats: Dict[str, Any], n_f_evals: int):
    """``term.init`` for a solve whose evaluation count is known in one piece (the CUDA routes
    count loop iterations on the device): the same (B,) CPU int64 ``n_f_evals`` as ``init`` followed
    by ``n`` calls of ``vf`` (terms.py:42-58), but as a stride-0 expansion of one element -- zeros +
    fill of the literal tensor cost milliseconds of host time at B = 2^20..2^24, in which the GPU
    idles.  A term subclass may track more than the plain ODETerm: it keeps its own protocol."""
    if type(term_) is ODETerm:
        if term_.with_stats:
            stats["n_f_evals"] = torch.full((1,), n_f_evals, dtype=torch.long).expand(problem.batch_size)
        return
    term_.init(problem, stats)
    if "n_f_evals" in stats:
        stats["n_f_evals"].fill_(n_f_evals)


def plain_term_of(term_) -> bool:
    """A plain ODETerm around f(t, y): nothing but the evaluation count depends on the calls of f (the
    step- and stage-fused routes evaluate the built-in field inside their kernels, not through term.vf)."""
    return type(term_) is ODETerm and not term_.with_args


class AutoDiffAdjoint(nn.Module):
    def __init__(self, step_method: SingleStepMethod, step_size_controller: StepSizeController, *,
                 max_steps: Optional[int] = None, backprop_through_step_size_control: bool = True):
        super().__init__()
        self.step_method = step_method
        self.step_size_controller = step_size_controller
        self.max_steps = max_steps
        self.backprop_through_step_size_control = backprop_through_step_size_control
        #: how many loop iterations the host may run ahead of the device (staged route)
        self.lookahead = 3
        #: staged route: capture one loop iteration (6 x (stage kernel, f) + finish kernel) into a
        #: CUDA graph after the first, eagerly launched iteration and replay it.  Removes the launch
        #: latency that dominates small problems.  ``None`` (default) = automatic: on when ``f`` is one of
        #: this package's kernel fields (``fields.TanhMLP256``: known to be capturable and free of host
        #: side effects), off for user code, because the Python body of ``f`` then runs only once
        #: more (at capture): an opaque ``f`` must be capturable (no host sync, no data-dependent Python
        #: control flow) and must not count its calls or log on the host -- set ``True`` to opt in.
        self.use_cuda_graph: Optional[bool] = None
        #: loop iterations recorded into one graph: a replay then costs one graph launch and one poll of the control
        #: block per ``graph_iterations`` iterations (iterations after the stop flag are no-ops on the device)
        self.graph_iterations = 4
        #: stage-wise route with the built-in ``fields.Heat1D`` as f and no ``t_eval``: run a whole loop
        #: iteration as ONE pass over y (stage values and the stencil's neighbours stay on chip,
        #: ``tode_heat_step``) instead of 6 x (stage kernel, f) + finish.  With ``fields.TanhMLP256`` as
        #: f: every stage combination is formed inside the MLP kernel's operand load
        #: (``tode_mlp_tanh256_stage_forward``) instead of a stage kernel of its own.  Same bits.
        self.use_step_fusion = True
        #: bookkeeping of the last solve: route taken and number of kernels launched through the C-ABI
        self.last_run = {}
        # staged route with use_cuda_graph: persistent buffers + recorded iteration graph per
        # problem signature, reused by later solves (capture + instantiation cost ~10-50 ms)
        self._plans = {}
        self._rings = {}
        # solve_from_host: one CUDA stream per chunk and device (host_pipeline.py)
        self._host_streams = {}
        # handle under which the torch.compile operator finds this solver (compile_ops.py)
        from .compile_ops import solver_handle

        self._compile_handle = solver_handle(self)

    def _poll_ring(self, dev, look):
        """Pinned host mirror of the control block + events, allocated once per solver and device
        (cudaHostAlloc costs far more than a loop iteration)."""
        key = (str(dev), look)
        ring = self._rings.get(key)
        if ring is None:
            ring = (torch.zeros((look + 1, _cabi.CTL_WORDS), dtype=torch.int32, device="cpu").pin_memory(),
                    [torch.cuda.Event() for _ in range(look + 1)])
            self._rings[key] = ring
        return ring

    # ------------------------------------------------------------------------------------
    def _kernel_route(self) -> bool:
        """Built-in components whose arithmetic the CUDA kernels implement."""
        return (type(self.step_method) in (Dopri5, Tsit5) and self.step_method.fusable()
                and type(self.step_size_controller) in (IntegralController, PIDController)
                and self.step_size_controller.fusable())

    def solve(self, problem: InitialValueProblem, term: Optional[ODETerm] = None,
              dt0: Optional[torch.Tensor] = None, args: Any = None) -> Solution:
        if term is None:
            term_ = self.step_method.term
            assert term_ is not None, "pass the ODE term to solve() or to the step method"
        else:
            term_ = term
        if not self._kernel_route():
            self._refuse_silent_gradients(problem, term_, args)
            return self._solve_generic(problem, term, term_, dt0, args)
        if torch.compiler.is_compiling() and term is None and args is None:
            # inside torch.compile: one opaque operator with a fake kernel (compile_ops.py)
            from .compile_ops import solve_compiled

            return solve_compiled(self, problem, dt0)

        _launch.require_cuda(problem.y0, problem.t_start, problem.t_end, problem.t_eval, dt0)
        if torch.is_grad_enabled() and problem.batch_size > 0:
            from .autodiff import grad_leaves, solve_with_grad

            # everything the solution depends on differentiably: y0, the term's parameters, and whatever else
            # one evaluation of f reaches -- a model captured by a closure, tensors inside ``args``
            leaves = grad_leaves(term_, problem, args)
            if problem.y0.requires_grad or leaves:
                # forward = the CUDA loop (recorded), backward = recompute-based (autodiff.py)
                with torch.cuda.device(problem.device):
                    return solve_with_grad(self, problem, term_, dt0, args, leaves)
        if problem.batch_size == 0:
            return self._empty_solution(problem, term_)
        with torch.no_grad(), torch.cuda.device(problem.device):
            f = self._fused_eligible(problem, term_)
            if f is not None:
                sol = self._solve_fused(problem, term_, f, dt0)
                if sol is not None:
                    return sol
            return self._solve_staged(problem, term_, dt0, args)

    def _refuse_silent_gradients(self, problem, term_, args):
        """The generic route drives kernel-backed protocol ops (``Dopri5.step``, ``Heun.step``,
        ``IntegralController.adapt_step_size``, the built-in interpolants ...) whose outputs carry no
        ``grad_fn``: back-propagating through such a solve would silently yield zero / wrong gradients
        where the reference differentiates its eager loop.  Refuse instead."""
        if not torch.is_grad_enabled() or not problem.y0.is_cuda:
            return
        from .autodiff import grad_leaves
        from .single_step_methods import Euler, ExplicitRungeKutta

        kernel_backed = (isinstance(self.step_method, (ExplicitRungeKutta, Euler))
                         or isinstance(self.step_size_controller, (IntegralController, PIDController,
                                                                   FixedStepController)))
        if kernel_backed and (problem.y0.requires_grad or grad_leaves(term_, problem, args)):
            raise NotImplementedError(
                "gradients through this solver configuration are not implemented: the solve would run on the "
                "generic route, whose built-in components are forward-only CUDA kernels (gradients are "
                "available for Dopri5 / Tsit5 with IntegralController / PIDController and the rms / max norm, "
                "and through BacksolveAdjoint / JointBacksolveAdjoint); wrap the solve in torch.no_grad() "
                "if no gradient is needed")

    @staticmethod
    def _empty_solution(problem, term_) -> Solution:
        """Empty batch: nothing to launch (empty tensors have no device pointer to hand over)."""
        dev, Tn = problem.device, problem.n_evaluation_points
        stats: Dict[str, Any] = {}
        term_.init(problem, stats)
        zeros = torch.zeros(0, dtype=torch.long, device=dev)
        stats["n_steps"], stats["n_accepted"], stats["n_initialized"] = zeros, zeros.clone(), zeros.clone()
        ys = problem.y0.new_empty((0, max(Tn, 1), problem.n_features))
        ts = problem.t_eval if problem.t_eval is not None else problem.t_end[:, None]
        return Solution(ts=ts, ys=ys, stats=stats, status=zeros.clone())

    # ------------------------------------------------------------------------------------
    # route 1: fused whole-solve kernel
    # ------------------------------------------------------------------------------------
    def _solve_fused(self, problem, term_, field: BuiltinField, dt0) -> Optional[Solution]:
        return self._fused_finish(self._fused_launch(problem, term_, field, dt0))

    def _fused_eligible(self, problem, term_) -> Optional[BuiltinField]:
        """The built-in analytic field of a problem the fused whole-solve kernel covers, else None."""
        f = term_.f
        if (self._kernel_route() and isinstance(f, BuiltinField) and not term_.with_args
                and problem.n_features <= 4 and (f.n_features is None or f.n_features == problem.n_features)):
            return f
        return None

    def _fused_launch(self, problem, term_, field: BuiltinField, dt0, peers=None, rows=None) -> Dict[str, Any]:
        """Allocate the outputs and enqueue the fused kernel on the current stream -- no host
        synchronisation.  ``_fused_finish`` reads the batch summary (the one sync of the solve).

        ``peers``: a ``distributed.SymmetricWorkspace`` -- the kernel then also stores every result
        into each rank's gathered buffers (peer memory over NVLink) and publishes the iteration
        count to every rank (``tode_solution.peer_*``); ``rows``: the row block ``[a, b)`` of this rank's
        shard the problem is (default: the whole shard)."""
        lib = _cabi.lib()
        method, ctrl = self.step_method, self.step_size_controller
        dev, D, Tt = problem.device, problem.data_dtype, problem.time_dtype
        B, F, Tn = problem.batch_size, problem.n_features, problem.n_evaluation_points
        cab_t = method.to_cabi()
        cab_c = ctrl.to_cabi(method.convergence_order(), D, self.max_steps)
        y0 = _launch.dense16(problem.y0)
        t_start, t_end = problem.t_start.contiguous(), problem.t_end.contiguous()
        prob = _cabi.Problem()
        prob.B, prob.F, prob.T = B, F, Tn
        prob.data_dtype, prob.time_dtype = _launch.dtype_id(D), _launch.dtype_id(Tt)
        prob.y0, prob.t_start, prob.t_end = y0.data_ptr(), t_start.data_ptr(), t_end.data_ptr()
        t_eval = None
        if problem.t_eval is not None:
            te = problem.t_eval
            if te.stride(0) == 0 and (Tn <= 1 or te.stride(1) == 1):
                t_eval, prob.t_eval_stride_b = te, 0
            else:
                t_eval, prob.t_eval_stride_b = te.contiguous(), Tn
            prob.t_eval = t_eval.data_ptr()
        dt0_c = None if dt0 is None else dt0.to(Tt).contiguous()
        prob.dt0 = _launch.ptr(dt0_c)

        if peers is None:
            ys = torch.empty((B, max(Tn, 1), F), dtype=D, device=dev)
            n_steps = torch.empty(B, dtype=torch.long, device=dev)
            n_accepted = torch.empty(B, dtype=torch.long, device=dev)
            n_init = torch.empty(B, dtype=torch.long, device=dev)
            status = torch.empty(B, dtype=torch.long, device=dev)
        else:  # this rank's rows of its own gathered buffers: the kernel writes them exactly once
            ys, n_steps, n_accepted, n_init, status = peers.own_rows(B, max(Tn, 1), F, D, rows)
        summary = torch.empty(_cabi.SUMMARY_WORDS, dtype=torch.int32, device=dev)
        sol = _cabi.SolutionOut()
        sol.ys, sol.n_steps, sol.n_accepted = ys.data_ptr(), n_steps.data_ptr(), n_accepted.data_ptr()
        sol.n_initialized, sol.status, sol.summary = n_init.data_ptr(), status.data_ptr(), summary.data_ptr()
        fp = (C.c_double * _cabi.MAX_FIELD_PARAMS)(*field.params())
        if peers is not None:
            peers.fill(sol, B, max(Tn, 1), F, D, rows)

        def run(cap: int):
            _cabi.check(lib.tode_solve_fused(field.field_id, fp, C.byref(cab_t), C.byref(cab_c),
                                             C.byref(prob), C.byref(sol), cap, _launch.stream_ptr(dev)),
                        "tode_solve_fused")

        run(0)
        return dict(run=run, summary=summary, ys=ys, n_steps=n_steps, n_accepted=n_accepted, n_init=n_init,
                    status=status, problem=problem, term=term_, n_stage_evals=cab_t.n_stages - 1,  # FSAL
                    n_init_evals=2 if dt0 is None else 1,
                    keep=(y0, t_start, t_end, t_eval, dt0_c))  # inputs stay alive until the kernel ran

    def _fused_finish(self, ctx: Dict[str, Any], summary_host=None) -> Optional[Solution]:
        """``summary_host``: the batch summary if the caller already copied it to the host."""
        iters, first_fail, nonmono, launches = (ctx["summary"].tolist() if summary_host is None else summary_host)[:4]
        if nonmono:
            return None  # t_eval rows not monotone in time: the staged route has the general mode
        self.last_run = {"route": "fused", "kernel_launches": launches, "iterations": iters}
        if first_fail != _INT32_MAX and first_fail < iters:
            # a failure stops the WHOLE batch at that iteration (adjoints.py:186-190): replay
            # with every sample limited to the iterations the reference would have executed
            ctx["run"](first_fail)
            iters = ctx["summary"].tolist()[0]
            self.last_run = {"route": "fused+replay", "kernel_launches": 2 * launches, "iterations": iters}
        problem = ctx["problem"]
        stats: Dict[str, Any] = {}
        _uniform_stats(ctx["term"], problem, stats, ctx["n_init_evals"] + ctx["n_stage_evals"] * iters)
        stats["n_steps"], stats["n_accepted"], stats["n_initialized"] = ctx["n_steps"], ctx["n_accepted"], ctx["n_init"]
        ts = problem.t_eval if problem.t_eval is not None else problem.t_end[:, None]
        return Solution(ts=ts, ys=ctx["ys"], stats=stats, status=ctx["status"])

    # ------------------------------------------------------------------------------------
    # route 2: stage-wise kernels around an opaque f
    # ------------------------------------------------------------------------------------
    def _step_fusable(self, problem, term_, args, record, general: bool = False) -> bool:
        """Problems whose loop iteration ``tode_heat_step`` covers: the built-in stencil field as a plain
        ``f(t, y)``, rows of whole 16-byte vectors, ``t_eval`` rows (if any) monotone in the direction of
        time (``general``: the scan-all mask mode of the stage-wise kernels)."""
        vec = 16 // problem.y0.element_size()
        return (self.use_step_fusion and record is None and not general and type(term_.f) is Heat1D
                and plain_term_of(term_) and args is None and problem.n_features % vec == 0
                and problem.n_features >= 2 * vec)

    def _solve_staged(self, problem, term_, dt0, args, general: bool = False, record=None,
                      step_fusion: bool = True, iter_cap: int = 0) -> Solution:
        lib = _cabi.lib()
        method, ctrl = self.step_method, self.step_size_controller
        dev, D, Tt = problem.device, problem.data_dtype, problem.time_dtype
        B, F, Tn = problem.batch_size, problem.n_features, problem.n_evaluation_points
        step_fusion = step_fusion and self._step_fusable(problem, term_, args, record, general)
        stage_fusion = (self.use_step_fusion and record is None and type(term_.f) is TanhMLP256 and plain_term_of(term_)
                        and args is None and D == torch.float32 and F == TanhMLP256.WIDTH
                        and term_.f.weights.device == dev)
        cab_t = method.to_cabi()
        cab_c = ctrl.to_cabi(method.convergence_order(), D, self.max_steps)
        cab_c.iter_cap = int(iter_cap)
        S = cab_t.n_stages
        plan = None
        use_graph = self.use_cuda_graph
        if use_graph is None:  # automatic: only for fields known to be pure and capturable
            use_graph = type(term_.f) is TanhMLP256 and plain_term_of(term_) and args is None
        if use_graph and record is None:
            te = problem.t_eval
            key = (str(dev), B, F, Tn, D, Tt, general, step_fusion, stage_fusion, id(term_.f), id(args), dt0 is None,
                   int(self.graph_iterations),
                   None if te is None else (te.stride(0) == 0), bytes(cab_t), bytes(cab_c))
            plan = self._plans.get(key)
            if plan is None:
                if len(self._plans) >= 4:
                    self._plans.clear()
                # the plan keeps f and args alive: their ids are part of the key, and an id may be reused
                # by another object once the original is collected
                plan = {"st": _launch.StagedState(problem, S, bool(cab_c.pid), general=general, persistent=True),
                        "graph": None, "kp": None, "ks": None, "f": term_.f, "args": args}
                self._plans[key] = plan
            else:
                assert plan["f"] is term_.f and plan["args"] is args
                plan["st"].rebind(problem)
            st = plan["st"]
        else:
            st = _launch.StagedState(problem, S, bool(cab_c.pid), general=general)
        stream = _launch.stream_ptr(dev)
        tab_p, ctrl_p, st_p = C.byref(cab_t), C.byref(cab_c), C.byref(st.c)
        stats: Dict[str, Any] = {}
        plain_term = type(term_) is ODETerm
        if plain_term:
            # the per-call ``n_f_evals += 1`` of terms.py:58 on a 1-element stand-in (an O(B) host
            # pass per f call otherwise); the (B,) tensor is produced once at the end
            call_stats = {"n_f_evals": torch.zeros(1, dtype=torch.long)} if term_.with_stats else {}
        else:
            term_.init(problem, stats)
            call_stats = stats

        def vf(t, y):
            out = term_.vf(t, y, call_stats, args)
            if out.dtype != D:
                raise TypeError(f"f returned {out.dtype}, expected the dtype of y0 ({D})")
            if not out.is_contiguous() or out.data_ptr() % 16:
                out = _launch.dense16(out)
            return out

        # ---- initial step size / state (adjoints.py:94-126) ---------------------------------
        st.f0.copy_(vf(st.t_start, st.y))
        n_init_evals = 1
        if dt0 is None:
            y1, t1 = st.y_stage[0], st.t_nodes[0]
            _cabi.check(lib.tode_init_step_a(tab_p, ctrl_p, st_p, y1.data_ptr(), t1.data_ptr(), stream),
                        "tode_init_step_a")
            f1 = vf(t1, y1)
            n_init_evals = 2
            _cabi.check(lib.tode_init_step_b(tab_p, ctrl_p, st_p, f1.data_ptr(), stream), "tode_init_step_b")
        else:
            dt0_c = dt0.to(Tt).contiguous()
            _cabi.check(lib.tode_init_with_dt0(tab_p, ctrl_p, st_p, dt0_c.data_ptr(), stream),
                        "tode_init_with_dt0")
        if general and Tn:
            st.not_yet[:, 0] = (st.cursor == 0).to(torch.uint8)

        # ---- the loop: never blocks on the iteration just launched ---------------------------
        look = max(1, int(self.lookahead))
        pinned, events = self._poll_ring(dev, look)
        if plan is not None and plan["kp"] is not None:
            kp, ks = plan["kp"], plan["ks"]
        else:
            kp = _cabi.KPtrs()
            kp[0] = st.f0.data_ptr()
            ks = [st.f0] + [None] * (S - 1)
        stage, finish = lib.tode_erk_stage, lib.tode_erk_finish
        y_stage, t_nodes = st.y_stage, st.t_nodes

        if step_fusion:
            # second (y, f0) buffer pair + per-sample selector: an accepted step flips the selector
            # instead of copying y <- y1, f0 <- k[S-1]
            y_alt, f_alt = y_stage[0], y_stage[1]
            if plan is not None and "sel" in plan:
                sel = plan["sel"].zero_()
            else:
                sel = torch.zeros(B, dtype=torch.uint8, device=dev)
                if plan is not None:
                    plan["sel"] = sel

        def launch_fused_iteration(stream):
            rc = lib.tode_heat_step(tab_p, ctrl_p, st_p, term_.f.kappa, y_alt.data_ptr(), f_alt.data_ptr(),
                                    sel.data_ptr(), stream)
            if rc:
                _cabi.check(rc, "tode_heat_step")

        def launch_mlp_iteration(stream):
            # fields.TanhMLP256: the stage combination is formed while the tcgen05 kernel loads its
            # activation tile (tode_mlp_tanh256_stage_forward) -- no stage kernel, y_i is only stored
            # for the last stage (the finish kernel's y1)
            mlp = term_.f
            for i in range(1, S):
                ks[i] = torch.empty_like(st.y)  # kept alive until the finish kernel has consumed it
                kp[i] = ks[i].data_ptr()
            if self.use_step_fusion == "stages":  # one launch per stage (round 1)
                for i in range(1, S):
                    y_out = y_stage[S - 2].data_ptr() if i == S - 1 else None
                    rc = lib.tode_mlp_tanh256_stage_forward(tab_p, i, st_p, kp, y_out, mlp.weights.data_ptr(),
                                                            mlp.biases.data_ptr(), kp[i], mlp.n_layers, stream)
                    if rc:
                        _cabi.check(rc, "tode_mlp_tanh256_stage_forward")
            else:
                # all six stage evaluations in ONE launch: rows are independent, so a CTA walks the stages of
                # its rows without any grid-wide synchronisation (tode_mlp_tanh256_step_forward)
                rc = lib.tode_mlp_tanh256_step_forward(tab_p, st_p, kp, y_stage[S - 2].data_ptr(),
                                                       mlp.weights.data_ptr(), mlp.biases.data_ptr(), mlp.n_layers,
                                                       stream)
                if rc:
                    _cabi.check(rc, "tode_mlp_tanh256_step_forward")
            rc = finish(tab_p, ctrl_p, st_p, kp, y_stage[S - 2].data_ptr(), stream)
            if rc:
                _cabi.check(rc, "tode_erk_finish")

        def launch_iteration(stream):
            for i in range(1, S):
                y_i = y_stage[i - 1]
                rc = stage(tab_p, i, st_p, kp, y_i.data_ptr(), stream)
                if rc:
                    _cabi.check(rc, "tode_erk_stage")
                k_i = vf(t_nodes[i], y_i)
                ks[i] = k_i  # keep alive until the finish kernel has consumed it
                kp[i] = k_i.data_ptr()
            rc = finish(tab_p, ctrl_p, st_p, kp, y_stage[S - 2].data_ptr(), stream)
            if rc:
                _cabi.check(rc, "tode_erk_finish")

        if step_fusion:
            launch_iteration = launch_fused_iteration
        elif stage_fusion:
            launch_iteration = launch_mlp_iteration
        launched = iters_launched = 0
        ctl_host = None
        graph = plan["graph"] if plan is not None else None
        per_replay = max(1, int(self.graph_iterations))
        while True:
            if record is not None:
                record.snapshot(st)
            if graph is not None:
                graph.replay()
                iters_launched += per_replay
            else:
                launch_iteration(stream)
                iters_launched += 1
                if plan is not None and launched == 0:
                    graph = self._capture_iteration(launch_iteration, dev, per_replay)
                    plan["graph"], plan["kp"], plan["ks"] = graph, kp, ks
            slot = launched % (look + 1)
            pinned[slot].copy_(st.ctl, non_blocking=True)
            events[slot].record()
            launched += 1
            if launched >= look:
                old = (launched - look) % (look + 1)
                events[old].synchronize()
                ctl_host = pinned[old].tolist()
                if ctl_host[_cabi.CTL_NONMONO] and not general and Tn:
                    # t_eval rows are not monotone in time: redo with the scan-all mask
                    if record is not None:
                        record.snaps.clear()
                    return self._solve_staged(problem, term_, dt0, args, general=True, record=record,
                                              iter_cap=iter_cap)
                if ctl_host[_cabi.CTL_STOP]:
                    break
        torch.cuda.current_stream(dev).synchronize()
        ctl_host = st.ctl.tolist()
        iters = ctl_host[_cabi.CTL_ITERS]
        if record is not None:
            record.snapshot(st)  # state after the last iteration
        if step_fusion and bool(st.status.any()):
            # the step-fused kernels compute what an all-successful solve needs (no end-point value
            # for a step that fails without reaching t_end): redo on the stage-wise kernels
            return self._solve_staged(problem, term_, dt0, args, general=general, record=record, step_fusion=False)
        if Tn == 0 and iter_cap == 0 and bool((st.status != 0).any()) and bool(st.running.any()):
            # a failure aborted the batch while other samples were still running: the reference returns, for
            # those, the interpolant of this last iteration evaluated at t_end (adjoints.py:298-301).  The
            # finish kernel writes an end value only for a sample that finishes or fails itself -- unless it
            # knows the batch's last iteration: replay with that iteration as the cap (rare path).
            if record is not None:
                record.snaps.clear()
            return self._solve_staged(problem, term_, dt0, args, general=general, record=record,
                                      step_fusion=False, iter_cap=iters)
        route = "step-fused" if step_fusion else ("stage-fused" if stage_fusion else "staged")
        self.last_run = {"route": route + "+graph" if graph is not None else route, "iterations": iters,
                         "general": general,
                         "iterations_launched": iters_launched,
                         # 6 stage kernels + finish (3 launches in split mode) per launched iteration
                         # (step-fused heat route and MLP field with all stages in one launch: 2), + init
                         "kernel_launches_min": iters_launched * (2 if (step_fusion or (
                             stage_fusion and self.use_step_fusion != "stages")) else S) + (2 if dt0 is None else 1)}
        # speculative iterations after the stop flag are no-ops on the device
        if plain_term:
            _uniform_stats(term_, problem, stats, n_init_evals + (S - 1) * iters)
        elif "n_f_evals" in stats:
            stats["n_f_evals"].fill_(n_init_evals + (S - 1) * iters)
        stats["n_steps"] = st.n_steps.to(torch.long)
        stats["n_accepted"] = st.n_accepted.to(torch.long)
        if Tn:
            if general:
                # adjoints.py:289-292
                stats["n_initialized"] = torch.searchsorted(
                    st.not_yet.int(), torch.ones((B, 1), dtype=torch.int, device=dev)).squeeze(dim=1)
            else:
                stats["n_initialized"] = st.cursor.to(torch.long)
            ts = problem.t_eval
        else:
            stats["n_initialized"] = torch.ones(B, dtype=torch.long, device=dev)
            ts = problem.t_end[:, None]
        # persistent (graph-cached) buffers are overwritten by the next solve: hand out a copy
        ys = st.y_eval.clone() if st.persistent else st.y_eval
        return Solution(ts=ts, ys=ys, stats=stats, status=st.status.to(torch.long))

    @staticmethod
    def _capture_iteration(launch_iteration, dev, repeat=1):
        """Record ``repeat`` loop iterations into a CUDA graph: inside the capture torch's current stream
        is the capturing side stream, so the kernels launched through the C-ABI are handed that
        stream; f's outputs live in the graph's private memory pool (static addresses)."""
        graph = torch.cuda.CUDAGraph()
        torch.cuda.current_stream(dev).synchronize()
        with torch.cuda.graph(graph):
            for _ in range(repeat):
                launch_iteration(_launch.stream_ptr(dev))
        return graph

    # ------------------------------------------------------------------------------------
    # route 3: foreign plug-ins (the reference's operator API)
    # ------------------------------------------------------------------------------------
    def _solve_generic(self, problem, term, term_, dt0, args) -> Solution:
        method, controller = self.step_method, self.step_size_controller
        dev, B = problem.device, problem.batch_size
        t_end, t_eval = problem.t_end, problem.t_eval
        sign = problem.time_direction.to(dtype=problem.time_dtype)
        lo, hi = torch.minimum(problem.t_start, t_end), torch.maximum(problem.t_start, t_end)
        stats: Dict[str, Any] = {}
        term_.init(problem, stats)
        t, y = problem.t_start, problem.y0
        dt, ctrl_state, f0 = controller.init(term, problem, method.convergence_order(), dt0,
                                              stats=stats, args=args)
        meth_state = method.init(term, problem, f0, stats=stats, args=args)
        dt = torch.clamp(dt, lo - t, hi - t)
        if not self.backprop_through_step_size_control:
            dt = dt.detach()
        n_steps = torch.zeros(B, dtype=torch.long, device=dev)
        n_accepted = torch.zeros(B, dtype=torch.long, device=dev)
        running = torch.ones(B, dtype=torch.bool, device=dev)
        pending = y_eval = None
        if t_eval is not None:
            y_eval = y.new_empty((B, problem.n_evaluation_points, problem.n_features))
            pending = torch.ones_like(t_eval, dtype=torch.bool)
            at_start = t_eval[:, 0] == t
            y_eval[at_start, 0] = y[at_start]
            pending[at_start, 0] = False
        status = None
        while True:
            result, interp_data, meth_next, meth_status = method.step(
                term, running, y, t, dt, meth_state, stats=stats, args=args)
            accept, dt_next, ctrl_next, ctrl_status = controller.adapt_step_size(
                t, dt, y, result, ctrl_state, stats)
            if not self.backprop_through_step_size_control:
                dt_next = dt_next.detach()
            commit = accept & running
            t = torch.where(commit, t + dt, t)
            y = torch.where(commit[:, None], result.y, y)
            meth_state = method.merge_states(commit, meth_next, meth_state)
            n_steps += running
            n_accepted += commit
            running = torch.addcmul(-sign * t_end, sign, t) < 0.0

            status = meth_status
            if status is None:
                status = ctrl_status
            elif ctrl_status is not None:
                status = torch.maximum(status, ctrl_status)
            if self.max_steps is not None:
                base = status if status is not None else status_codes.SUCCESS
                status = torch.where(n_steps >= self.max_steps, status_codes.REACHED_MAX_STEPS, base)
            go_on = running.any()
            if status is not None:
                go_on = go_on & (status == status_codes.SUCCESS).all()

            if t_eval is not None:
                crossed = (torch.addcmul(-sign[:, None] * t_eval, sign[:, None], t[:, None]) >= 0.0) & pending
                if crossed.any():
                    interp = method.build_interpolation(interp_data)
                    rows, cols = crossed.nonzero(as_tuple=True)
                    y_eval[rows, cols] = interp.evaluate(t_eval[rows, cols], rows)
                    pending = pending & ~crossed

            dt = torch.clamp(torch.where(running, dt_next, dt), lo - t, hi - t)
            ctrl_state = controller.merge_states(running, ctrl_next, ctrl_state)
            if bool(go_on):
                continue
            break

        if status is None:
            status = torch.zeros((), dtype=torch.long, device=dev).expand(B)
        stats["n_steps"], stats["n_accepted"] = n_steps, n_accepted
        if t_eval is not None:
            stats["n_initialized"] = torch.searchsorted(
                pending.int(), torch.ones((B, 1), dtype=torch.int, device=dev)).squeeze(dim=1)
            return Solution(ts=t_eval, ys=y_eval, stats=stats, status=status)
        interp = method.build_interpolation(interp_data)
        y_end = interp.evaluate(t_end, torch.arange(B, device=dev))
        stats["n_initialized"] = torch.ones(B, dtype=torch.long, device=dev)
        return Solution(ts=t_end[:, None], ys=y_end[:, None], stats=stats, status=status)

    def __repr__(self):
        return (f"AutoDiffAdjoint(step_method={self.step_method}, "
                f"step_size_controller={self.step_size_controller}, max_steps={self.max_steps}, "
                f"backprop_through_step_size_control={self.backprop_through_step_size_control})")

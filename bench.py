#!/usr/bin/env python
"""Benchmark of the B200-native batch-parallel adaptive RK solve loop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1..c5] [--impl reference]

One "step" = one complete solve of one batch of synthetic initial value problems through the public API
(``AutoDiffAdjoint.solve``).  Metric (BASELINE.json): accepted RK steps/s in sample-steps (sum of
``stats["n_accepted"]`` / device time), whole job.

Main line = BASELINE.json ``configs[1]``: Van der Pol mu=10, Tsit5 + PIDController(1e-8, 1e-8, 0.2, 0.5, 0), batch
2^20 per GPU, dim 2, fp64, t in [0, 20], no t_eval (inputs from a CPU
``torch.Generator().manual_seed(1234 + rank)``).  For N > 1 (torchrun, one rank per GPU) every rank solves its own
2^20-sample slice (weak scaling, no collective in the step loop) and the step ends with the full-batch Solution on
every rank (peer stores of the fused kernel into symmetric memory).

The printed JSON line also carries
  per_config        the other four BASELINE configs in the same run (3 timed steps each): value, ms_per_step,
                    roofline, route; N > 1: strong scaling of the config's batch, the step ending with the full
                    Solution on every rank, the shard-only time and the exchange next to it
  roofline          dominant kernel of the workload (algorithmic bytes / event time / measured peak)
  roofline_kernels  the HBM-bound stage / finish kernels of the stage-wise path, timed live on 2^24 x 2 fp32
                    operands (each operand 128 MiB > L2)
  fp64_issue        achieved vs measured double-precision FMA rate (the fused C2 kernel is fp64-issue-bound)
  parity_checked    number of samples of the TIMED batch whose counts and ys bits equal the oracle's (rows 0..n-1)
  cpu_baseline      torchode itself (baseline/_ref, unmodified) eager on the host cores, bounded sample;
  cpu_baseline_port the oracle port (oracle/, C + OpenMP) on a bounded sample
  reference_cuda    torchode eager on the SAME GPU on the first 65 536 samples, and the parity of this repo's
                    result for those samples with it
  e2e               same metric with HOST buffers: H2D of the inputs and D2H of ys + stats inside the timed region
``--impl reference`` times the reference itself on the host cores: torch.compile(solver.solve) (compilation
excluded) if it compiles within its budget, else eager; without baseline/_ref the oracle port (kind "port").
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torchode_b200 as to  # noqa: E402
from torchode_b200 import _cabi, _launch  # noqa: E402
from torchode_b200.fields import LinearDecay, LotkaVolterra, VanDerPol  # noqa: E402

METRIC = "accepted_rk_steps_per_sec"
UNIT = "sample-steps/s"


# --------------------------------------------------------------------------------------------
# workloads
# --------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, name, batch):
        self.name, self.batch = name, batch

    def describe(self):
        raise NotImplementedError

    def describe_reference(self):
        """The workload without this repo's route: what the reference arm runs."""
        return self.describe().split(", ROUTE: ")[0]


class C2(Workload):
    """Van der Pol mu=10, Tsit5 + PID(1e-8,1e-8,.2,.5,0), fp64, t in [0,20], no t_eval."""

    data_dtype, dtype_name = torch.float64, "f64"

    def host_inputs(self, rank, batch):
        g = torch.Generator().manual_seed(1234 + rank)
        y0 = torch.rand(batch, 2, generator=g, dtype=torch.float64) * 4 - 2
        return dict(y0=y0, t_start=torch.zeros(batch, dtype=torch.float64),
                    t_end=torch.full((batch,), 20.0, dtype=torch.float64), t_eval=None)

    def components(self):
        field = VanDerPol(10.0)
        term = to.ODETerm(field)
        return field, to.Tsit5(term), to.PIDController(1e-8, 1e-8, 0.2, 0.5, 0.0, term=term)

    def describe(self):
        return ("configs[1]: Van der Pol mu=10, Tsit5+PID(1e-8,1e-8,0.2,0.5,0), fp64, t in [0,20], "
                "no t_eval, ROUTE: fused whole-solve kernel")

    def algorithmic_bytes(self, batch, T):
        return batch * (16 + 16 + 16 + 32)  # y0, t_start+t_end, ys, 3 int64 stats + status


class C3(Workload):
    """Lotka-Volterra, Dopri5 + I(1e-6,1e-3), fp32, 100 shared t_eval points in [0,10]."""

    data_dtype, dtype_name = torch.float32, "f32"

    def host_inputs(self, rank, batch):
        g = torch.Generator().manual_seed(1234 + rank)
        y0 = 1 + torch.rand(batch, 2, generator=g)
        t_row = torch.linspace(0, 10, 100)
        return dict(y0=y0, t_start=torch.zeros(batch), t_end=torch.full((batch,), 10.0), t_eval=t_row)

    def components(self):
        field = LotkaVolterra()
        term = to.ODETerm(field)
        return field, to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term)

    def describe(self):
        return ("configs[2]: Lotka-Volterra, Dopri5+I(1e-6,1e-3), fp32, 100 t_eval points in [0,10] "
                "(broadcast row), ROUTE: fused whole-solve kernel (packed fp32, persistent grid with lane refill)")

    def algorithmic_bytes(self, batch, T):
        return batch * (8 + 8 + T * 8 + 32)


class C1(Workload):
    """README example."""

    data_dtype, dtype_name = torch.float32, "f32"

    def host_inputs(self, rank, batch):
        return dict(y0=torch.tensor([[1.2], [5.0]]), t_start=torch.tensor([0.0, 3.0]),
                    t_end=torch.tensor([5.0, 4.0]),
                    t_eval=torch.stack((torch.linspace(0, 5, 10), torch.linspace(3, 4, 10))))

    def components(self):
        field = LinearDecay(-0.5)
        term = to.ODETerm(field)
        return field, to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term)

    def describe(self):
        return "configs[0]: README example, Dopri5+I(1e-6,1e-3), f=-0.5y, batch 2, 10 t_eval"

    def algorithmic_bytes(self, batch, T):
        return batch * (4 + 8 + T * 8 + 32)


def _mlp_field(device=None):
    """configs[3]: Sequential(Linear(256,256), Tanh, Linear, Tanh, Linear), default init under
    torch.manual_seed(1234), weights x3 -> the tcgen05 field."""
    from torchode_b200.fields import TanhMLP256

    torch.manual_seed(1234)
    seq = torch.nn.Sequential(torch.nn.Linear(256, 256), torch.nn.Tanh(), torch.nn.Linear(256, 256),
                              torch.nn.Tanh(), torch.nn.Linear(256, 256))
    with torch.no_grad():
        for p in seq.parameters():
            p.mul_(3.0)
    f = TanhMLP256.from_sequential(seq)
    return f if device is None else f.to(device)


class C4(Workload):
    """Neural ODE: 3x256 tanh MLP (tcgen05 bf16 field), Dopri5 + I(1e-6,1e-3), fp32 state, t in [0,10]."""

    data_dtype, dtype_name = torch.float32, "f32"
    staged, graph = True, True

    def host_inputs(self, rank, batch):
        g = torch.Generator().manual_seed(1234 + rank)
        return dict(y0=torch.randn(batch, 256, generator=g), t_start=torch.zeros(batch),
                    t_end=torch.full((batch,), 10.0), t_eval=None)

    def components(self, device=None):
        field = _mlp_field(device)
        term = to.ODETerm(field)
        return field, to.Dopri5(term), to.IntegralController(1e-6, 1e-3, term=term)

    def describe(self):
        return ("configs[3]: neural ODE, 3x256 tanh MLP field on tcgen05 (bf16 GEMM, fp32 state), Dopri5+"
                "I(1e-6,1e-3), t in [0,10], ROUTE: stage-wise route with CUDA-graph replay")

    def algorithmic_bytes(self, batch, T):
        return None

    def numpy_field(self):
        f = _mlp_field()
        W = f.weights.float().numpy()
        b = f.biases.numpy()

        def bf16(x):
            return torch.from_numpy(x).to(torch.bfloat16).float().numpy()

        def fn(t, y):
            h = bf16(y)
            for l in range(3):
                h = h @ W[l].T + b[l]
                if l < 2:
                    h = bf16(np.tanh(h))
            return h.astype(np.float32)
        return fn


class C5(Workload):
    """1-D heat equation, method of lines (opaque stencil f), Tsit5 + I(1e-6,1e-3), fp32, dim 2^20."""

    data_dtype, dtype_name = torch.float32, "f32"
    staged, graph = True, False
    N = 1 << 20
    KAPPA = 25.0

    def host_inputs(self, rank, batch, n=None):
        n = n or self.N
        g = torch.Generator().manual_seed(1234 + rank)
        x = torch.linspace(0, 1, n)
        amp = torch.rand(batch, 3, generator=g)
        y0 = sum(amp[:, k - 1:k] * torch.sin(k * torch.pi * x)[None] for k in (1, 2, 3))
        return dict(y0=y0, t_start=torch.zeros(batch), t_end=torch.ones(batch), t_eval=None)

    def components(self, device=None):
        from torchode_b200.fields import Heat1D

        field = Heat1D(self.KAPPA)  # one-pass stencil kernel (bit-identical to the PyTorch expression)
        term = to.ODETerm(field)
        return field, to.Tsit5(term), to.IntegralController(1e-6, 1e-3, term=term)

    def describe(self):
        return ("configs[4]: 1-D heat equation method of lines (fields.Heat1D stencil kernel as f), "
                "Tsit5+I(1e-6,1e-3), batch 64, dim 2^20, fp32, ROUTE: step-fused route (tode_heat_step: one pass per iteration)")

    def algorithmic_bytes(self, batch, T):
        return None

    def numpy_field(self):
        kappa = np.float32(self.KAPPA)

        def fn(t, y):
            out = np.zeros_like(y)
            out[:, 1:-1] = kappa * ((y[:, 2:] - np.float32(2) * y[:, 1:-1]) + y[:, :-2])
            return out
        return fn


WORKLOADS = {"c2": (C2, 1 << 20), "c3": (C3, 1 << 24), "c1": (C1, 2), "c4": (C4, 8192), "c5": (C5, 64)}


def make_problem(host, device):
    dev = {k: (None if v is None else v.to(device)) for k, v in host.items()}
    t_eval = dev["t_eval"]
    if t_eval is not None and t_eval.ndim == 1:
        t_eval = t_eval.expand(dev["y0"].shape[0], -1)  # stride-0 row: never materialised
    return to.InitialValueProblem(dev["y0"], dev["t_start"], dev["t_end"], t_eval)


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (the child starts with the
    warm-up; ``stop(since=...)`` keeps the samples taken after the timed region began)."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self, since=None):
        import datetime

        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            if since is not None:
                try:
                    when = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if when < since:
                        continue
                except ValueError:
                    pass
            parts = parts[1:]
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                  "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        busy = [s for s in sm if s >= 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def flush_l2(buf):
    buf.add_(1)  # 512 MiB read+write > 126 MB L2


def ev_pair():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


# --------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores, bounded sample
# --------------------------------------------------------------------------------------------
def cpu_run(workload, sample_batch, repeats=1, host=None, want_output=False):
    """The oracle port on ``sample_batch`` samples (``host``: explicit inputs, e.g. the first rows of the
    batch the GPU solved).  Returns (accepted steps / s, seconds, accepted[, oracle output])."""
    from oracle import oracle as orc
    from torchode_b200.single_step_methods import ExplicitRungeKutta  # noqa: F401

    if getattr(workload, "staged", False):
        return cpu_run_opaque(workload, sample_batch)
    field, method, ctrl = workload.components()
    if host is None:
        host = workload.host_inputs(0, sample_batch)
    tab = method.to_cabi()
    cc = ctrl.to_cabi(method.convergence_order(), host["y0"].dtype)
    t_eval = host["t_eval"]
    if t_eval is not None and t_eval.ndim == 1:
        t_eval = np.broadcast_to(t_eval.numpy(), (sample_batch, t_eval.shape[0]))
    elif t_eval is not None:
        t_eval = t_eval.numpy()
    args = (field.field_id, field.params(), tab, cc, host["y0"].numpy(), host["t_start"].numpy(),
            host["t_end"].numpy(), t_eval)
    orc.lib()
    best, out = None, None
    for _ in range(repeats):
        t0 = time.perf_counter()
        out = orc.solve_builtin(*args)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    acc = int(out["n_accepted"].sum())
    if want_output:
        return acc / best, best, acc, out
    return acc / best, best, acc


def cpu_run_opaque(workload, sample_batch):
    """Opaque-f workloads: the oracle's C ops driven around a numpy restatement of f."""
    from oracle import driver

    _, method, ctrl = workload.components()
    if workload.name == "c5":
        host = workload.host_inputs(0, sample_batch, n=1 << 16)  # bounded: 2^16 of the 2^20 grid points
    else:
        host = workload.host_inputs(0, sample_batch)
    tab = method.to_cabi()
    cc = ctrl.to_cabi(method.convergence_order(), host["y0"].dtype)
    t0 = time.perf_counter()
    out = driver.solve_opaque(workload.numpy_field(), tab, cc, host["y0"].numpy(), host["t_start"].numpy(),
                              host["t_end"].numpy())
    dt = time.perf_counter() - t0
    acc = int(out["n_accepted"].sum())
    return acc / dt, dt, acc


def cpu_sample_size(workload):
    return {"c2": 1 << 17, "c3": 1 << 20, "c1": 2, "c4": 1024, "c5": 8}[workload.name]


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_threads():
    """OpenMP threads the oracle port runs with: all host cores this process may use (torchrun
    exports OMP_NUM_THREADS=1, which would make the CPU arm look 10-20x slower than it is)."""
    from oracle import oracle as orc

    lib = orc.lib()
    lib.orc_set_threads.restype = C.c_int
    lib.orc_set_threads.argtypes = [C.c_int]
    return int(lib.orc_set_threads(cpu_cores()))


# --------------------------------------------------------------------------------------------
# the reference itself (torchode 1.0.1, staged verbatim under baseline/_ref): CPU eager, CPU
# torch.compile(solver.solve), and eager on the B200 -- bounded samples of the same seeded workload
# --------------------------------------------------------------------------------------------
REF_SAMPLE = {"c1": 2, "c2": 8192, "c3": 65536, "c4": 1024, "c5": 8}        # CPU (SURVEY.md 8(d) probes)
REF_SAMPLE_CUDA = {"c1": 2, "c2": 65536, "c3": 1 << 20, "c4": 8192, "c5": 8}  # torchode eager on the GPU
REF_C5_GRID = 1 << 16


def ref_host_inputs(workload, sb):
    if workload.name == "c5":
        return workload.host_inputs(0, sb, n=REF_C5_GRID)
    return workload.host_inputs(0, sb)


def ref_sample_text(workload, sb, what):
    extra = f" x {REF_C5_GRID} of the 2^20 grid points" if workload.name == "c5" else ""
    return f"{sb} of {workload.batch} samples{extra} of the same seeded workload per step, {what}"


def time_reference(name, host, device, mode, steps, warmup, budget_s=None):
    """Times ``steps`` solves of the staged reference (after ``warmup`` untimed ones; for
    mode="compiled" the first warm-up call contains the compilation and is reported separately).
    Returns dict(value, ms_per_step, accepted, n_steps_mean, steps, compile_s)."""
    from baseline import reference

    solver, problem = reference.build(name, host, device)
    solve = solver.solve
    compile_s = None
    is_cuda = torch.device(device).type == "cuda"
    sync = torch.cuda.synchronize if is_cuda else (lambda: None)
    with torch.no_grad():
        if mode == "compiled":
            # torch.compile(solver.solve) -- torch.compile(solver).solve(...) compiles nothing
            solve = torch.compile(solver.solve)
            t0 = time.perf_counter()
            solve(problem)
            compile_s = time.perf_counter() - t0
        times, sol = [], None
        t_begin = time.perf_counter()
        for i in range(warmup + steps):
            sync()
            t0 = time.perf_counter()
            sol = solve(problem)
            sync()
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if budget_s is not None and len(times) >= 1 and time.perf_counter() - t_begin > budget_s:
                break  # bounded leg (context numbers only; the reference arm proper runs all K steps)
    acc = int(sol.stats["n_accepted"].sum())
    sec = sum(times) / len(times)
    return {"value": acc / sec, "ms_per_step": 1e3 * sec, "accepted": acc, "steps": len(times),
            "n_steps_mean": float(sol.stats["n_steps"].float().mean()), "compile_s": compile_s,
            "status_nonzero": int((sol.status != 0).sum())}


def _reference_child(args, workload, out):
    """``--ref-child MODE``: one leg of the reference arm in its own process (a compile that fails or
    overruns its budget must not take the arm down); prints one JSON object."""
    mode = args.ref_child
    if mode == "compiled":
        if os.path.exists("/usr/bin/g++"):
            os.environ["CXX"] = "/usr/bin/g++"  # the image's default g++ wrapper (/opt/gcc/bin) lacks libgomp.spec
        os.environ.setdefault("TORCHINDUCTOR_CACHE_DIR", os.path.join(ROOT, "baseline", "_ref", "inductor_cache"))
    torch.set_num_threads(cpu_cores())
    device = "cuda" if mode == "cuda" else "cpu"
    sb = args.batch or (REF_SAMPLE_CUDA if mode == "cuda" else REF_SAMPLE)[workload.name]
    res = time_reference(workload.name, ref_host_inputs(workload, sb), device,
                         "compiled" if mode == "compiled" else "eager", args.steps, args.warmup,
                         budget_s=args.ref_budget)
    res["sample_batch"] = sb
    res["threads"] = torch.get_num_threads()
    out.emit(json.dumps(res))
    return 0


def reference_leg(workload, mode, steps, warmup, timeout_s, budget_s=None, batch=None):
    """Runs one leg (eager / compiled / cuda) of the reference in a child process; dict or {"error": ...}."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload.name,
           "--ref-child", mode, "--steps", str(steps), "--warmup", str(warmup)]
    if budget_s is not None:
        cmd += ["--ref-budget", str(budget_s)]
    if batch is not None:
        cmd += ["--batch", str(batch)]
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)  # torchrun exports OMP_NUM_THREADS=1
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout_s,
                              env=env)
    except subprocess.TimeoutExpired:
        return {"error": f"timed out after {timeout_s} s"}
    if proc.returncode != 0:
        return {"error": f"rc {proc.returncode}: {proc.stderr.strip().splitlines()[-1] if proc.stderr.strip() else ''}"}
    try:
        return json.loads(proc.stdout.strip().splitlines()[-1])
    except (ValueError, IndexError):
        return {"error": "no JSON from the child", "stdout": proc.stdout[-300:]}


def run_reference_arm(args, workload, out):
    """``--impl reference``: the reference's own implementation of the path on the host cores.
    With baseline/_ref staged: torchode itself -- K timed steps of torch.compile(solver.solve) (the
    north star's CPU path; compilation excluded) if the compile leg succeeds within its budget, else K
    timed eager steps; the other leg and torchode eager on the B200 ride along as context.  Without
    baseline/_ref: the oracle port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from baseline import reference

    threads = cpu_cores()
    line = {
        "impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": workload.dtype_name, "data": "synthetic", "gpu_launches": 0,
    }
    if reference.available():
        sb = REF_SAMPLE[workload.name]
        legs = {}
        legs["compiled"] = reference_leg(workload, "compiled", args.steps, args.warmup, timeout_s=args.ref_compile_timeout)
        primary = "compiled" if "error" not in legs["compiled"] else "eager"
        if primary == "eager":
            legs["eager"] = reference_leg(workload, "eager", args.steps, args.warmup, timeout_s=1200)
        else:
            legs["eager"] = reference_leg(workload, "eager", min(args.steps, 3), 1, timeout_s=600, budget_s=60)
        if "error" in legs[primary]:
            raise SystemExit(f"reference arm failed: {legs}")
        if torch.cuda.is_available():
            legs["cuda_eager"] = reference_leg(workload, "cuda", min(args.steps, 3), 1, timeout_s=600, budget_s=60)
        r = legs[primary]
        what = ("torchode 1.0.1 (baseline/_ref, unmodified), CPU, "
                + ("torch.compile(solver.solve), compilation excluded" if primary == "compiled" else "eager"))
        line.update({
            "value": r["value"], "ms_per_step": r["ms_per_step"],
            "config": {"workload": workload.describe_reference(), "batch_per_step": sb, "threads": r["threads"],
                       "reference_path": primary},
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                             "sample": ref_sample_text(workload, sb, what)},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reference_legs": legs,
        })
    else:
        sb = cpu_sample_size(workload)
        n_thr = cpu_threads()
        values, times = [], []
        for i in range(args.warmup + args.steps):
            v, t, _ = cpu_run(workload, sb)
            if i >= args.warmup:
                values.append(v)
                times.append(t)
        value = sum(values) / len(values)
        sample = (f"{sb} of {workload.batch} samples of the same seeded workload per step, oracle port "
                  f"(plain C + OpenMP, lock-step like the reference); baseline/_ref is not staged")
        line.update({
            "value": value, "ms_per_step": 1e3 * sum(times) / len(times),
            "config": {"workload": workload.describe(), "batch_per_step": sb},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_thr, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
    out.emit(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------
# live roofline of the stage-wise (path A) kernels
# --------------------------------------------------------------------------------------------
def measure_path_a_kernels(device, hbm_peak, reps=5, B=1 << 24, F=2, dtype=torch.float32):
    """Times tode_erk_stage (i = 1..6) and tode_erk_finish alone, CUDA events on the launch stream,
    on operands far larger than L2, every sample running.  Algorithmic bytes per launch:
    stage i: (i + 2) rows of F elements + dt + running flag per sample; finish: 9 rows read, the
    per-sample scalars read/written, 2 rows written for every accepted sample."""
    lib = _cabi.lib()
    g = torch.Generator().manual_seed(7)
    y0 = (1 + torch.rand(B, F, generator=g, dtype=dtype)).to(device)
    zeros = torch.zeros(B, device=device, dtype=dtype)
    problem = to.InitialValueProblem(y0, zeros, torch.full((B,), 1e6, device=device, dtype=dtype))
    method = to.Dopri5()
    ctrl = to.IntegralController(1e-6, 1e-3)
    cab_t, cab_c = method.to_cabi(), ctrl.to_cabi(5, dtype)
    st = _launch.StagedState(problem, 7, False)
    ks = [st.f0] + [torch.empty_like(y0) for _ in range(6)]
    base = torch.randn(B, F, generator=g, dtype=dtype).to(device) * 0.5
    for j, k in enumerate(ks):
        k.copy_(base + 1e-3 * j)  # smooth "derivatives": small error estimate, steps get accepted
    dt0 = torch.full((B,), 1e-3, device=device, dtype=dtype)
    stream = _launch.stream_ptr(device)
    _cabi.check(lib.tode_init_with_dt0(C.byref(cab_t), C.byref(cab_c), C.byref(st.c), dt0.data_ptr(), stream),
                "init")
    kp = _launch.kptrs(ks)
    e = et = y0.element_size()
    tag = f"{'f32' if e == 4 else 'f64'},B={B},F={F}"
    out = []
    for i in range(1, 7):
        times = []
        for _ in range(reps + 1):
            e0, e1 = ev_pair()
            e0.record()
            _cabi.check(lib.tode_erk_stage(C.byref(cab_t), i, C.byref(st.c), kp, st.y_stage[i - 1].data_ptr(),
                                           stream), "stage")
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = statistics.median(times[1:])
        byts = (i + 2) * B * F * e + B * (et + 1)
        out.append({"kernel": f"erk_stage_kernel<{tag},NK={i}>", "bytes": byts, "ms": ms,
                    "achieved": byts / ms / 1e6, "frac": byts / ms / 1e6 / hbm_peak})
    # finish: reset the mutable per-sample state before every launch
    times, n_acc = [], 0
    saved = dict(t=st.t.clone(), dt=st.dt.clone(), y=st.y.clone(), f0=st.f0.clone())
    for _ in range(reps + 1):
        st.t.copy_(saved["t"]); st.dt.copy_(saved["dt"]); st.y.copy_(saved["y"]); st.f0.copy_(saved["f0"])
        st.running.fill_(1); st.n_steps.zero_(); st.n_accepted.zero_(); st.ctl.zero_()
        e0, e1 = ev_pair()
        e0.record()
        _cabi.check(lib.tode_erk_finish(C.byref(cab_t), C.byref(cab_c), C.byref(st.c), kp,
                                        st.y_stage[5].data_ptr(), stream), "finish")
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        n_acc = int(st.n_accepted.sum())
    ms = statistics.median(times[1:])
    # reads: y, y1, k0..k6 (9 rows) + t, dt, t_start, t_end, n_steps, running; writes: t, dt, n_steps,
    # status, running, 6 t_nodes; accepted rows additionally write y and f0 (+ n_accepted r/w)
    byts = 9 * B * F * e + B * (4 * et + 4 + 1) + B * (2 * et + 4 + 4 + 1 + 6 * et) + n_acc * (2 * F * e + 8)
    out.append({"kernel": f"erk_finish_kernel<{tag}>", "bytes": byts, "ms": ms,
                "achieved": byts / ms / 1e6, "frac": byts / ms / 1e6 / hbm_peak,
                "accepted_fraction": n_acc / B})
    return out


def measure_fp64_peak(device):
    lib = _cabi.lib()
    n = int(lib.tode_bench_fp64_fma_threads())
    sink = torch.empty(n, dtype=torch.float64, device=device)
    n_fma = C.c_int64(0)
    best = None
    for _ in range(4):
        e0, e1 = ev_pair()
        e0.record()
        _cabi.check(lib.tode_bench_fp64_fma(4096, sink.data_ptr(), C.byref(n_fma), _launch.stream_ptr(device)),
                    "fp64 peak")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return n_fma.value / best / 1e9  # FMA per ms / 1e9 = tera-FMA per second


# --------------------------------------------------------------------------------------------
class _StdoutGuard:
    """Exactly ONE JSON line on stdout: everything else (NCCL banners, library chatter) that would
    be written to fd 1 during the run goes to stderr."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    with _StdoutGuard() as out:
        return _main(out)


class Env:
    """Process-wide context of one bench run."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device("cuda", self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=self.device)
            self.dist = dist
        self.hbm_peak, self.peak_src = peaks()
        self.l2buf = torch.zeros(128 << 20, dtype=torch.float32, device=self.device)  # 512 MiB

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def allreduce(self, values, op):
        t = torch.tensor(values, dtype=torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return t.tolist()


def redraw_failing_rows(workload, solver, host, device, rank):
    """configs[2] at its full batch: a handful of the 2^24 seeded samples (3 on rank 0) overflow fp32 in an
    over-long early step and end with INFINITE_NORM -- in the reference too -- which aborts the WHOLE batch
    at iteration 4 (adjoints.py:186-190).  To time a batch that completes, those rows (and only those) get
    fresh draws from a second seeded generator.  Untimed set-up; returns the number of rows replaced."""
    g = torch.Generator().manual_seed(4321 + rank)
    replaced = 0
    for _ in range(8):
        with torch.no_grad():
            sol = solver.solve(make_problem(host, device))
        bad = (sol.status != 0).nonzero().flatten().cpu()
        if bad.numel() == 0:
            break
        host["y0"][bad] = 1 + torch.rand(bad.numel(), 2, generator=g)
        replaced += int(bad.numel())
    return replaced


def run_workload(env, name, batch, steps, warmup, *, main, extras):
    """Times ``steps`` solves of one workload (after ``warmup`` untimed ones) on this rank's GPU -- for
    world > 1 on this rank's slice, the step ending with the full-batch Solution on every rank.
    ``main``: the workload of the JSON line proper (weak scaling: ``batch`` samples per GPU; e2e with host
    buffers is measured); else a ``per_config`` entry (strong scaling of the config's batch, 3 steps)."""
    from torchode_b200.distributed import SymmetricWorkspace, gather_solution, shard_bounds, solve_sharded_symmetric

    cls, _ = WORKLOADS[name]
    world, rank, device, dist = env.world, env.rank, env.device, env.dist
    if main or world == 1:
        B = batch
        workload = cls(name, B)
        host = workload.host_inputs(rank, B)
        scaling = "weak"
    else:  # strong scaling: this rank's slice of the config's batch (same seeded global inputs on every rank)
        workload = cls(name, batch)
        full = workload.host_inputs(0, batch)
        lo, hi = shard_bounds(batch, rank, world)
        assert (hi - lo) * world == batch, "per_config entries need batches divisible by the world size"
        host = {k: (v if (v is None or (k == "t_eval" and v.ndim == 1)) else v[lo:hi].clone()) for k, v in full.items()}
        B = hi - lo
        scaling = "strong"
    staged_wl = getattr(workload, "staged", False)
    field, method, ctrl = workload.components(device) if staged_wl else workload.components()
    solver = to.AutoDiffAdjoint(method, ctrl)
    # (C4's kernel field gets CUDA-graph replay automatically; C5's step-fused route launches 2 kernels per
    # 0.36 ms iteration and runs without)
    redrawn = 0
    if name == "c3":
        redrawn = redraw_failing_rows(workload, solver, host, device, rank)
    problem = make_problem(host, device)
    T = problem.n_evaluation_points
    F = int(problem.n_features)

    # N > 1, fused route: the gathered Solution is assembled in every rank's symmetric (peer-mapped)
    # buffers while the shards solve -- statistics by the kernel's own peer stores, dense-output blocks
    # by bulk pushes that overlap the next chunk's solve; stage-wise workloads gather with NCCL
    ws, ts_full = None, None
    if world > 1 and not staged_wl:
        try:
            ws = SymmetricWorkspace(B, T, F, problem.data_dtype, device)
        except Exception as exc:  # no symmetric memory on this box: every rank falls back to NCCL
            print(f"[rank {rank}] symmetric memory unavailable ({type(exc).__name__}: {exc}); NCCL gather", file=sys.stderr)
            ws = None
        if not int(env.allreduce([1 if ws is not None else 0], "MIN")[0]):
            ws = None
    if ws is not None:
        if T == 0:
            ts_full = problem.t_end.new_empty((B * world, 1))
            dist.all_gather_into_tensor(ts_full, problem.t_end[:, None].contiguous())
        else:
            ts_full = problem.t_eval[:1].expand(B * world, -1)  # the workloads share one t_eval row
    n_chunks_mg = 8 if (ws is not None and T > 0) else 1

    def step():
        if ws is not None:
            return solve_sharded_symmetric(solver, problem, ws, ts=ts_full, chunks=n_chunks_mg)
        sol = solver.solve(problem)
        if world > 1:
            sol = gather_solution(sol, B * world, ts=None if T == 0 else problem.t_eval)
        return sol

    with torch.no_grad():
        sampler = ClockSampler(torch.cuda.current_device()) if (rank == 0 and main) else None
        for _ in range(warmup):
            flush_l2(env.l2buf)
            step()
        env.barrier()
        times = []
        t_epoch, t_wall = time.time(), time.perf_counter()
        for _ in range(steps):
            flush_l2(env.l2buf)
            e0, e1 = ev_pair()
            e0.record()
            step()
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        env.barrier()
        t_wall = time.perf_counter() - t_wall
        clocks = sampler.stop(since=t_epoch - 0.05) if sampler is not None else None
        timed_route = dict(solver.last_run)

        # the shard alone (no gather), for the multi-GPU lines: what the exchange costs on top
        solve_only_ms = None
        if world > 1:
            ts_ = []
            for _ in range(3):
                flush_l2(env.l2buf)
                e0, e1 = ev_pair()
                e0.record()
                solver.solve(problem)
                e1.record()
                e1.synchronize()
                ts_.append(e0.elapsed_time(e1))
            solve_only_ms = env.allreduce([statistics.median(ts_)], "MAX")[0]

        local = solver.solve(problem)  # per-rank statistics of ONE step (every step solves the same inputs)
        acc_local = int(local.stats["n_accepted"].sum())
        attempted_local = int(local.stats["n_steps"].sum())
        iters = (int(local.stats["n_f_evals"][0]) - 2) // 6
        last_run = dict(solver.last_run) if world == 1 else timed_route
        n_status = int((local.status != 0).sum())
        mean_steps = float(local.stats["n_steps"].float().mean())

        total_ms = env.allreduce([sum(times)], "MAX")[0]
        acc = env.allreduce([acc_local], "SUM")[0]
        ms_per_step = total_ms / steps
        value = acc / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel -------------------------------------------------------
    kernel_ms = statistics.median(times) if world == 1 else (solve_only_ms or ms_per_step)
    alg_bytes = workload.algorithmic_bytes(B, T)
    route = last_run.get("route", "")
    step_fused = route.startswith("step-fused")
    e = 4 if workload.dtype_name == "f32" else 8
    if alg_bytes is None:
        # stage-wise workloads: solver-owned algorithmic traffic = 44 F e per attempted sample-step
        # (the user's f is not part of it); step-fused route: y and f0 read, y1 and
        # k6 written per attempted step (accepting a step flips a buffer selector: no commit copy)
        alg_bytes = (4 if step_fused else 44) * F * e * attempted_local
    fused = route.startswith("fused")
    roofline = {
        "kernel": ("solve_fused_f2_kernel" if (fused and name == "c3") else "solve_fused_kernel") if fused else
                  "heat_step_kernel + finish_split_control_kernel (whole step incl. f)" if step_fused else
                  "erk_stage_kernel x6 + erk_finish_kernel (whole staged step incl. the user's f)",
        "bound": "hbm", "achieved": alg_bytes / kernel_ms / 1e6, "peak": env.hbm_peak, "unit": "GB/s",
        "frac": alg_bytes / kernel_ms / 1e6 / env.hbm_peak, "traffic": None,
        "peak_source": env.peak_src, "algorithmic_bytes_per_launch": alg_bytes,
    }
    if fused:
        roofline["note"] = ("whole solve in registers: HBM is touched only for inputs / outputs; the kernel is bound by "
                            + ("fp64 issue (see fp64_issue)" if name == "c2" else
                               "instruction issue (about 15 thread instructions per byte of output against a machine "
                               "balance of 5.6: at most 0.37 of the copy bandwidth at 100 % issue; DESIGN.md section 4)")
                            + "; the HBM-bound kernels of the stage-wise path are in roofline_kernels")
    res = {
        "value": value, "ms_per_step": ms_per_step, "scaling": scaling, "dtype": workload.dtype_name,
        "config": {"workload": workload.describe(), "batch_per_gpu": B, "global_batch": B * world, "features": F,
                   "t_eval_points": T, "loop_iterations": iters, "mean_n_steps": mean_steps,
                   "samples_with_failure_status": n_status},
        "roofline": roofline, "route": last_run, "steps": steps, "warmup": warmup,
        "gpu_launches": int(last_run.get("kernel_launches", last_run.get("kernel_launches_min", 0))) * steps,
    }
    if name == "c3" and fused:
        # the dense-output kernel is bound by instruction issue, not HBM: its issue-slot
        # utilisation next to the HBM figure.  Warp instructions per sample are a property of kernel + workload
        # read off the committed ncu capture (661.4 M for 2^20 samples, profiles/r02_ncu_f2_final.txt), not measured live
        sm_clock = 1.965e9
        props = torch.cuda.get_device_properties(device)
        instr = 630.7 * B
        res["issue_slots"] = {
            "bound": "issue", "warp_instr_per_sample": 630.7,
            "warp_instr_source": "static: profiles/r02_ncu_f2_final.txt (smsp__inst_executed.sum / samples)",
            "achieved_ginstr_per_s": instr / (kernel_ms * 1e-3) / 1e9,
            "peak_ginstr_per_s": props.multi_processor_count * 4 * sm_clock / 1e9,
            "frac": instr / (kernel_ms * 1e-3) / (props.multi_processor_count * 4 * sm_clock),
            "active_lanes_per_instr": 20.8,
            "note": "peak = SMs x 4 schedulers x 1.965 GHz; 20.8 of 32 lanes active (t_eval loop: a warp takes as many "
                    "trips as its busiest lane has points)"}
    if name == "c3":
        res["config"]["rows_redrawn"] = redrawn
        res["config"]["rows_redrawn_why"] = ("samples whose solve ends in INFINITE_NORM (fp32 overflow in an over-long "
                                             "early step, identically in the reference) abort the whole batch at "
                                             "iteration 4; they are re-drawn (untimed set-up) so the batch completes")
    if world > 1:
        res["solve_only_ms"] = solve_only_ms
        res["value_results_left_sharded"] = acc / (solve_only_ms * 1e-3)  # no exchange of ys / statistics
        res["config"]["multi_gpu"] = (
            ("independent batch slices; statistics by the fused kernel's peer stores (symmetric memory over NVLink), "
             f"dense-output blocks pushed to every peer in {n_chunks_mg} chunks under the next chunk's solve, "
             "iteration / failure counts by system-scope atomics, two signal-pad barriers per step, no collective")
            if ws is not None else "independent batch slices, NCCL all-gather of ys / statistics after the solve")
        if T > 0:
            pushed = B * T * F * e * (world - 1)
            res["exchange"] = {
                "bytes_sent_per_rank": pushed, "bytes_received_per_rank": pushed,
                "ms_on_top_of_the_solve": ms_per_step - solve_only_ms,
                "nvlink_gbs_per_rank_over_the_step": pushed / ms_per_step / 1e6,
                "limiter": "every rank receives the other ranks' dense output (the north star's all-gather of solutions): "
                           "at 900 GB/s per direction that alone takes bytes_received_per_rank / 900e9 s, "
                           f"{pushed / 900e9 * 1e3:.2f} ms here against {solve_only_ms:.2f} ms of solve"}
    # (configs[4] carries an e2e figure too: 64 rows of 4 MB -- the case solve_from_host cuts by bytes)
    if not main and not (name == "c5" and env.world == 1):
        del ws
        return res, None

    # ---- end to end through the public API with HOST buffers (every rank, max over ranks) --------
    host_pinned = {k: (None if v is None else v.pin_memory()) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host_pinned.values() if v is not None)
    te_h = host_pinned["t_eval"]
    if te_h is not None and te_h.ndim == 1:
        te_h = te_h.expand(B, -1)
    host_problem = to.InitialValueProblem(host_pinned["y0"], host_pinned["t_start"], host_pinned["t_end"], te_h)
    n_chunks = 8  # upper bound: solve_from_host runs fewer when the chunks would be launch-bound
    e2e_times, e2e_plain, d2h, host_out, hsol = [], [], 0, None, None
    with torch.no_grad():
        for i in range(2 + max(3, steps // 2)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            hsol = to.solve_from_host(solver, host_problem, device, chunks=n_chunks, out=hsol)
            dt = time.perf_counter() - t0
            chunks_run = solver.last_run.get("chunks")
            if i >= 2:
                e2e_times.append(dt)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            s = solver.solve(make_problem(host_pinned, device))
            outs = [s.ys, s.stats["n_steps"], s.stats["n_accepted"], s.stats["n_initialized"], s.status]
            if host_out is None:  # pinned result buffers, allocated once (first, untimed iteration)
                host_out = [torch.empty(o.shape, dtype=o.dtype, pin_memory=True) for o in outs]
            for h, o in zip(host_out, outs):
                h.copy_(o, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            d2h = sum(o.numel() * o.element_size() for o in host_out)
            if i >= 2:
                e2e_plain.append(dt)
    if n_status == 0:  # with failing samples the chunks stop separately (documented per-chunk scope)
        assert int(hsol.stats["n_accepted"].sum()) == acc_local, "host-pipelined solve disagrees"
    e2e_s = env.allreduce([statistics.median(e2e_times), statistics.median(e2e_plain)], "MAX")
    res["e2e"] = {"value": acc / e2e_s[0], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "ms_per_step": e2e_s[0] * 1e3,
                  "api": f"to.solve_from_host(solver, host_problem, device, chunks={n_chunks}): H2D, solve and D2H of "
                         "the chunks overlap on their own streams",
                  "chunks_run": chunks_run,
                  "unpipelined": {"value": acc / e2e_s[1], "ms_per_step": e2e_s[1] * 1e3,
                                  "api": "problem.to(device); solver.solve; results.to(pinned host)"}}
    if not main:
        del ws
        return res, None
    res["wall_s_timed_region"] = t_wall
    res["step_ms_rank0"] = [round(t, 4) for t in times]
    if clocks is not None:
        res["clocks"] = clocks
    ctx = dict(workload=workload, host=host, local=local, solver=solver, problem=problem, kernel_ms=kernel_ms,
               attempted_local=attempted_local, B=B)
    return res, ctx


def main_extras(env, line, ctx):
    """Rank 0, N = 1: measured context for the default line (none of it inside the timed region)."""
    workload, host, local, B = ctx["workload"], ctx["host"], ctx["local"], ctx["B"]
    device = env.device
    try:
        peak_fma = measure_fp64_peak(device)  # tera-FMA/s
        # fp64-pipe instructions (DFMA+DMUL+DADD+DSETP) per attempted sample-step of the fused Tsit5+PID Van der
        # Pol kernel: a property of the compiled kernel, read off the ncu opcode mix of the committed capture
        # (profiles/r01_ncu_fused_c2_v6.txt), not measured live
        ops_per_step = 210
        achieved = ctx["attempted_local"] / (ctx["kernel_ms"] * 1e-3) * ops_per_step / 1e12
        line["fp64_issue"] = {
            "peak_tfma_per_s": peak_fma, "peak_source": "tode_bench_fp64_fma, measured live",
            "fp64_pipe_instr_per_attempted_step": ops_per_step,
            "fp64_pipe_instr_source": "static: opcode mix of the kernel (profiles/r01_ncu_fused_c2_v6.txt)",
            "achieved_tinstr_per_s": achieved, "frac": achieved / peak_fma,
            "note": "only meaningful for the fp64 workload (c2)"}
    except Exception as exc:  # measurement aid only
        line["fp64_issue"] = {"error": str(exc)}
    try:
        line["roofline_kernels"] = measure_path_a_kernels(device, env.hbm_peak)
    except Exception as exc:
        line["roofline_kernels"] = {"error": str(exc)}
    # ---- CPU baselines on a bounded sample of the same seeded inputs + bit-for-bit parity of the GPU result ----
    try:
        sb = min(cpu_sample_size(workload), B)
        threads = cpu_threads()
        if getattr(workload, "staged", False):
            v, t, _ = cpu_run(workload, sb)
        else:
            first = {k: (v if (v is None or (k == "t_eval" and v.ndim == 1)) else v[:sb].contiguous())
                     for k, v in host.items()}
            v, t, _, ref = cpu_run(workload, sb, host=first, want_output=True)
            # the GPU solved these very samples (rows 0 .. sb-1 of its batch): every count and every ys bit
            same = (np.array_equal(local.stats["n_steps"][:sb].cpu().numpy(), ref["n_steps"])
                    and np.array_equal(local.stats["n_accepted"][:sb].cpu().numpy(), ref["n_accepted"])
                    and np.array_equal(local.status[:sb].cpu().numpy(), ref["status"])
                    and np.array_equal(local.ys[:sb].cpu().numpy(), ref["ys"], equal_nan=True))
            line["parity_checked"] = sb if same else 0
            line["parity"] = {"samples": sb, "against": "oracle port (oracle/), same seeded inputs: rows 0..n-1 of the "
                              "timed batch", "bit_exact": bool(same),
                              "compared": "n_steps, n_accepted, status, every bit of ys"}
        line["cpu_baseline_port"] = {
            "value": v, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{sb} of {B} samples of the same seeded workload, oracle port (plain C + OpenMP), {t:.1f} s"}
    except Exception as exc:
        line["cpu_baseline_port"] = {"error": str(exc)}
    from baseline import reference

    if reference.available():
        sb = REF_SAMPLE[workload.name]
        leg = reference_leg(workload, "eager", 2, 1, timeout_s=300, budget_s=30)
        if "error" not in leg:
            line["cpu_baseline"] = {
                "value": leg["value"], "unit": UNIT, "cores": leg["threads"], "kind": "reference",
                "sample": ref_sample_text(workload, sb, f"torchode 1.0.1 (baseline/_ref, unmodified) eager on the host "
                                          f"cores, {leg['ms_per_step'] / 1e3:.1f} s per solve; torch.compile(solver.solve) "
                                          "is timed by --impl reference")}
        else:
            line["cpu_baseline"] = dict(line.get("cpu_baseline_port", {}), reference_error=leg["error"])
        # torchode eager on the SAME GPU, same seeded inputs (the like-for-like competitor),
        # and this repo's result for those samples next to it
        sbc = min(REF_SAMPLE_CUDA[workload.name], B)
        try:
            first = {k: (v if (v is None or (k == "t_eval" and v.ndim == 1)) else v[:sbc].contiguous())
                     for k, v in host.items()}
            rc = time_reference(workload.name, first, str(device), "eager", 2, 1, budget_s=30)
            rsolver, rproblem = reference.build(workload.name, first, str(device))
            with torch.no_grad():
                rsol = rsolver.solve(rproblem)
            ns_same = (rsol.stats["n_steps"] == local.stats["n_steps"][:sbc]) & (
                rsol.stats["n_accepted"] == local.stats["n_accepted"][:sbc])
            err = (rsol.ys - local.ys[:sbc]).abs()
            scale = rsol.ys.abs().amax(dim=-1, keepdim=True).clamp_min(1e-30)
            line["reference_cuda"] = {
                "value": rc["value"], "unit": UNIT, "ms_per_step": rc["ms_per_step"],
                "sample": f"{sbc} of {B} samples (rows 0..n-1 of the timed batch), torchode 1.0.1 eager, device=cuda",
                "parity_vs_this_repo": {"samples": sbc, "count_mismatches": int((~ns_same).sum()),
                                        "ys_max_err_rel_to_state_norm": float((err / scale).max())}}
        except Exception as exc:
            line["reference_cuda"] = {"error": f"{type(exc).__name__}: {exc}"}
    else:
        line["cpu_baseline"] = dict(line.get("cpu_baseline_port", {}))


PER_CONFIG = ("c1", "c3", "c4", "c5")


def _main(out):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))  # c2 = BASELINE configs[1]
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU (default: the config's)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip cpu baselines / kernel rooflines / per_config")
    ap.add_argument("--ref-child", default=None, choices=["eager", "compiled", "cuda"],
                    help="internal: one leg of the reference arm (own process)")
    ap.add_argument("--ref-budget", type=float, default=None, help="internal: stop a context leg after this many s")
    ap.add_argument("--ref-compile-timeout", type=float, default=900.0,
                    help="reference arm: give up on torch.compile(solver.solve) after this many seconds")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    cls, batch = WORKLOADS[args.workload]
    if args.impl == "reference":
        workload = cls(args.workload, args.batch or batch)
        if args.ref_child:
            return _reference_child(args, workload, out)
        return run_reference_arm(args, workload, out)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference "
                         "for the CPU arm")
    env = Env()
    res, ctx = run_workload(env, args.workload, args.batch or batch, args.steps, args.warmup, main=True,
                            extras=not args.no_extras)
    line = {"metric": METRIC, "value": res.pop("value"), "unit": UNIT, "n_gpus": env.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res.pop("ms_per_step"),
            "higher_is_better": True, "scaling": res.pop("scaling"), "vs_baseline": None,
            "dtype": res.pop("dtype"), "data": "synthetic"}
    res["config"]["l2"] = "512 MiB buffer rewritten between timed steps (L2 flush)"
    line.update(res)
    if not args.no_extras and args.workload == "c2" and args.batch is None:
        # the other BASELINE configs in the same run: 3 timed steps each (2 warm-up), L2 flushed between steps;
        # N > 1: strong scaling of the config's own batch, the step ends with the full Solution on every rank
        per = {}
        for name in PER_CONFIG:
            if env.world > 1 and name == "c1":
                continue  # two samples do not shard
            try:
                torch.cuda.empty_cache()
                r, _ = run_workload(env, name, WORKLOADS[name][1], 3, 2, main=False, extras=False)
                r["metric"], r["unit"] = METRIC, UNIT
                per[name] = r
            except Exception as exc:  # a per_config failure must not take the default line down
                per[name] = {"error": f"{type(exc).__name__}: {exc}"}
            env.barrier()
        line["per_config"] = per
    if env.rank == 0 and not args.no_extras and env.world == 1:
        main_extras(env, line, ctx)
    if env.rank == 0:
        out.emit(json.dumps(line))
    if env.world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

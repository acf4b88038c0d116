"""Loads the staged reference (torchode 1.0.1, verbatim copy under baseline/_ref/) and builds the benchmark
workloads with the reference's own classes -- its public API, its stock code path
(README.md:44-56), nothing of this repo's kernels or engine on that path.  The vector fields are plain PyTorch
expressions / modules (the analytic ones are torchode_b200.fields' nn.Modules, which only use torch ops)."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "torchode", "__init__.py"))


def load():
    """The reference package (``import torchode``) from baseline/_ref; raises if it was not staged."""
    if not available():
        raise RuntimeError("the reference is not staged: run scripts/stage_reference.sh in the build container "
                           "(baseline/_ref/ is git-ignored and travels to the GPU box with the snapshot)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import torchode  # noqa: E402  (the reference; this repo's package is torchode_b200)

    assert os.path.dirname(os.path.abspath(torchode.__file__)).startswith(REF_DIR), torchode.__file__
    return torchode


def mlp_sequential(device=None):
    """configs[3]'s field as a user of the reference writes it: fp32 Sequential, default
    init under torch.manual_seed(1234), all parameters x3."""
    torch.manual_seed(1234)
    seq = torch.nn.Sequential(torch.nn.Linear(256, 256), torch.nn.Tanh(), torch.nn.Linear(256, 256),
                              torch.nn.Tanh(), torch.nn.Linear(256, 256))
    with torch.no_grad():
        for p in seq.parameters():
            p.mul_(3.0)
    return seq if device is None else seq.to(device)


def heat_field(kappa):
    def f(t, y):
        out = torch.zeros_like(y)
        out[:, 1:-1] = kappa * ((y[:, 2:] - 2 * y[:, 1:-1]) + y[:, :-2])
        return out
    return f


def build(name, host, device="cpu"):
    """(solver, problem) of workload ``name`` in the reference's classes.  ``host`` = the dict of CPU tensors
    bench.py's Workload.host_inputs returns (y0, t_start, t_end, t_eval or a shared 1-D t_eval row)."""
    to = load()
    from torchode_b200.fields import LinearDecay, LotkaVolterra, VanDerPol  # plain torch nn.Modules

    dev = torch.device(device)
    if name == "c1":
        f, method, ctrl = LinearDecay(-0.5), to.Dopri5, lambda term: to.IntegralController(1e-6, 1e-3, term=term)
    elif name == "c2":
        f, method = VanDerPol(10.0), to.Tsit5
        ctrl = lambda term: to.PIDController(1e-8, 1e-8, 0.2, 0.5, 0.0, term=term)
    elif name == "c3":
        f, method, ctrl = LotkaVolterra(), to.Dopri5, lambda term: to.IntegralController(1e-6, 1e-3, term=term)
    elif name == "c4":
        seq = mlp_sequential(dev)
        f, method, ctrl = (lambda t, y: seq(y)), to.Dopri5, lambda term: to.IntegralController(1e-6, 1e-3, term=term)
    elif name == "c5":
        f, method, ctrl = heat_field(25.0), to.Tsit5, lambda term: to.IntegralController(1e-6, 1e-3, term=term)
    else:
        raise ValueError(name)
    term = to.ODETerm(f)
    solver = to.AutoDiffAdjoint(method(term=term), ctrl(term)).to(dev)
    t_eval = host.get("t_eval")
    if t_eval is not None and t_eval.ndim == 1:
        t_eval = t_eval.expand(host["y0"].shape[0], -1)
    mv = lambda x: None if x is None else x.to(dev)
    problem = to.InitialValueProblem(y0=mv(host["y0"]), t_start=mv(host["t_start"]), t_end=mv(host["t_end"]),
                                     t_eval=mv(t_eval))
    return solver, problem

#!/usr/bin/env python
"""Reference-vs-reference noise floor of the fp32 golden cases (round 2).

For every fp32 case of tests/golden/ that is compared free-running, the REAL reference is run twice on the
stored inputs: with the field f, and with f * (1 + 2^-23) -- every f value moves by about one fp32 ulp, a
rounding-level change of the problem.  Recorded per case: the largest relative ys difference over the
samples whose step counts agree, and the fraction of samples whose counts differ.  The free-running parity
tests (tests/test_oracle_golden.py, tests/test_gpu_parity.py) assert the north star's 1e-5 and relax it, per
case, only to this measured floor.

    PYTHONPATH=baseline/_ref:. python tests/golden/measure_noise_floor.py   -> tests/golden/noise_floor.json
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import torchode as to  # the reference (baseline/_ref)  # noqa: E402

import make_golden as mg  # noqa: E402
from helpers import BENIGN, load_case  # noqa: E402

torch.set_num_threads(1)
EPS = 2.0 ** -23


def solve(case, perturb):
    f0 = mg.field_of(str(case["field"]), case["params"].tolist())
    f = (lambda t, y: f0(t, y) * (1 + EPS)) if perturb else f0
    term = to.ODETerm(f)
    spec = dict(zip(case["ctrl_keys"].tolist(), case["ctrl_vals"].tolist()))
    spec = {k: (v if k in ("kind", "norm") else float(v)) for k, v in spec.items()}
    step = {"dopri5": to.Dopri5, "tsit5": to.Tsit5}[str(case["method"])](term=term)
    ctrl = mg.make_controller(spec, term)
    ms = int(case["max_steps"])
    solver = to.AutoDiffAdjoint(step, ctrl, max_steps=None if ms < 0 else ms)
    tt = lambda k: torch.from_numpy(case[k]) if k in case else None
    with torch.no_grad():
        return solver.solve(to.InitialValueProblem(y0=tt("y0"), t_start=tt("t_start"), t_end=tt("t_end"),
                                                   t_eval=tt("t_eval")), dt0=tt("dt0"))


if __name__ == "__main__":
    out = {}
    for name in BENIGN + ["large_c3_lv_f32_B4096"]:
        case = load_case(name)
        if case["y0"].dtype != np.float32:
            continue
        a, b = solve(case, False), solve(case, True)
        ni = case["n_initialized"]
        stored_valid = np.arange(case["ys"].shape[1])[None, :, None] < ni[:, None, None]  # the rest is new_empty memory
        assert np.array_equal(np.where(stored_valid, a.ys.numpy(), 0), np.where(stored_valid, case["ys"], 0),
                              equal_nan=True), name  # the stored run is reproducible
        same = ((a.stats["n_steps"] == b.stats["n_steps"]) & (a.stats["n_accepted"] == b.stats["n_accepted"])).numpy()
        ya, yb = a.ys.numpy(), b.ys.numpy()
        n_init = np.minimum(a.stats["n_initialized"].numpy(), b.stats["n_initialized"].numpy())
        valid = (np.arange(ya.shape[1])[None, :, None] < n_init[:, None, None]) & np.isfinite(ya) & np.isfinite(yb)
        valid &= same[:, None, None]
        with np.errstate(all="ignore"):
            rel = np.where(valid, np.abs(ya - yb) / np.maximum(np.abs(ya), 1e-30), 0.0)
        out[name] = {"ys_rel_max_same_counts": float(rel.max()), "count_mismatch_fraction": float(1 - same.mean()),
                     "samples": int(same.size)}
        print(name, out[name])
    with open(os.path.join(HERE, "noise_floor.json"), "w") as fh:
        json.dump({"how": "reference vs reference with f * (1 + 2^-23), fp32 cases, torch CPU eager 1 thread",
                   "cases": out}, fh, indent=1, sort_keys=True)

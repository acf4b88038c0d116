#!/usr/bin/env python
"""Golden fixtures at sizes that reach the multi-chunk code paths, from the REAL reference (round 2):

  large_heat_f32_tsit5_F16384 / _F65536   configs[4] with rows of 4096 / 16384 16-byte vectors (B = 4): the
                                          split-mode initial step / finish and heat_step_kernel over many chunks
  large_c2_vdp_f64_B4096                  configs[1] at 4096 samples (128 warps of the fused kernel)
  large_c3_lv_f32_B4096                   configs[2] at 4096 samples, 100 shared t_eval points

Run in the build container (the reference staged by scripts/stage_reference.sh):

    PYTHONPATH=baseline/_ref:. python tests/golden/make_golden_large.py

Stored: inputs and the reference's Solution, no per-iteration trace.  CPU eager, 1 thread.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (imports the reference as `to`)
import make_golden_heat as mh  # noqa: E402

if __name__ == "__main__":
    mh.run("large_heat_f32_tsit5_F16384", 4, 16384, torch.float32, "tsit5", 11, False)
    mh.run("large_heat_f32_tsit5_F65536", 4, 65536, torch.float32, "tsit5", 12, False)

    g = torch.Generator().manual_seed(4321)
    B = 4096
    y0 = torch.rand(B, 2, generator=g, dtype=torch.float64) * 4 - 2
    mg.run_case("large_c2_vdp_f64_B4096", field="vdp", params=[10.0], method="tsit5",
                ctrl=dict(kind="pid", atol=1e-8, rtol=1e-8, pcoeff=0.2, icoeff=0.5, dcoeff=0.0), y0=y0,
                t_start=torch.zeros(B, dtype=torch.float64), t_end=torch.full((B,), 20.0, dtype=torch.float64))
    y0 = 1 + torch.rand(B, 2, generator=g)
    mg.run_case("large_c3_lv_f32_B4096", field="lv", params=[1.5, 1.0, 1.0, 3.0], method="dopri5",
                ctrl=dict(kind="integral", atol=1e-6, rtol=1e-3), y0=y0,
                t_eval=torch.linspace(0, 10, 100).repeat(B, 1))

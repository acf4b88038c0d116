"""The rounding contract depends on ptxas NOT contracting certain packed-fp32 instruction pairs.

ptxas 12.9 fuses ``mul.rn.f32x2`` + ``add.rn.f32x2`` (and ``FMUL2`` feeding an ``FFMA2`` whose multiplier is +1)
into ONE ``FFMA2`` even under ``-fmad=false`` -- one rounding where the reference's un-fused sums have two
(profiles/r01_ptxas_f32x2_contraction.txt).  The packed kernels therefore subtract negated products through
``fma(p, -1, acc)``, which this ptxas leaves alone.  The GPU parity tests would catch a change of that behaviour;
this test catches it at build time, without a GPU, on the SASS of the built objects (cuobjdump)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "torchode_b200", "csrc", "build")


def sass_of(obj, symbol_regex):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    path = os.path.join(BUILD, obj)
    if not os.path.exists(path):
        pytest.skip(f"{obj} not built (python -c 'import __graft_entry__ as g; g.build()')")
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    chunks = re.split(r"\n\s*Function : ", out)
    picked = [c for c in chunks[1:] if re.match(symbol_regex, c.split("\n", 1)[0])]
    assert picked, f"no function matching {symbol_regex} in {obj}"
    return picked


def count(sass, pattern):
    return len(re.findall(pattern, sass))


def test_dense_output_kernel_keeps_its_unfused_products():
    # solve_fused_f2_kernel<LOTKA_VOLTERRA, MINB, CK = 0>: per step 6 x 1 (field), 6 (error estimate),
    # 6 (midpoint value of the Dopri5 interpolant) + 6 (Tsit5 rows, 3 x 6) subtractions of negated products
    (sass,) = sass_of("fused_f32f32.o", r"_ZN4tode21solve_fused_f2_kernelILi2ELi\dELi0EE")
    minus_one = count(sass, r"FFMA2 R\d+, R\d+\.F32x2\.HI_LO, -1, ")
    assert minus_one >= 6 + 6 + 6 + 18, minus_one
    assert count(sass, r"\bFMUL2\b") >= 40
    # the packed stage chains themselves are fused multiply-adds by contract (runge_kutta.py:261-263)
    assert count(sass, r"\bFFMA2\b") - minus_one >= 21


def test_heat_step_kernel_keeps_its_unfused_products():
    (sass,) = sass_of("heat_step.o", r"_ZN4tode4heat16heat_step_kernelIffLi4EE")
    # error estimate: 6 subtractions of negated products per pair, two pairs
    assert count(sass, r"FFMA2 R\d+, R\d+\.F32x2\.HI_LO, -1, ") >= 12
    # the stencil's 2 c is c + c (FADD2), never a packed multiplication contracted with the subtraction
    assert count(sass, r"\bFADD2\b") >= 24


def test_tensor_core_and_tmem_instructions_are_in_the_mlp_kernel():
    sass = "\n".join(sass_of("mlp_field.o", r"_ZN4tode3mlp18mlp_tanh256_kernelILi(64|128)ELb[01]EE"))
    for op in ("UTCHMMA", "LDTM", "UTCBAR", "UTMALDG"):  # tcgen05.mma, tcgen05.ld, tcgen05.commit, TMA tile loads
        assert count(sass, op) > 0, op
    # the step-fused instantiation keeps the step's y tile in TMEM (tcgen05.st) and has no local-memory traffic
    step = "\n".join(sass_of("mlp_field.o", r"_ZN4tode3mlp18mlp_tanh256_kernelILi64ELb1EE"))
    assert count(step, "STTM") > 0
    assert count(step, "STL") == 0 and count(step, "LDL") == 0

"""The C-ABI library loads without a GPU and exports every symbol the header declares."""
import ctypes
import os
import re

from torchode_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "torchode_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tode_[a-z0-9_]+)\s*\(", text)))


def test_header_and_ctypes_mirror_agree():
    assert declared_symbols() == sorted(_cabi.PROTOTYPES)


def test_library_loads_and_exports_every_symbol():
    lib = _cabi.lib()  # raises if the .so is missing or a prototype cannot be bound
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.tode_abi_version() == _cabi.ABI_VERSION
    assert lib.tode_error_string(0) == b"success"
    assert lib.tode_scratch_elems(10, 3) >= 20


def test_struct_layouts_match_the_header():
    # sizes computed from the C declarations (LP64, natural alignment)
    assert ctypes.sizeof(_cabi.Tableau) == 16 + 8 * (7 + 49 + 7 + 7 + 21)
    assert ctypes.sizeof(_cabi.Controller) == 16 + 8 * 11 + 8 + 8  # + iter_cap (ABI 3)
    assert ctypes.sizeof(_cabi.State) == 3 * 8 + 8 + 3 * 8 + 8 + 16 * 8 + 8
    assert ctypes.sizeof(_cabi.Problem) == 3 * 8 + 8 + 4 * 8 + 8 + 8
    assert ctypes.sizeof(_cabi.SolutionOut) == 8 * 8 + 8 + 8 + 6 * 8 * _cabi.MAX_PEERS  # + peer replicas (ABI 2)


def test_argument_errors_are_reported_not_thrown():
    lib = _cabi.lib()
    tab, st = _cabi.Tableau(), _cabi.State()
    tab.n_stages = 7
    rc = lib.tode_erk_stage(ctypes.byref(tab), 1, ctypes.byref(st), _cabi.KPtrs(), None, None)
    assert rc == -1  # TODE_EINVAL: NULL pointers, no CUDA call was made
    assert b"invalid" in lib.tode_error_string(rc)

"""The C-ABI library loads and exports every symbol include/pde_b200.h declares;
the opcode numbering of header, product and oracle agree (CPU, no compute)."""
import ctypes
import os
import re

import pytest

from conftest import REPO

HEADER = os.path.join(REPO, "include", "pde_b200.h")


def _header_text():
    with open(HEADER) as f:
        return f.read()


def test_library_exports_every_declared_symbol():
    from pde_engine_b200 import _lib
    txt = re.sub(r"/\*.*?\*/", "", _header_text(), flags=re.S)
    declared = set(re.findall(r"\b(pde_[a-z0-9_]+)\s*\(", txt))
    declared -= {"pde_validate_out"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTED)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.lib.pde_abi_version() == 6


def test_opcode_tables_agree():
    from oracle import bytecode as bc
    txt = _header_text()
    enum = dict((k, int(v, 16)) for k, v in re.findall(r"(PDE_OP_[A-Z0-9_]+)\s*=\s*(0x[0-9A-Fa-f]+)", txt))
    want = {
        "PDE_OP_END": bc.OP_END, "PDE_OP_VAR0": bc.OP_VAR0, "PDE_OP_VAR1": bc.OP_VAR1, "PDE_OP_PRIM0": bc.OP_PRIM0,
        "PDE_OP_ADD": bc.OP_ADD, "PDE_OP_SUB": bc.OP_SUB, "PDE_OP_MUL": bc.OP_MUL, "PDE_OP_DIV": bc.OP_DIV,
        "PDE_OP_NEG": bc.OP_NEG, "PDE_OP_ABS": bc.OP_ABS, "PDE_OP_SQRT": bc.OP_SQRT, "PDE_OP_EXP": bc.OP_EXP,
        "PDE_OP_FN_NEG": bc.OP_FN_NEG, "PDE_OP_FN_INV": bc.OP_FN_INV, "PDE_OP_FN_SQUARE": bc.OP_FN_SQUARE,
        "PDE_OP_FN_POW32": bc.OP_FN_POW32, "PDE_OP_FN_POWN32": bc.OP_FN_POWN32, "PDE_OP_FN_EXPNEG": bc.OP_FN_EXPNEG,
        "PDE_OP_POW0": bc.OP_POW0, "PDE_OP_CONST0": bc.OP_CONST0,
    }
    assert enum == want
    defs = dict(re.findall(r"#define\s+(PDE_N_[A-Z]+)\s+(\d+)", txt))
    assert (int(defs["PDE_N_PRIM"]), int(defs["PDE_N_POW"]), int(defs["PDE_N_CONST"])) == (bc.N_PRIM, bc.N_POW, bc.N_CONST)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point fails loudly (PDE_E_NODEVICE)."""
    import pde_engine_b200 as pb
    from pde_engine_b200 import _lib
    if pb.device_count() > 0:
        pytest.skip("a GPU is visible")
    sess = pb.Session.for_problem("force_free")
    prog = pb.ResidualProgram.for_problem("force_free")
    out = _lib.ValidateOut()
    rc = _lib.lib.pde_validate(sess._h, prog._h, None, None, 0, 48, None, None, None, 0, 64, 1e-10, 8, 0.5, 0.0625, 0, 3, 4,
                               ctypes.byref(out), None)
    assert rc == _lib.PDE_E_NODEVICE
    assert b"no CPU fallback" in _lib.lib.pde_last_error()
    n = ctypes.c_int64()
    es = sess.compile(["rho", "z"])
    db = (ctypes.c_int32 * 2)(0, 2)
    assert _lib.lib.pde_enumerate_count(es._h, db, 2, 1, ctypes.byref(n), None) == _lib.PDE_E_NODEVICE


def test_product_never_imports_oracle():
    """The product path must not route through the oracle."""
    pkg = os.path.join(REPO, "pde_engine_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f

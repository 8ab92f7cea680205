"""ctypes binding of libpde_b200.so (the C-ABI declared in include/pde_b200.h).

There is NO CPU fallback: if the shared library is missing the import fails
loudly with build instructions, and every compute entry point of the library
returns PDE_E_NODEVICE when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PDE_B200_LIB") or os.path.join(_HERE, "libpde_b200.so")   # env override: kernel-variant experiments


class PdeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libpde_b200 error {code}: {msg}")
        self.code = code


PDE_E_NODEVICE = -5

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: pde_engine_b200 is a CUDA-only implementation (no CPU fallback). "
        "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
        "`pde_engine_b200/csrc/build.sh` (nvcc, sm_100a).")

lib = C.CDLL(LIB_PATH)

c_void_p, c_int, c_int64, c_double, c_char_p = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_char_p
P = C.POINTER


class ValidateOut(C.Structure):
    _fields_ = [
        ("ratio_max", c_void_p), ("resid_max", c_void_p), ("scale_at", c_void_p),
        ("n_finite", c_void_p), ("n_votes", c_void_p), ("ref_rs", c_void_p), ("survivor_bits", c_void_p),
        ("confirm", c_void_p), ("scratch", c_void_p),
    ]


def _sig(name, restype, *argtypes):
    f = getattr(lib, name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


_sig("pde_abi_version", c_int)
_sig("pde_last_error", c_char_p)
_sig("pde_device_count", c_int)
_sig("pde_launch_count", c_int64)
_sig("pde_session_create", c_int, c_char_p, c_char_p, P(c_char_p), P(c_double), c_int, P(c_void_p))
_sig("pde_session_free", None, c_void_p)
_sig("pde_session_tables", c_int, c_void_p, P(c_double), P(c_int), P(c_double), P(c_int))
_sig("pde_session_const_key", c_char_p, c_void_p, c_int)
_sig("pde_session_pow_key", c_char_p, c_void_p, c_int)
_sig("pde_compile_exprs", c_int, c_void_p, P(c_char_p), c_int, P(c_void_p))
_sig("pde_compile_exprs_packed", c_int, c_void_p, c_void_p, c_void_p, c_int, P(c_void_p))
_sig("pde_compile_exprs_blob", c_int, c_void_p, c_void_p, C.c_size_t, c_int, P(c_void_p))
_sig("pde_exprset_free", None, c_void_p)
_sig("pde_exprset_size", c_int, c_void_p, P(c_int), P(c_int), P(c_int))
_sig("pde_exprset_export", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
_sig("pde_exprset_programs", c_int, c_void_p, c_int, c_void_p, c_void_p)
_sig("pde_enumerate_count", c_int, c_void_p, P(C.c_int32), c_int, c_int, P(c_int64), c_void_p)
_sig("pde_enumerate", c_int, c_void_p, P(C.c_int32), c_int, c_int, c_int64, c_int64, c_int,
     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
_sig("pde_dedup", c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, P(c_int64), c_void_p)
_sig("pde_enumerate_csr_size", c_int, c_void_p, P(C.c_int32), c_int, c_int, c_int64, c_int64, c_int, P(c_int64), c_void_p)
_sig("pde_enumerate_csr", c_int, c_void_p, P(C.c_int32), c_int, c_int, c_int64, c_int64, c_int,
     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
_sig("pde_dedup_csr", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, P(c_int64), c_void_p)
_sig("pde_synth_trees", c_int, C.c_uint64, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p)
_sig("pde_compile_residual", c_int, c_int, P(c_double), c_int, P(c_void_p))
_sig("pde_compile_residual_program", c_int, c_int, c_int, P(c_double), c_int, P(C.c_uint32), c_int, P(c_void_p))
_sig("pde_program_free", None, c_void_p)
_sig("pde_program_info", c_int, c_void_p, P(c_int), P(c_int), P(c_int))
_sig("pde_program_point_table", c_int, c_void_p, c_void_p, c_int, c_void_p)
_sig("pde_validate", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
     c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_int, c_double, c_double, c_int, c_int, c_int, P(ValidateOut), c_void_p)
_sig("pde_validate_csr", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
     c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_int, c_double, c_double, c_int, c_int, c_int, P(ValidateOut), c_void_p)
_sig("pde_eval_points", c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
     c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p)
_sig("pde_fingerprint", c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_int,
     c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p)
_sig("pde_fp64_peak", c_int, c_int, P(c_double), c_void_p)
_sig("pde_fp64_peak_3op", c_int, c_int, P(c_double), c_void_p)

EXPORTED = [
    "pde_abi_version", "pde_last_error", "pde_device_count", "pde_launch_count",
    "pde_session_create", "pde_session_free", "pde_session_tables", "pde_session_const_key", "pde_session_pow_key",
    "pde_compile_exprs", "pde_compile_exprs_packed", "pde_compile_exprs_blob", "pde_exprset_free", "pde_exprset_size", "pde_exprset_export", "pde_exprset_programs",
    "pde_enumerate_count", "pde_enumerate", "pde_dedup", "pde_enumerate_csr_size", "pde_enumerate_csr", "pde_dedup_csr", "pde_synth_trees",
    "pde_compile_residual", "pde_compile_residual_program", "pde_program_free", "pde_program_info", "pde_program_point_table",
    "pde_validate", "pde_validate_csr", "pde_eval_points", "pde_fingerprint", "pde_fp64_peak", "pde_fp64_peak_3op",
]


def check(rc: int) -> None:
    if rc != 0:
        raise PdeError(rc, (lib.pde_last_error() or b"").decode(errors="replace"))

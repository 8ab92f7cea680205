"""pde_engine_b200 -- B200-native hot path of pde-engine (enumerate + validate).

CUDA only: importing this package loads libpde_b200.so and fails loudly when it
is missing.  See DESIGN.md / INTEGRATION.md.
"""
from . import _lib  # noqa: F401  (loads the shared library or raises ImportError)
from .core import (Session, ExprSet, ResidualProgram, validate, eval_points, enumerate_count,
                   enumerate_candidates, enumerate_candidates_csr, csr_rows, dedup, dedup_csr, synth_trees, fp64_peak, fp64_peak_3op, device_count, launch_count,
                   PROBLEM_FORCE_FREE, PROBLEM_KERR)

__all__ = ["Session", "ExprSet", "ResidualProgram", "validate", "eval_points", "enumerate_count",
           "enumerate_candidates", "enumerate_candidates_csr", "csr_rows", "dedup", "dedup_csr", "synth_trees", "fp64_peak", "fp64_peak_3op", "device_count", "launch_count",
           "PROBLEM_FORCE_FREE", "PROBLEM_KERR"]

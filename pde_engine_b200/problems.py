"""Problem plugins -- host-side mirror of the reference's problems/__init__.py.

Same dataclass fields (PI:34-63), same primitives / symbols / constants
(PI:66-108, 259-302), same ``load_problem`` aliases and error (PI:355-361), so
code written against the reference's plugin API runs unchanged.  The only
difference: ``spec.validator`` is a ``GpuBatchValidator`` that filters on the
device and delegates survivors to the CPU validator you pass in (in production
the reference's own validator object; ``None`` = filter only).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional

import sympy as sp

from .generator import UNARY_NAMES, BINARY_NAMES, DEAD_BINARY_NAMES


def _unary_ops() -> Dict[str, Callable]:
    # expression_operations.py:30-60 (semantics of every unary opcode)
    return {
        "neg": lambda x: -x,
        "inv": lambda x: 1 / x,
        "sqrt": lambda x: sp.sqrt(x),
        "square": lambda x: x ** 2,
        "pow_3_2": lambda x: x ** sp.Rational(3, 2),
        "pow_neg_3_2": lambda x: x ** sp.Rational(-3, 2),
        "exp": lambda x: sp.exp(x),
        "exp_neg": lambda x: sp.exp(-x),
    }


def _binary_ops() -> Dict[str, Callable]:
    # expression_operations.py:11-28
    return {
        "add": lambda x, y: x + y,
        "sub": lambda x, y: x - y,
        "mul": lambda x, y: x * y,
        "div": lambda x, y: x / y,
        "geom_sum": lambda x, y: x / (1 - y),
    }


def _special_ops() -> Dict[str, Callable]:
    # expression_operations.py:62-77 -- present in the vocabulary, dead in the live generator (LBF:170-195)
    return {
        "sqrt_shift_neg": lambda x, y: sp.sqrt((x - 1) ** 2 + y ** 2),
        "sqrt_shift_pos": lambda x, y: sp.sqrt((x + 1) ** 2 + y ** 2),
        "exp_mul": lambda x, y: x * sp.exp(y),
        "log_mul": lambda x, y: x * sp.log(y),
    }


assert tuple(_unary_ops()) == UNARY_NAMES and tuple(_binary_ops()) == BINARY_NAMES and tuple(_special_ops()) == DEAD_BINARY_NAMES


@dataclass
class ProblemSpec:
    """Specification container for a PDE discovery problem (PI:34-63)."""
    name: str
    slug: str
    symbols: Dict[str, sp.Symbol]
    constants: Dict[str, sp.Symbol]
    primitives: List[sp.Basic]
    unary_ops: Dict[str, Callable]
    binary_ops: Dict[str, Callable]
    special_ops: Dict[str, Callable]
    all_binary_ops: Dict[str, Callable]
    validator: Any
    known_solutions: Dict[str, str]
    output_root: str

    def get_output_dir(self) -> str:
        os.makedirs(self.output_root, exist_ok=True)
        return self.output_root

    def sympify_locals(self) -> dict:
        """The locals mapping the engine parses candidate strings with (GM:85-93)."""
        locs: dict = {}
        locs.update(self.symbols)
        locs.update(self.constants)
        locs.update(self.unary_ops)
        return locs


def _spec(name, slug, symbols, constants, primitives, validator, known, make_gpu, gpu_kwargs) -> ProblemSpec:
    spec = ProblemSpec(
        name=name, slug=slug, symbols=symbols, constants=constants, primitives=primitives,
        unary_ops=_unary_ops(), binary_ops=_binary_ops(), special_ops=_special_ops(),
        all_binary_ops={**_binary_ops(), **_special_ops()}, validator=validator, known_solutions=known,
        output_root=os.path.join("problems", slug, "outputs"))
    if make_gpu:
        from .validator import GpuBatchValidator
        spec.validator = GpuBatchValidator(validator, slug, sympify_locals=spec.sympify_locals(), **gpu_kwargs)
    return spec


def _create_force_free_problem(cpu_validator=None, make_gpu=True, **gpu_kwargs) -> ProblemSpec:
    rho = sp.Symbol("rho", real=True, positive=True)     # PI:70-71
    z = sp.Symbol("z", real=True)
    primitives = [rho, z, rho ** 2 + z ** 2, rho / z, sp.Integer(1)]   # PI:73-79
    known = {                                            # PI:85-93
        "rho**2": "Vertical field",
        "rho**2*z": "X-point",
        "1 - z/sqrt(rho**2 + z**2)": "Radial",
        "rho**2/(rho**2 + z**2)**(3/2)": "Dipolar",
        "sqrt(rho**2 + z**2) - z": "Parabolic",
        "sqrt(z**2 + (rho - 1)**2) - sqrt(z**2 + (rho + 1)**2)": "Hyperbolic",
        "rho**2*exp(-2*z)": "Bent",
    }
    return _spec("Force-Free Foliations", "force_free", {"rho": rho, "z": z}, {}, primitives,
                 cpu_validator, known, make_gpu, gpu_kwargs)


def _create_kerr_magnetosphere_problem(cpu_validator=None, make_gpu=True, **gpu_kwargs) -> ProblemSpec:
    r = sp.Symbol("r", real=True, positive=True)         # PI:263-266
    x = sp.Symbol("x", real=True)
    M = sp.Symbol("M", real=True, positive=True)
    a = sp.Symbol("a", real=True)
    Delta = r ** 2 - 2 * M * r + a ** 2
    G = 1 - (2 * M * r) / (r ** 2 + a ** 2 * x ** 2)
    primitives = [r, x, sp.Integer(1), sp.Rational(1, 3), (1 - x), a ** 2, r ** 2 + a ** 2 * x ** 2, Delta, G]   # PI:271-281
    return _spec("Kerr Magnetosphere (linear surrogate)", "kerr_magnetosphere", {"r": r, "x": x}, {"M": M, "a": a},
                 primitives, cpu_validator, {"1 - x": "Monopole (a -> 0 limit)"}, make_gpu, gpu_kwargs)


class SymbolicResidualValidator:
    """CPU confirmation for a custom plugin: the validator protocol (PI:52) over `lhs(u)` -- valid iff the residual
    simplifies to zero identically (what KerrMagnetosphereValidator.validate does with its own `_lhs`, KV:284-292)."""

    def __init__(self, lhs: Callable, coords):
        self._lhs_fn, self.coords = lhs, tuple(coords)

    def _lhs(self, u):
        return self._lhs_fn(u)

    def validate(self, u, check_regularity: bool = True, fast_point_only: bool = False, **kw):
        try:
            if not any(sp.sympify(u).has(c) for c in self.coords):
                return False, "Trivial constant solution excluded"
            res = sp.simplify(self._lhs_fn(u).doit())
            return (True, "Valid (exact zero)") if res == 0 else (False, f"PDE residual != 0 | residual: {str(res)[:80]}")
        except Exception as e:   # KV:344-345
            return False, f"Validation error: {e}"

    def describe(self):
        return {"method_name": "pde_engine_b200.problems.SymbolicResidualValidator.validate", "math_definition": "lhs(u) = 0"}


def custom_problem(name: str, slug: str, lhs: Callable, base: str = "force_free", primitives: Optional[List] = None,
                   known_solutions: Optional[Dict[str, str]] = None, cpu_validator: Any = "symbolic",
                   params: Optional[dict] = None, **gpu_kwargs) -> ProblemSpec:
    """A new plugin with NO CUDA and no rebuild (BASELINE north_star item 3).  `lhs(u)` states the PDE like the
    reference's `KerrMagnetosphereValidator._lhs` (KV:77-91): it is called once with a generic `Function('u')` of the
    base problem's coordinates and compiled into a device residual program (residual_compiler.compile_residual);
    the ProblemSpec keeps the reference's fields (PI:34-63).  `base` supplies the coordinate system, constants,
    default primitives and the collocation grid; `params` gives numeric values for symbolic constants in lhs."""
    from .residual_compiler import compile_residual
    from .validator import GpuBatchValidator
    b = load_problem(base, make_gpu=False)
    coords = tuple(b.symbols.values())
    u = sp.Function("u")(*coords)
    cr = compile_residual(lhs(u), u, coords, params=params, description=name)
    if cpu_validator == "symbolic":
        cpu_validator = SymbolicResidualValidator(lhs, coords)
    spec = ProblemSpec(
        name=name, slug=slug, symbols=b.symbols, constants=b.constants,
        primitives=list(primitives) if primitives is not None else b.primitives,
        unary_ops=_unary_ops(), binary_ops=_binary_ops(), special_ops=_special_ops(),
        all_binary_ops={**_binary_ops(), **_special_ops()}, validator=cpu_validator,
        known_solutions=dict(known_solutions or {}), output_root=os.path.join("problems", slug, "outputs"))
    spec.coordinate_system = b.slug
    spec.compiled_residual = cr
    spec.validator = GpuBatchValidator(cpu_validator, b.slug, sympify_locals=spec.sympify_locals(), program=cr.program(),
                                       math_definition=f"{sp.sstr(lhs(u))} = 0", **gpu_kwargs)
    return spec


def load_problem(name: str, cpu_validator=None, make_gpu: bool = True, **gpu_kwargs) -> ProblemSpec:
    """PI:355-361 (same aliases, same error)."""
    key = (name or "").strip().lower()
    if key in ("force_free", "forcefree", "foliation", "foliations"):
        return _create_force_free_problem(cpu_validator, make_gpu, **gpu_kwargs)
    if key in ("kerr", "kerr_magnetosphere", "kerr-magnetosphere"):
        return _create_kerr_magnetosphere_problem(cpu_validator, make_gpu, **gpu_kwargs)
    raise ValueError(f"Unknown problem '{name}'. Available: 'force_free', 'kerr_magnetosphere'")


__all__ = ["ProblemSpec", "load_problem", "custom_problem", "SymbolicResidualValidator"]

"""Run-time residual programs: a plugin's PDE, stated as a SymPy formula over a generic u, compiled into the
straight-line device program of `pde_compile_residual_program` (include/pde_b200.h) -- no CUDA, no rebuild.

The reference's plugin seam is `ProblemSpec.validator` (problems/__init__.py:34-63); its validators state the PDE
symbolically: `KerrMagnetosphereValidator._lhs(u)` accepts any u, including `Function('u')(r, x)`
(problems/kerr_magnetosphere/validator.py:77-91), the force-free one builds det M from Lie derivatives
(problems/force_free/validator.py:305-347).  `compile_residual` takes such a formula:

    r, x = sp.symbols("r x"); u = sp.Function("u")(r, x)
    cr = compile_residual(validator._lhs(u), u, (r, x), params={M: 1, a: sp.Rational(1, 10)})
    prog = cr.program()                      # core.ResidualProgram (device handle)
    table = cr.point_table(pts)              # [n_cols, P] float64: the coefficient functions at the grid

Requirements: the formula is a POLYNOMIAL in u and its partial derivatives (order <= 4) whose coefficients depend
only on the coordinates (and numeric parameters).  Each distinct coefficient function becomes a column of the point
table (evaluated on the host in float64 by numpy), each numeric factor a program constant.

`ProgramBuilder` is the assembler underneath (virtual registers -> file slots by liveness); tools/gen_residual.py
uses it to emit the two built-in residuals in the exact schedule of their CUDA specialisations (residual_programs.py),
which is how the tests show the interpreter and the specialisations agree bit for bit.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

# include/pde_b200.h: enum pde_residual_op, PDE_R_*
R_END, R_MUL, R_ACC0, R_ACC, R_LDA, R_STA, R_ADDA, R_OUT = range(8)
R_MAX_WORDS, R_MAX_CONSTS, R_MAX_COLS, R_MAX_FILE = 2048, 16, 8, 51


def jidx(i: int, j: int) -> int:
    return (i + j) * (i + j + 1) // 2 + j


def n_coef(order: int) -> int:
    return (order + 1) * (order + 2) // 2


def word(op: int, a: int = 0, b: int = 0, dst: int = 0, neg: bool = False) -> int:
    return op | (a << 4) | (b << 12) | (dst << 20) | ((1 if neg else 0) << 28)


def decode(w: int) -> Tuple[int, int, int, int, bool]:
    return w & 15, (w >> 4) & 255, (w >> 12) & 255, (w >> 20) & 255, bool((w >> 28) & 1)


class ProgramBuilder:
    """Assembler for residual programs.  Operands are ("d", g) | ("col", k) | ("const", k) | a virtual register
    returned by `mul` / `sta`.  Products are scheduled lazily (right before their first use) and file slots are
    assigned by liveness, so the file stays small; the arithmetic (which products and fused multiply-adds are
    formed, in which order they are accumulated) is exactly what the caller spelled out."""

    def __init__(self, order: int, n_cols: int, consts: Sequence[float] = ()):
        assert order in (2, 4), "jets of order 2 and 4 are instantiated on the device"
        assert 0 <= n_cols <= R_MAX_COLS
        self.order, self.n_cols = order, n_cols
        self.consts: List[float] = [float(c) for c in consts]
        self._muls: Dict[int, Tuple] = {}          # vreg -> (x, y) not yet emitted
        self._ops: List[Tuple] = []                # (op, x, y, dst_vreg, neg)
        self._next = 0

    # ---- operands -------------------------------------------------------
    def d(self, i: int, j: int):
        assert 0 <= i + j <= self.order
        return ("d", jidx(i, j))

    def col(self, k: int):
        assert 0 <= k < self.n_cols
        return ("col", k)

    def const(self, value: float):
        value = float(value)
        if value not in self.consts:
            self.consts.append(value)
        assert len(self.consts) <= R_MAX_CONSTS, "too many program constants"
        return ("const", self.consts.index(value))

    def _vreg(self):
        self._next += 1
        return ("v", self._next)

    def _use(self, x):
        if x[0] == "v" and x[1] in self._muls:        # emit the product now (and what it needs)
            a, b = self._muls.pop(x[1])
            self._use(a)
            self._use(b)
            self._ops.append((R_MUL, a, b, x, False))

    # ---- instructions ---------------------------------------------------
    def mul(self, x, y):
        v = self._vreg()
        self._muls[v[1]] = (x, y)
        return v

    def acc0(self, x, y, neg=False):
        self._use(x); self._use(y)
        self._ops.append((R_ACC0, x, y, None, neg))

    def acc(self, x, y, neg=False):
        self._use(x); self._use(y)
        self._ops.append((R_ACC, x, y, None, neg))

    def lda(self, x, neg=False):
        self._use(x)
        self._ops.append((R_LDA, x, None, None, neg))

    def adda(self, x, neg=False):
        self._use(x)
        self._ops.append((R_ADDA, x, None, None, neg))

    def sta(self):
        v = self._vreg()
        self._ops.append((R_STA, None, None, v, False))
        return v

    def out(self):
        self._ops.append((R_OUT, None, None, None, False))

    # ---- assembly -------------------------------------------------------
    def assemble(self) -> Tuple[List[int], int]:
        """(words, n_file).  Raises ValueError when the program does not fit the device limits."""
        nc = n_coef(self.order)
        base = nc + self.n_cols + len(self.consts)
        ops = self._ops
        if not ops or ops[-1][0] != R_OUT:
            raise ValueError("a residual program ends with out()")
        last = {}
        for k, (_, x, y, dst, _) in enumerate(ops):
            for o in (x, y):
                if o is not None and o[0] == "v":
                    last[o[1]] = k
        slot_of, free, top = {}, [], base

        def slot(o):
            if o is None:
                return 0
            if o[0] == "d":
                return o[1]
            if o[0] == "col":
                return nc + o[1]
            if o[0] == "const":
                return nc + self.n_cols + o[1]
            return slot_of[o[1]]

        words = []
        for k, (op, x, y, dst, neg) in enumerate(ops):
            a, b = slot(x), slot(y)
            for o in (x, y):                           # operands that die here free their slot for this op's result
                if o is not None and o[0] == "v" and last.get(o[1]) == k and o[1] in slot_of:
                    free.append(slot_of.pop(o[1]))
            dd = 0
            if dst is not None:
                if dst[1] not in last:                  # never read: still needs a place to be written
                    last[dst[1]] = k
                if free:
                    dd = min(free); free.remove(dd)
                else:
                    dd = top; top += 1
                slot_of[dst[1]] = dd
                if last[dst[1]] == k:
                    free.append(slot_of.pop(dst[1]))
            words.append(word(op, a, b, dd, neg))
        if top > R_MAX_FILE:
            raise ValueError(f"residual program needs a file of {top} > {R_MAX_FILE} entries")
        if len(words) > R_MAX_WORDS:
            raise ValueError(f"residual program has {len(words)} > {R_MAX_WORDS} words")
        return words, top


def interpret(words: Sequence[int], order: int, n_cols: int, consts: Sequence[float], d: np.ndarray,
              cols: np.ndarray, magnitudes: bool = False, theta: Optional[np.ndarray] = None) -> np.ndarray:
    """Host restatement of the device interpreter (validate.cuh: res_run) on arrays of points -- used by the CPU tests
    and by `CompiledResidual.check`.  d [n_coef, P] partial derivatives, cols [n_cols, P]; magnitudes=True runs the
    scale pass (|d_g| + theta[|g|], signs dropped).  Products and sums are separate roundings; ACC is evaluated as
    x*y + acc in float64 (numpy has no fma): equal to the device up to one rounding per fused multiply-add."""
    nc = n_coef(order)
    P = d.shape[1]
    F = np.zeros((256, P))
    if magnitudes:
        th = np.zeros(order + 1) if theta is None else np.asarray(theta, float)
        for n in range(order + 1):
            for j in range(n + 1):
                F[jidx(n - j, j)] = np.abs(d[jidx(n - j, j)]) + (th[n] if th.ndim == 1 else th[n])
        F[nc:nc + n_cols] = np.abs(cols)
        for k, c in enumerate(consts):
            F[nc + n_cols + k] = abs(c)
    else:
        F[:nc] = d
        F[nc:nc + n_cols] = cols
        for k, c in enumerate(consts):
            F[nc + n_cols + k] = c
    acc = np.zeros(P)
    with np.errstate(all="ignore"):
        for w in words:
            op, a, b, dst, neg = decode(w)
            if op in (R_OUT, R_END):
                break
            if op == R_STA:
                F[dst] = acc
                continue
            x = -F[a] if (neg and not magnitudes) else F[a]
            if op == R_LDA:
                acc = x.copy()
            elif op == R_ADDA:
                acc = acc + x
            elif op == R_MUL:
                F[dst] = x * F[b]
            elif op == R_ACC0:
                acc = x * F[b]
            else:
                acc = x * F[b] + acc
    return acc


@dataclass
class CompiledResidual:
    order: int
    n_cols: int
    consts: List[float]
    words: List[int]
    n_file: int
    columns: list = field(default_factory=list)              # SymPy expressions of the coefficient functions
    coords: tuple = ()
    _table_fn: Optional[Callable] = None
    description: str = ""

    def point_table(self, pts: np.ndarray) -> np.ndarray:
        """pts [2, P] (SoA) -> [max(n_cols, 1), P] float64."""
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        P = pts.shape[1]
        tab = np.zeros((max(self.n_cols, 1), P))
        if self.n_cols:
            with np.errstate(all="ignore"):
                vals = self._table_fn(pts[0], pts[1])
            for k, v in enumerate(vals):
                tab[k] = np.broadcast_to(np.asarray(v, dtype=np.float64), (P,))
        return tab

    def program(self):
        from . import core
        return core.ResidualProgram.from_words(self.order, self.n_cols, self.consts, self.words, table_fn=self.point_table)


def compile_residual(lhs, u, coords, params: Optional[dict] = None, description: str = "") -> CompiledResidual:
    """Compile `lhs` (SymPy, in u = Function(...)(x0, x1) and its derivatives) into a device residual program."""
    import sympy as sp
    x0, x1 = coords
    expr = sp.sympify(lhs).doit()
    if params:
        expr = expr.subs(params)
    # u and its partial derivatives -> polynomial generators D_i_j
    rep, gens, ij_of = {}, [], {}

    def gen(i, j):
        key = (i, j)
        if key not in ij_of:
            s = sp.Symbol(f"D_{i}_{j}")
            ij_of[key] = s
            gens.append(s)
        return ij_of[key]

    for der in expr.atoms(sp.Derivative):
        if der.expr != u:
            raise ValueError(f"derivative of something other than u: {der}")
        cnt = {v: int(k) for v, k in der.variable_count}
        if set(cnt) - {x0, x1}:
            raise ValueError(f"derivative with respect to a non-coordinate: {der}")
        rep[der] = gen(cnt.get(x0, 0), cnt.get(x1, 0))
    expr = expr.xreplace(rep)
    if expr.has(u):
        expr = expr.xreplace({u: gen(0, 0)})
    if expr.atoms(sp.Function) - expr.atoms(sp.exp, sp.log, sp.sin, sp.cos, sp.Abs, sp.sign) or any(
            isinstance(f, sp.core.function.AppliedUndef) for f in expr.atoms(sp.Function)):
        raise ValueError("the residual still contains an undefined function after substituting u's derivatives")
    max_order = max((i + j for i, j in ij_of), default=0)
    if max_order > 4:
        raise ValueError(f"partial derivatives of order {max_order} > 4")
    order = 2 if max_order <= 2 else 4
    gens.sort(key=lambda s: tuple(int(t) for t in s.name.split("_")[1:]))
    expr = sp.together(sp.expand(expr))
    num, den = sp.fraction(expr)
    if any(den.has(g) for g in gens):
        raise ValueError("the residual is not a polynomial in the derivatives of u")
    if not gens:
        raise ValueError("the residual does not involve u")
    try:
        poly = sp.Poly(sp.expand(num), *gens)
    except sp.PolynomialError as e:
        raise ValueError(f"the residual is not a polynomial in the derivatives of u: {e}") from None
    # coefficient = rational number * function of the point; distinct functions become table columns
    columns, terms = [], []
    for mon, coef in zip(poly.monoms(), poly.coeffs()):
        c = sp.simplify(coef / den)
        if c.free_symbols - {x0, x1}:
            raise ValueError(f"coefficient depends on something other than the coordinates: {c.free_symbols - {x0, x1}}")
        k, f = c.as_coeff_Mul()
        kf = float(k)
        col = None
        if f != 1:
            f = sp.simplify(f)
            for q, g in enumerate(columns):
                if sp.simplify(g - f) == 0:
                    col = q
                    break
            if col is None:
                columns.append(f)
                col = len(columns) - 1
        terms.append((kf, col, mon))
    if len(columns) > R_MAX_COLS:
        raise ValueError(f"{len(columns)} distinct coefficient functions > {R_MAX_COLS} table columns")
    b = ProgramBuilder(order, len(columns))
    gen_ij = [tuple(int(t) for t in s.name.split("_")[1:]) for s in gens]
    memo: Dict[tuple, tuple] = {}

    def product(factors):
        """Memoised left-to-right product chain of >= 1 operands."""
        if len(factors) == 1:
            return factors[0]
        key = tuple(factors)
        if key not in memo:
            memo[key] = b.mul(product(factors[:-1]), factors[-1])
        return memo[key]

    first = True
    for kf, col, mon in terms:
        factors = []
        if abs(kf) != 1.0:
            factors.append(b.const(abs(kf)))
        if col is not None:
            factors.append(b.col(col))
        for (i, j), e in zip(gen_ij, mon):
            factors += [b.d(i, j)] * e
        neg = kf < 0
        if len(factors) == 1:
            (b.lda if first else b.adda)(factors[0], neg)
        else:
            left, right = product(factors[:-1]), factors[-1]
            (b.acc0 if first else b.acc)(left, right, neg)
        first = False
    b.out()
    words, n_file = b.assemble()
    table_fn = sp.lambdify((x0, x1), columns, "numpy") if columns else None
    return CompiledResidual(order, len(columns), b.consts, words, n_file, columns, (x0, x1), table_fn, description)


__all__ = ["ProgramBuilder", "CompiledResidual", "compile_residual", "interpret", "word", "decode", "jidx", "n_coef"]

"""Oracle: expression string -> term-structured postfix bytecode (TEST INFRASTRUCTURE).

The reference turns strings into expressions with ``sympy.sympify``
(general_method_paper_reproduction.py:1257, lean_normalizer/lean_bridge.py:73),
i.e. with *Python's* expression grammar.  This restatement therefore uses
Python's own ``ast`` module as the parser, so operator precedence
(``**`` > unary minus > ``* /`` > ``+ -``, ``**`` right-associative) is Python's
by construction.  The product's C++ compiler (pde_engine_b200/csrc/compiler.cpp)
is a hand-written recursive-descent parser and is tested byte-for-byte
against this module.

Compilation pipeline (identical in the product):
  1. parse -> IR; fold sub-trees made only of integer literals and
     ``+ - * / **`` into exact rationals (|num|, den <= 2**53, else UNSUPPORTED);
  2. split the top level into additive terms along the LEFT spine of the
     ``+``/``-`` chain (what the reference's textual splice sees, LBF:170-195);
     a leading unary minus of a term (leftmost leaf of its ``* /`` chain)
     becomes the term's sign;
  3. emit each term body as postfix; nested sums are emitted by the same rule
     ("whole" form:  t1 [NEG]  (tk ADD|SUB)* ).
"""
from __future__ import annotations

import ast
import math
from dataclasses import dataclass, field
from fractions import Fraction
from typing import Dict, List, Optional, Sequence, Tuple

from . import bytecode as bc

LIMIT = 1 << 53


class Unsupported(Exception):
    pass


class TableFull(Exception):
    pass


@dataclass
class Session:
    """Per-problem symbol / constant / exponent tables (append-only)."""
    var_names: Tuple[str, str] = ("rho", "z")
    named_consts: Dict[str, float] = field(default_factory=dict)
    const_keys: List[str] = field(default_factory=lambda: ["1"])
    const_vals: List[float] = field(default_factory=lambda: [1.0])
    pow_keys: List[str] = field(default_factory=lambda: ["3/2", "-3/2", "2"])
    pow_vals: List[float] = field(default_factory=lambda: [1.5, -1.5, 2.0])

    def __post_init__(self):
        self.named_consts = dict(self.named_consts)
        self.named_consts.setdefault("E", math.e)

    @staticmethod
    def for_problem(slug: str) -> "Session":
        if slug in ("force_free",):
            return Session(("rho", "z"), {})
        if slug in ("kerr_magnetosphere", "kerr"):
            # M_value = 1, a_value = 1/10  (problems/__init__.py:283)
            return Session(("r", "x"), {"M": 1.0, "a": 0.1})
        raise ValueError(slug)

    def const_slot(self, key: str, val: float) -> int:
        try:
            return self.const_keys.index(key)
        except ValueError:
            if len(self.const_keys) >= bc.N_CONST:
                raise TableFull("const")
            self.const_keys.append(key)
            self.const_vals.append(val)
            return len(self.const_keys) - 1

    def pow_slot(self, key: str, val: float) -> int:
        try:
            return self.pow_keys.index(key)
        except ValueError:
            if len(self.pow_keys) >= bc.N_POW:
                raise TableFull("pow")
            self.pow_keys.append(key)
            self.pow_vals.append(val)
            return len(self.pow_keys) - 1


def frac_key(f: Fraction) -> str:
    return str(f.numerator) if f.denominator == 1 else f"{f.numerator}/{f.denominator}"


def _check(f: Fraction) -> Fraction:
    if abs(f.numerator) > LIMIT or f.denominator > LIMIT:
        raise Unsupported("rational overflow")
    return f


# IR nodes: ('const', Fraction) ('nconst', name) ('var', k) ('neg', x)
#           ('bin', '+-*/', l, r) ('pow', base, Fraction) ('call', opcode, x)

def _to_ir(node: ast.AST, sess: Session):
    if isinstance(node, ast.Expression):
        return _to_ir(node.body, sess)
    if isinstance(node, ast.Constant):
        if isinstance(node.value, bool) or not isinstance(node.value, int):
            raise Unsupported("non-integer literal")
        return ("const", _check(Fraction(node.value)))
    if isinstance(node, ast.Name):
        if node.id == sess.var_names[0]:
            return ("var", 0)
        if node.id == sess.var_names[1]:
            return ("var", 1)
        if node.id in sess.named_consts:
            return ("nconst", node.id)
        raise Unsupported(f"name {node.id}")
    if isinstance(node, ast.UnaryOp):
        x = _to_ir(node.operand, sess)
        if isinstance(node.op, ast.UAdd):
            return x
        if isinstance(node.op, ast.USub):
            if x[0] == "const":
                return ("const", -x[1])
            return ("neg", x)
        raise Unsupported("unary op")
    if isinstance(node, ast.BinOp):
        l = _to_ir(node.left, sess)
        r = _to_ir(node.right, sess)
        if isinstance(node.op, ast.Pow):
            if r[0] != "const":
                raise Unsupported("non-constant exponent")
            k = r[1]
            if l[0] == "const" and k.denominator == 1:
                if abs(k.numerator) > 64:
                    raise Unsupported("exponent too large")
                if l[1] == 0 and k < 0:
                    raise Unsupported("division by zero")
                return ("const", _check(l[1] ** int(k)))
            return ("pow", l, k)
        sym = {ast.Add: "+", ast.Sub: "-", ast.Mult: "*", ast.Div: "/"}.get(type(node.op))
        if sym is None:
            raise Unsupported("binary op")
        if l[0] == "const" and r[0] == "const":
            a, b = l[1], r[1]
            if sym == "+":
                return ("const", _check(a + b))
            if sym == "-":
                return ("const", _check(a - b))
            if sym == "*":
                return ("const", _check(a * b))
            if b == 0:
                raise Unsupported("division by zero")
            return ("const", _check(a / b))
        return ("bin", sym, l, r)
    if isinstance(node, ast.Call):
        if not isinstance(node.func, ast.Name) or len(node.args) != 1 or node.keywords:
            raise Unsupported("call form")
        opc = bc.FUNC_OPCODES.get(node.func.id)
        if opc is None:
            raise Unsupported(f"function {node.func.id}")
        return ("call", opc, _to_ir(node.args[0], sess))
    raise Unsupported(type(node).__name__)


def _split_terms(ir) -> List[Tuple[int, tuple]]:
    """Left spine of the +/- chain -> [(sign, body)], leading minus extracted."""
    chain = []
    while ir[0] == "bin" and ir[1] in "+-":
        chain.append((+1 if ir[1] == "+" else -1, ir[3]))
        ir = ir[2]
    chain.append((+1, ir))
    chain.reverse()
    return [_extract_sign(s, t) for s, t in chain]


def _extract_sign(sign: int, t) -> Tuple[int, tuple]:
    # descend the leftmost leaf of the * / chain
    if t[0] == "neg":
        return _extract_sign(-sign, t[1])
    if t[0] == "bin" and t[1] in "*/":
        s2, l2 = _extract_sign(sign, t[2])
        if l2 is not t[2]:
            return s2, ("bin", t[1], l2, t[3])
    return sign, t


def _emit(ir, sess: Session, out: bytearray) -> None:
    terms = _split_terms(ir)
    for k, (sign, body) in enumerate(terms):
        _emit_term(body, sess, out)
        if k == 0:
            if sign < 0:
                out.append(bc.OP_NEG)
        else:
            out.append(bc.OP_ADD if sign > 0 else bc.OP_SUB)


def _emit_term(t, sess: Session, out: bytearray) -> None:
    kind = t[0]
    if kind == "const":
        out.append(bc.OP_CONST0 + sess.const_slot(frac_key(t[1]), t[1].numerator / t[1].denominator))
    elif kind == "nconst":
        out.append(bc.OP_CONST0 + sess.const_slot(t[1], sess.named_consts[t[1]]))
    elif kind == "var":
        out.append(bc.OP_VAR0 + t[1])
    elif kind == "neg":
        _emit(t[1], sess, out)
        out.append(bc.OP_NEG)
    elif kind == "bin":
        if t[1] in "+-":
            _emit(t, sess, out)
        else:
            _emit(t[2], sess, out)
            _emit(t[3], sess, out)
            out.append(bc.OP_MUL if t[1] == "*" else bc.OP_DIV)
    elif kind == "pow":
        _emit(t[1], sess, out)
        k = t[2]
        out.append(bc.OP_POW0 + sess.pow_slot(frac_key(k), k.numerator / k.denominator))
    elif kind == "call":
        _emit(t[2], sess, out)
        out.append(t[1])
    else:  # pragma: no cover
        raise AssertionError(kind)


@dataclass
class Compiled:
    """Term-structured program of one expression string."""
    terms: List[Tuple[int, bytes]]          # (sign, postfix body) in printed order
    flags: int = bc.FLAG_OK

    def whole(self) -> bytes:
        return whole_code(self.terms)


def whole_code(terms: Sequence[Tuple[int, bytes]]) -> bytes:
    out = bytearray()
    for k, (sign, body) in enumerate(terms):
        out += body
        if k == 0:
            if sign < 0:
                out.append(bc.OP_NEG)
        else:
            out.append(bc.OP_ADD if sign > 0 else bc.OP_SUB)
    return bytes(out)


def compile_expr(s: str, sess: Session) -> Compiled:
    nc0, np0 = len(sess.const_keys), len(sess.pow_keys)
    c = _compile_expr(s, sess)
    if c.flags:  # failed compiles leave the append-only tables untouched
        del sess.const_keys[nc0:], sess.const_vals[nc0:], sess.pow_keys[np0:], sess.pow_vals[np0:]
    return c


def _compile_expr(s: str, sess: Session) -> Compiled:
    try:
        ir = _to_ir(ast.parse(s.strip(), mode="eval"), sess)
        terms = []
        for sign, body in _split_terms(ir):
            out = bytearray()
            _emit_term(body, sess, out)
            terms.append((sign, bytes(out)))
        c = Compiled(terms)
        if len(c.whole()) > 255:
            return Compiled([], bc.FLAG_TOO_LONG)
        return c
    except TableFull:
        return Compiled([], bc.FLAG_TABLE_FULL)
    except (Unsupported, SyntaxError, RecursionError, OverflowError, ZeroDivisionError):
        return Compiled([], bc.FLAG_UNSUPPORTED)


# ----------------------------------------------------------------------------
# The device enumerator's splice rule, restated on term lists.
# ----------------------------------------------------------------------------

def splice(op: int, a: Compiled, b: Optional[Compiled]) -> Optional[bytes]:
    """Postfix program of the candidate ``op(a, b)`` under the reference's
    *textual* splice (LBF:170-195).  op: 0..7 unary, 8..12 binary (after swap).
    Returns None when an operand is not device-compilable."""
    if a.flags or (b is not None and b.flags):
        return None
    if op < 8:
        return a.whole() + bytes([bc.UNARY_OPCODES[op]])
    ta, tb = list(a.terms), list(b.terms)
    name = ("add", "sub", "mul", "div", "geom_sum")[op - 8]
    if name == "add":
        terms = ta + tb
    elif name == "sub":
        terms = ta + [(-tb[0][0], tb[0][1])] + tb[1:]
    elif name == "mul":
        sa, la = ta[-1]
        sb, fb = tb[0]
        body = la + fb + (bytes([bc.OP_NEG]) if sb < 0 else b"") + bytes([bc.OP_MUL])
        terms = ta[:-1] + [(sa, body)] + tb[1:]
    elif name == "div":
        sa, la = ta[-1]
        terms = ta[:-1] + [(sa, la + b.whole() + bytes([bc.OP_DIV]))]
    else:  # geom_sum: last(a) / (1 - t1(b) +- t2(b) ...)
        sa, la = ta[-1]
        den = bytearray([bc.OP_CONST0 + bc.CONST_ONE])
        den += tb[0][1]
        if tb[0][0] < 0:
            den.append(bc.OP_NEG)
        den.append(bc.OP_SUB)
        for s, body in tb[1:]:
            den += body
            den.append(bc.OP_ADD if s > 0 else bc.OP_SUB)
        terms = ta[:-1] + [(sa, la + bytes(den) + bytes([bc.OP_DIV]))]
    return whole_code(terms)


# ----------------------------------------------------------------------------
# Decompiler (structural tests): postfix -> sympy expression
# ----------------------------------------------------------------------------

def to_sympy(code: bytes, sess: Session, locals_: Optional[dict] = None, prim_exprs: Sequence = ()):
    """Rebuild a SymPy expression from a postfix program.

    ``locals_`` maps names to symbols / callables exactly like the reference's
    ``_sympify_locals`` (GM:85-93); without it the six opaque names stay
    undefined functions like in the reference's normaliser (LB:73)."""
    import sympy as sp

    locals_ = locals_ or {}

    def sym(name):
        return locals_.get(name, sp.Symbol(name))

    def fn(name, x):
        f = locals_.get(name)
        return f(x) if f is not None else sp.Function(name)(x)

    st = []
    for op in code:
        if op == bc.OP_VAR0:
            st.append(sym(sess.var_names[0]))
        elif op == bc.OP_VAR1:
            st.append(sym(sess.var_names[1]))
        elif bc.OP_PRIM0 <= op < bc.OP_PRIM0 + bc.N_PRIM:
            st.append(prim_exprs[op - bc.OP_PRIM0])
        elif bc.OP_CONST0 <= op < bc.OP_CONST0 + bc.N_CONST:
            key = sess.const_keys[op - bc.OP_CONST0]
            if key == "E":
                st.append(sp.E)
            elif key in sess.named_consts:
                st.append(sym(key))
            else:
                st.append(sp.Rational(key))
        elif bc.is_binary(op):
            b = st.pop()
            a = st.pop()
            st.append({bc.OP_ADD: a + b, bc.OP_SUB: a - b, bc.OP_MUL: a * b, bc.OP_DIV: a / b}[op])
        elif op == bc.OP_NEG:
            st.append(-st.pop())
        elif op == bc.OP_ABS:
            st.append(sp.Abs(st.pop()))
        elif op == bc.OP_SQRT:
            st.append(sp.sqrt(st.pop()))
        elif op == bc.OP_EXP:
            st.append(sp.exp(st.pop()))
        elif bc.OP_FN_NEG <= op <= bc.OP_FN_EXPNEG:
            name = ("neg", "inv", "square", "pow_3_2", "pow_neg_3_2", "exp_neg")[op - bc.OP_FN_NEG]
            st.append(fn(name, st.pop()))
        elif bc.OP_POW0 <= op < bc.OP_POW0 + bc.N_POW:
            st.append(st.pop() ** sp.Rational(sess.pow_keys[op - bc.OP_POW0]))
        else:
            raise ValueError(hex(op))
    assert len(st) == 1
    return st[0]

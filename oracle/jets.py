"""Oracle: float64 Taylor-mode jets in two variables (TEST INFRASTRUCTURE).

Generalises the reference's own 2nd-order forward-mode evaluator
``_ad_eval_point`` (problems/force_free/validator.py:70-180) to arbitrary
order N and to arrays of points:

  leaf rules        FFV:78-83      -> ``var`` / ``const``
  Add               FFV:102-110    -> ``add`` / ``sub``
  Mul (Leibniz)     FFV:113-129    -> ``mul``  (truncated Cauchy product)
  Pow, numeric exp  FFV:132-141    -> ``powk``
  sqrt / exp        FFV:151-163    -> ``sqrt`` / ``exp``

A jet is ``numpy.ndarray[NC, P]`` of *normalised* Taylor coefficients
``c[i,j] = d_0^i d_1^j f / (i! j!)`` with ``NC = (N+1)(N+2)/2``, ordered by
total degree then by j:  idx(i, j) = n(n+1)/2 + j,  n = i + j
(order 4: u; u_r,u_z; u_rr,u_rz,u_zz; ...).  With normalised coefficients a
product is a plain convolution.

The higher-order recurrences use the radial Euler operator D = x d_x + y d_y
(D c_g = |g| c_g):   e = exp(b):  |g| e_g = sum_{b!=0} |b| b_b e_{g-b};
p = b^k:  |g| b_0 p_g = sum_{b!=0} (k|b| - |g-b|) b_b p_{g-b};
r = 1/b:  r_g = -r_0 sum_{b!=0} b_b r_{g-b}.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List, Sequence, Tuple

import numpy as np


def ncoef(order: int) -> int:
    return (order + 1) * (order + 2) // 2


def idx(i: int, j: int) -> int:
    n = i + j
    return n * (n + 1) // 2 + j


@lru_cache(maxsize=None)
def multi_indices(order: int) -> Tuple[Tuple[int, int], ...]:
    out = []
    for n in range(order + 1):
        for j in range(n + 1):
            out.append((n - j, j))
    return tuple(out)


@lru_cache(maxsize=None)
def product_pairs(order: int):
    """For each g: list of (b, g-b) index pairs with b <= g componentwise."""
    mi = multi_indices(order)
    table = []
    for (gi, gj) in mi:
        lst = []
        for bi in range(gi + 1):
            for bj in range(gj + 1):
                lst.append((idx(bi, bj), idx(gi - bi, gj - bj), bi + bj))
        table.append(lst)
    return table


def var(order: int, k: int, x: np.ndarray) -> np.ndarray:
    j = np.zeros((ncoef(order), x.shape[0]))
    j[0] = x
    if order >= 1:
        j[1 + k] = 1.0
    return j


def const(order: int, c: float, npts: int) -> np.ndarray:
    j = np.zeros((ncoef(order), npts))
    j[0] = c
    return j


def add(a, b):
    return a + b


def sub(a, b):
    return a - b


def neg(a):
    return -a


def mul(a, b, order: int):
    out = np.zeros_like(a)
    for g, lst in enumerate(product_pairs(order)):
        acc = 0.0
        for (ib, ic, _) in lst:
            acc = acc + a[ib] * b[ic]
        out[g] = acc
    return out


def square(a, order: int):
    return mul(a, a, order)


def inv(b, order: int):
    r = np.zeros_like(b)
    with np.errstate(all="ignore"):
        r0 = 1.0 / b[0]
        r[0] = r0
        for g, lst in enumerate(product_pairs(order)):
            if g == 0:
                continue
            acc = 0.0
            for (ib, ic, nb) in lst:
                if nb == 0:
                    continue
                acc = acc + b[ib] * r[ic]
            r[g] = -r0 * acc
    return r


def div(a, b, order: int):
    q = np.zeros_like(a)
    with np.errstate(all="ignore"):
        r0 = 1.0 / b[0]
        for g, lst in enumerate(product_pairs(order)):
            acc = a[g]
            for (ib, ic, nb) in lst:
                if nb == 0:
                    continue
                acc = acc - b[ib] * q[ic]
            q[g] = acc * r0
    return q


def exp(b, order: int):
    e = np.zeros_like(b)
    mi = multi_indices(order)
    with np.errstate(all="ignore"):
        e[0] = np.exp(b[0])
        for g, lst in enumerate(product_pairs(order)):
            if g == 0:
                continue
            ng = mi[g][0] + mi[g][1]
            acc = 0.0
            for (ib, ic, nb) in lst:
                if nb == 0:
                    continue
                acc = acc + nb * b[ib] * e[ic]
            e[g] = acc / ng
    return e


def _pow0(b0: np.ndarray, k: float) -> np.ndarray:
    """b0 ** k in real arithmetic: NaN where SymPy would go complex."""
    with np.errstate(all="ignore"):
        if float(k).is_integer():
            return np.power(b0, int(k)) if k >= 0 else 1.0 / np.power(b0, int(-k))
        return np.where(b0 >= 0, np.power(np.abs(b0), k), np.nan)


def powk(b, k: float, order: int):
    """b ** k, constant real exponent k."""
    p = np.zeros_like(b)
    mi = multi_indices(order)
    with np.errstate(all="ignore"):
        p[0] = _pow0(b[0], k)
        inv_b0 = 1.0 / b[0]
        for g, lst in enumerate(product_pairs(order)):
            if g == 0:
                continue
            ng = mi[g][0] + mi[g][1]
            acc = 0.0
            for (ib, ic, nb) in lst:
                if nb == 0:
                    continue
                acc = acc + (k * nb - (ng - nb)) * b[ib] * p[ic]
            p[g] = acc * inv_b0 / ng
        if float(k).is_integer() and k >= 0:
            # polynomial case: exact even when b0 == 0 (0 * inf above) -> use products
            q = const(order, 1.0, b.shape[1])
            for _ in range(int(k)):
                q = mul(q, b, order)
            bad = ~np.isfinite(p).all(axis=0) & np.isfinite(b).all(axis=0)
            p[:, bad] = q[:, bad]
    return p


def sqrt(b, order: int):
    return powk(b, 0.5, order)


def absj(b):
    with np.errstate(all="ignore"):
        s = np.sign(b[0])
        s = np.where(b[0] == 0, np.nan, s)  # |x| is not differentiable at 0
    out = b * s
    out[0] = np.abs(b[0])
    return out


def derivatives(j: np.ndarray, order: int) -> np.ndarray:
    """Normalised Taylor coefficients -> partial derivatives (multiply by i! j!)."""
    out = np.empty_like(j)
    with np.errstate(all="ignore"):
        for g, (i, jj) in enumerate(multi_indices(order)):
            out[g] = j[g] * (math.factorial(i) * math.factorial(jj))
    return out


def diff(j: np.ndarray, k: int, order: int) -> np.ndarray:
    """d/dx_k of a jet; the result is valid to order-1 (top degree set to 0)."""
    out = np.zeros_like(j)
    for g, (i, jj) in enumerate(multi_indices(order)):
        if i + jj >= order:
            continue
        if k == 0:
            out[g] = (i + 1) * j[idx(i + 1, jj)]
        else:
            out[g] = (jj + 1) * j[idx(i, jj + 1)]
    return out


# ----------------------------------------------------------------------------
# postfix interpreter
# ----------------------------------------------------------------------------

def evaluate(code: bytes, pts: np.ndarray, order: int, const_vals: Sequence[float],
             pow_vals: Sequence[float], prim_jets: Sequence[np.ndarray] = ()) -> np.ndarray:
    """Jet of a postfix program at ``pts[P,2]`` -> ``[NC, P]`` (NaN/inf where
    the real evaluation leaves the domain)."""
    from . import bytecode as bc

    P = pts.shape[0]
    st: List[np.ndarray] = []
    with np.errstate(all="ignore"):
        for op in code:
            if op == bc.OP_VAR0 or op == bc.OP_VAR1:
                st.append(var(order, op - bc.OP_VAR0, pts[:, op - bc.OP_VAR0]))
            elif bc.OP_PRIM0 <= op < bc.OP_PRIM0 + bc.N_PRIM:
                st.append(np.array(prim_jets[op - bc.OP_PRIM0], copy=True))
            elif op >= bc.OP_CONST0:
                st.append(const(order, const_vals[op - bc.OP_CONST0], P))
            elif bc.is_binary(op):
                b = st.pop()
                a = st.pop()
                if op == bc.OP_ADD:
                    st.append(a + b)
                elif op == bc.OP_SUB:
                    st.append(a - b)
                elif op == bc.OP_MUL:
                    st.append(mul(a, b, order))
                else:
                    st.append(div(a, b, order))
            elif op in (bc.OP_NEG, bc.OP_FN_NEG):
                st.append(-st.pop())
            elif op == bc.OP_ABS:
                st.append(absj(st.pop()))
            elif op == bc.OP_SQRT:
                st.append(sqrt(st.pop(), order))
            elif op == bc.OP_EXP:
                st.append(exp(st.pop(), order))
            elif op == bc.OP_FN_INV:
                st.append(inv(st.pop(), order))
            elif op == bc.OP_FN_SQUARE:
                st.append(square(st.pop(), order))
            elif op == bc.OP_FN_POW32:
                st.append(powk(st.pop(), 1.5, order))
            elif op == bc.OP_FN_POWN32:
                st.append(powk(st.pop(), -1.5, order))
            elif op == bc.OP_FN_EXPNEG:
                st.append(exp(-st.pop(), order))
            elif bc.OP_POW0 <= op < bc.OP_POW0 + bc.N_POW:
                k = pow_vals[op - bc.OP_POW0]
                a = st.pop()
                st.append(square(a, order) if k == 2.0 else powk(a, k, order))
            else:
                raise ValueError(hex(op))
    assert len(st) == 1
    return st[0]

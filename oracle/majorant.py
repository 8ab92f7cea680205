"""Oracle: round-off majorants of the jet interpreter (TEST INFRASTRUCTURE).

Why.  The reference accepts a candidate only when its residual is IDENTICALLY zero
(problems/force_free/validator.py:405-427, problems/kerr_magnetosphere/validator.py:283-294);
the device sees float64 jets.  The residual scale S = sum |monomial| bounds the round-off of
evaluating R from the partials of u, but not the round-off already inside those partials
(SURVEY 7 "hard parts": `(rho + z) - rho`; found in round 2: `z*inv(z)/rho`, which the reference
validates as 1/rho, has u_z = O(1e-16) instead of 0, so R and S are both noise and |R|/S = O(1)).
This module restates the bound the device carries next to every jet so that a float64 residual
can be called non-zero *by construction*, not by an empirical margin.

Majorant calculus.  For a jet c_g (normalised Taylor coefficients, g = (i, j)) let [C](t) be a
power series in ONE variable with non-negative coefficients such that sum_{|g| = n} |c_g| <= [C]_n.
Products, quotients and compositions of jets are majorised by the same operations on the series
(Cauchy's method of majorants).  Three numbers per jet, evaluated at a fixed radius t0 > 0:
    V >= |c_0|                          the value, summed WITHOUT cancellation
    D >= sum_{n >= 1} [C]_n t0^n        the non-constant part      (|c_g| <= D / t0^|g|)
    W    majorises the accumulated round-off in units of eps = 2^-52:
         |computed c_g - exact c_g| <= eps * W / t0^|g|            (first order in eps)
Local rounding errors are relative to the actual magnitudes (<= V + D =: M); cancellation is covered
because W is an absolute bound.  Only quotients and compositions look at the ACTUAL computed value of
their operand (d0 = |d_0|, a = |x_0|): the distance to the pole is a fact about the value, not about
its summands (1 - rho is small next to rho = 1 although V = 1 + rho).
Rules (`inf` = give up, the point does not vote):
    coordinate        V = |x|, D = t0                    W = 0
    constant c        V = |c|, D = 0                     W = |c|
    -x, |x|           unchanged
    a +- b            V = Va + Vb, D = Da + Db           W = Wa + Wb + Ma + Mb
    a * b             V = Va Vb, D = Va Db + Da Mb       W = Wa Mb + Ma Wb + 16 Ma Mb
    q = a / d         den = d0 - Dd  (<= 0: inf)
                      V = Va / d0, D = (V Dd + Da) / den W = (Wa + Mq Wd + 16 (Ma + Dd Mq)) / den
    F(x) in general   G >= sum_j |F_j| Dx^j,  G' >= dG/dDx:
                      V = G,  D = G' Dx   (G(D) - G(0) <= G'(D) D: G is convex and increasing)
                      W = G' Wx + c_F G
      x ** n, n = 0, 1, 2, ... (also square):  G = Mx^n,  G' = n Mx^(n-1),  c_F = 16 n
      1/x, inv(x), x ** -1:  the quotient rule with numerator 1 (V = 1, D = 0, W = 1)
      x ** k otherwise (sqrt: k = 1/2; Dx >= a: inf):
                      G = a^k (1 - Dx/a)^-|k|,  G' = G |k| / (a - Dx),  c_F = 64
                      (|binom(k, j)| <= binom(|k| + j - 1, j))
      exp(+-x):       G = exp(+-x_0 + Dx),  G' = G,  c_F = 64
The partial derivative d_g = i! j! c_g then carries  |delta d_g| <= eps W n! / t0^n  (n = |g|).
The device carries V, D, W in float32 next to the float64 jet (validate.cuh); its approximate
reciprocal / log2 / exp2 are covered by a 1e-4 safety factor per rule, so device and oracle agree to
about 1e-3 relative, which is all a bound needs.

Decision scale.  R is a polynomial in the partials with majorant S (absolute coefficients).  For a
true solution R(d_exact) = 0, so |R(d)| <= S(|d| + |delta|) - S(|d|) <= tau S(|d| + |delta| / tau)
for any tau <= 1 (every term of the difference has at least one delta factor).  The device votes
"non-zero" only when
    |R| > tau * S~,      S~ = S_iso(m_n + theta_n),   m_n = sum_{|g| = n} |d_g|,
                         theta_n = 2 eps W n! / (t0^n tau),
where S_iso >= S is the residual's majorant with every partial of order n replaced by m_n (one
polynomial in m_1..m_4 and 1/rho instead of 48 monomials).  `iso_tables()` derives S_iso from the
same SymPy expansion of FFV:305-347 that oracle.residuals uses.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import List, Sequence, Tuple

import numpy as np

from . import bytecode as bc
from . import jets as J

EPS = 2.0 ** -52
T0_DEFAULT = 0.0625


def _div_rule(Va, Da, Wa, d0, Dd, Wd):
    """q = a / d; d0 = |actual value of d|.  Returns (V', D', W')."""
    with np.errstate(all="ignore"):
        den = d0 - Dd
        ok = den > 0
        V = Va / d0
        D = np.where(ok, (V * Dd + Da) / den, np.inf)
        Mq = V + D
        W = np.where(ok, (Wa + Mq * Wd + 16.0 * (Va + Da + Dd * Mq)) / den, np.inf)
        return V, D, W


def _pow_rule(a, V, D, W, k: float):
    """x ** k for a real constant k; a = |actual x_0|.  Returns (V', D', W').  1/x (k = -1, inv(x)) is the quotient 1 / x."""
    if k == -1.0:
        one = np.ones_like(D)
        return _div_rule(one, np.zeros_like(D), one, a, D, W)
    with np.errstate(all="ignore"):
        if float(k).is_integer() and k >= 0:
            k = int(k)
            M = V + D
            if k == 0:
                return np.ones_like(D), np.zeros_like(D), np.ones_like(D)
            G1 = k * M ** (k - 1)
            G = M ** k
            return G, G1 * D, G1 * W + 16.0 * k * G
        ok = D < a
        ak = abs(k)
        G = np.where(ok, a ** k * (1.0 - D / a) ** (-ak), np.inf)
        G1 = np.where(ok, G * ak / (a - D), np.inf)
        return G, np.where(ok, G1 * D, np.inf), np.where(ok, G1 * W + 64.0 * G, np.inf)


def _exp_rule(x0, D, W, sign: float):
    with np.errstate(all="ignore"):
        G = np.exp(sign * x0 + D)
        return G, G * D, G * W + 64.0 * G


def evaluate(code: bytes, pts: np.ndarray, order: int, const_vals: Sequence[float], pow_vals: Sequence[float],
             prim_jets: Sequence[np.ndarray] = (), prim_maj: Sequence[Tuple[np.ndarray, np.ndarray]] = (),
             t0: float = T0_DEFAULT):
    """Jets of a postfix program at pts[P, 2] together with their majorants: (u[NC, P], V[P], D[P], W[P]).
    prim_maj[p] = (D, W) of PRIM(p) (its V is the |value| of its jet, as on the device)."""
    P = pts.shape[0]
    st: List[Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]] = []
    with np.errstate(all="ignore"):
        for op in code:
            if op == bc.OP_VAR0 or op == bc.OP_VAR1:
                x = pts[:, op - bc.OP_VAR0]
                st.append((J.var(order, op - bc.OP_VAR0, x), np.abs(x), np.full(P, t0), np.zeros(P)))
            elif bc.OP_PRIM0 <= op < bc.OP_PRIM0 + bc.N_PRIM:
                p = op - bc.OP_PRIM0
                j = np.array(prim_jets[p], copy=True)
                st.append((j, np.abs(j[0]), prim_maj[p][0].copy(), prim_maj[p][1].copy()))
            elif op >= bc.OP_CONST0:
                c = const_vals[op - bc.OP_CONST0]
                st.append((J.const(order, c, P), np.full(P, abs(c)), np.zeros(P), np.full(P, abs(c))))
            elif bc.is_binary(op):
                b, Vb, Db, Wb = st.pop()
                a, Va, Da, Wa = st.pop()
                Ma, Mb = Va + Da, Vb + Db
                if op == bc.OP_ADD or op == bc.OP_SUB:
                    st.append((a + b if op == bc.OP_ADD else a - b, Va + Vb, Da + Db, Wa + Wb + Ma + Mb))
                elif op == bc.OP_MUL:
                    st.append((J.mul(a, b, order), Va * Vb, Va * Db + Da * Mb, Wa * Mb + Ma * Wb + 16.0 * Ma * Mb))
                else:
                    V, D, W = _div_rule(Va, Da, Wa, np.abs(b[0]), Db, Wb)
                    st.append((J.div(a, b, order), V, D, W))
            elif op in (bc.OP_NEG, bc.OP_FN_NEG):
                a, V, D, W = st.pop()
                st.append((-a, V, D, W))
            elif op == bc.OP_ABS:
                a, V, D, W = st.pop()
                st.append((J.absj(a), V, D, W))
            elif op in (bc.OP_EXP, bc.OP_FN_EXPNEG):
                a, V, D, W = st.pop()
                sg = 1.0 if op == bc.OP_EXP else -1.0
                Vo, Do, Wo = _exp_rule(a[0], D, W, sg)
                st.append((J.exp(sg * a, order), Vo, Do, Wo))
            else:
                if op == bc.OP_SQRT:
                    k = 0.5
                elif op == bc.OP_FN_INV:
                    k = -1.0
                elif op == bc.OP_FN_SQUARE:
                    k = 2.0
                elif op == bc.OP_FN_POW32:
                    k = 1.5
                elif op == bc.OP_FN_POWN32:
                    k = -1.5
                elif bc.OP_POW0 <= op < bc.OP_POW0 + bc.N_POW:
                    k = pow_vals[op - bc.OP_POW0]
                else:
                    raise ValueError(hex(op))
                a, V, D, W = st.pop()
                Vo, Do, Wo = _pow_rule(np.abs(a[0]), V, D, W, k)
                if k == 2.0:
                    j = J.square(a, order)
                elif k == -1.0:
                    j = J.inv(a, order)
                else:
                    j = J.powk(a, k, order)
                st.append((j, Vo, Do, Wo))
    assert len(st) == 1
    return st[0]


# ----------------------------------------------------------------------------
# isotropic majorant of the force-free residual
# ----------------------------------------------------------------------------

@lru_cache(maxsize=None)
def iso_tables():
    """The four entries of FFV:341-347 as majorant polynomials in (m1, m2, m3, m4, w):
    tuple of dicts {(e1, e2, e3, e4, ew): coefficient}."""
    from . import residuals as Rz
    mi = J.multi_indices(Rz.FF_ORDER)
    out = []
    for tab in Rz.force_free_monomials():
        poly = {}
        for coef, expo in tab:
            e = [0, 0, 0, 0, 0]
            for v, k in enumerate(expo):
                if k == 0:
                    continue
                if v == len(mi):
                    e[4] += k                     # w = 1/rho
                else:
                    n = mi[v][0] + mi[v][1]
                    assert n >= 1                 # the residual never uses u itself
                    e[n - 1] += k
            key = tuple(e)
            poly[key] = poly.get(key, 0.0) + abs(coef)
        out.append(poly)
    return tuple(out)


def order_sums(d: np.ndarray, order: int) -> np.ndarray:
    """m_n = sum over |g| = n of |d_g|, n = 1..order  ->  [order, P]"""
    out = np.zeros((order, d.shape[1]))
    for g, (i, j) in enumerate(J.multi_indices(order)):
        if i + j >= 1:
            out[i + j - 1] += np.abs(d[g])
    return out


def theta(W: np.ndarray, order: int, tau: float, t0: float = T0_DEFAULT) -> np.ndarray:
    """theta_n = 2 eps W n! / (t0^n tau), n = 1..order  ->  [order, P]"""
    return np.stack([2.0 * EPS * W * math.factorial(n) / (t0 ** n * tau) for n in range(1, order + 1)])


def force_free_scale(u: np.ndarray, rho: np.ndarray, W: np.ndarray, tau: float, t0: float = T0_DEFAULT) -> np.ndarray:
    """S~ = a0 a3 + a1 a2 with a_k = entry k's isotropic majorant at m_n + theta_n."""
    d = J.derivatives(u, 4)
    with np.errstate(all="ignore"):
        m = order_sums(d, 4) + theta(W, 4, tau, t0)
        w = np.abs(1.0 / rho)
        a = []
        for poly in iso_tables():
            acc = np.zeros(u.shape[1])
            for (e1, e2, e3, e4, ew), c in poly.items():
                acc = acc + c * m[0] ** e1 * m[1] ** e2 * m[2] ** e3 * m[3] ** e4 * w ** ew
            a.append(acc)
        return a[0] * a[3] + a[1] * a[2]


def kerr_scale(u: np.ndarray, pts: np.ndarray, W: np.ndarray, tau: float, t0: float = T0_DEFAULT, M_: float = 1.0, a_: float = 0.1) -> np.ndarray:
    """S~ = sum_k |c_k| (|d_k| + theta_order(k)) for R = c1_r u_r + c1 u_rr + c2_x u_x + c2 u_xx (KV:77-91)."""
    from . import residuals as Rz
    d = J.derivatives(u, 2)
    c = np.abs(Rz.kerr_coeffs(pts, M_, a_))
    th = theta(W, 2, tau, t0)
    with np.errstate(all="ignore"):
        return (c[:, 1] * (np.abs(d[1]) + th[0]) + c[:, 0] * (np.abs(d[3]) + th[1]) +
                c[:, 3] * (np.abs(d[2]) + th[0]) + c[:, 2] * (np.abs(d[5]) + th[1]))

"""Oracle: the two PDE residual operators on float64 jets (TEST INFRASTRUCTURE).

force-free (problems/force_free/validator.py:305-347, Omega = 0):
    A = u_rr + u_zz - u_r / rho            FFV:319-320
    B = u_r**2 + u_z**2                    FFV:321
    L_T f = u_z d_rho f - u_r d_z f        FFV:335-339
    R = det [[L_T A, L_T B], [L_T^2 A, L_T^2 B]]     FFV:341-347

Kerr (problems/kerr_magnetosphere/validator.py:69-91):
    Delta = r^2 - 2 M r + a^2,  G = 1 - 2 M r / (r^2 + a^2 x^2)
    R = d_r[ G/(1-x^2) u_r ] + d_x[ G/Delta u_x ]

Two independent evaluations are provided and cross-checked in the tests:
  * ``*_structured``: the formula applied line by line to truncated jets
    (differentiate = shift coefficients, multiply = Cauchy product);
  * ``*_monomials``: the four Lie-derivative entries expanded by SymPy into
    polynomials in the partial derivatives of u and w = 1/rho; this form also
    defines the *scale*  S = |LT_A|_1 |L2T_B|_1 + |LT_B|_1 |L2T_A|_1  where
    |p|_1 = sum of |monomial| -- the quantity round-off in R is proportional
    to (SURVEY 7 "hard parts"); the device reports R and S per point.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Dict, List, Tuple

import numpy as np

from . import jets as J

FF_ORDER = 4
KERR_ORDER = 2


# ----------------------------------------------------------------------------
# force-free, structured (jet arithmetic)
# ----------------------------------------------------------------------------

def force_free_parts_structured(u: np.ndarray, rho: np.ndarray):
    """(LT_A, LT_B, L2T_A, L2T_B) at each point from the order-4 jet of u."""
    N = FF_ORDER
    P = u.shape[1]
    with np.errstate(all="ignore"):
        u_r = J.diff(u, 0, N)
        u_z = J.diff(u, 1, N)
        u_rr = J.diff(u_r, 0, N)
        u_zz = J.diff(u_z, 1, N)
        rho_j = J.var(N, 0, rho)
        A = u_rr + u_zz - J.div(u_r, rho_j, N)
        B = J.mul(u_r, u_r, N) + J.mul(u_z, u_z, N)

        def LT(f):
            return J.mul(u_z, J.diff(f, 0, N), N) - J.mul(u_r, J.diff(f, 1, N), N)

        LT_A = LT(A)
        LT_B = LT(B)
        L2T_A = LT(LT_A)
        L2T_B = LT(LT_B)
    return LT_A[0], LT_B[0], L2T_A[0], L2T_B[0]


# ----------------------------------------------------------------------------
# force-free, monomial tables (SymPy expands the same formula once)
# ----------------------------------------------------------------------------

@lru_cache(maxsize=None)
def force_free_monomials() -> Tuple[Tuple[Tuple[float, Tuple[int, ...]], ...], ...]:
    """Four parts; each a tuple of (coef, exponents) with exponents over the
    variables [d_0 .. d_14, w]: d_g = partial derivative idx g of u (NOT the
    normalised Taylor coefficient), w = 1/rho."""
    import sympy as sp

    rho, z = sp.symbols("rho z", positive=True)
    u = sp.Function("u")(rho, z)
    u_rho = u.diff(rho)
    u_z = u.diff(z)
    A = u_rho.diff(rho) + u_z.diff(z) - u_rho / rho
    B = u_rho ** 2 + u_z ** 2

    def LT(f):
        return u_z * f.diff(rho) - u_rho * f.diff(z)

    LT_A, LT_B = LT(A), LT(B)
    L2T_A, L2T_B = LT(LT_A), LT(LT_B)
    d = [sp.Symbol(f"d{g}") for g in range(J.ncoef(FF_ORDER))]
    w = sp.Symbol("w")
    gens = d + [w]

    def to_table(e):
        e = e.doit()
        rep = {}
        for der in e.atoms(sp.Derivative):
            cnt = dict(der.variable_count)
            rep[der] = d[J.idx(cnt.get(rho, 0), cnt.get(z, 0))]
        e = e.xreplace(rep).subs(rho, 1 / w)
        poly = sp.Poly(sp.expand(e), *gens)
        return tuple((float(c), tuple(int(k) for k in mon)) for mon, c in zip(poly.monoms(), poly.coeffs()))

    return tuple(to_table(p) for p in (LT_A, LT_B, L2T_A, L2T_B))


def _eval_parts(tables, vars_: np.ndarray):
    """vars_ [NV, P] -> signed sums [4, P] and abs sums [4, P]."""
    P = vars_.shape[1]
    val = np.zeros((len(tables), P))
    ab = np.zeros((len(tables), P))
    with np.errstate(all="ignore"):
        for k, tab in enumerate(tables):
            for coef, expo in tab:
                t = np.full(P, coef)
                for v, e in enumerate(expo):
                    for _ in range(e):
                        t = t * vars_[v]
                val[k] += t
                ab[k] += np.abs(t)
    return val, ab


def force_free_residual(u: np.ndarray, rho: np.ndarray):
    """R, S, parts[4,P] from the order-4 jet ``u[15,P]``."""
    d = J.derivatives(u, FF_ORDER)
    with np.errstate(all="ignore"):
        vars_ = np.vstack([d, (1.0 / rho)[None, :]])
        val, ab = _eval_parts(force_free_monomials(), vars_)
        R = val[0] * val[3] - val[1] * val[2]
        S = ab[0] * ab[3] + ab[1] * ab[2]
    return R, S, val


# ----------------------------------------------------------------------------
# Kerr
# ----------------------------------------------------------------------------

def kerr_coeffs(pts: np.ndarray, M: float, a: float) -> np.ndarray:
    """Per-point table [P,4] = (c1, c1_r, c2, c2_x) with
    c1 = G/(1-x^2), c2 = G/Delta  (KV:69-91)."""
    r, x = pts[:, 0], pts[:, 1]
    with np.errstate(all="ignore"):
        Sg = r * r + a * a * x * x
        G = 1.0 - 2.0 * M * r / Sg
        G_r = -2.0 * M / Sg + 4.0 * M * r * r / (Sg * Sg)
        G_x = 4.0 * M * r * a * a * x / (Sg * Sg)
        Delta = r * r - 2.0 * M * r + a * a
        om = 1.0 - x * x
        return np.stack([G / om, G_r / om, G / Delta, G_x / Delta], axis=1)


def kerr_residual(u: np.ndarray, pts: np.ndarray, M: float = 1.0, a: float = 0.1):
    """R, S from the order-2 jet ``u[6,P]`` (idx: u; u_r,u_x; u_rr,u_rx,u_xx)."""
    d = J.derivatives(u, KERR_ORDER)
    c = kerr_coeffs(pts, M, a)
    with np.errstate(all="ignore"):
        t = np.stack([c[:, 1] * d[1], c[:, 0] * d[3], c[:, 3] * d[2], c[:, 2] * d[5]])
        R = t.sum(axis=0)
        S = np.abs(t).sum(axis=0)
    return R, S, t


# ----------------------------------------------------------------------------
# grids
# ----------------------------------------------------------------------------

MASK64 = (1 << 64) - 1


def splitmix64_stream(seed: int):
    s = seed & MASK64
    while True:
        s = (s + 0x9E3779B97F4A7C15) & MASK64
        x = s
        x ^= x >> 30
        x = (x * 0xBF58476D1CE4E5B9) & MASK64
        x ^= x >> 27
        x = (x * 0x94D049BB133111EB) & MASK64
        x ^= x >> 31
        yield x


def _u01(g) -> float:
    return (next(g) >> 11) * (1.0 / (1 << 53))


def collocation_grid(problem: str, P: int, seed: int = 0x5EED9017) -> np.ndarray:
    """The reference's own test points first, then uniform points."""
    g = splitmix64_stream(seed)
    pts = np.empty((P, 2))
    if problem == "force_free":
        ref = [(4 / 5, 6 / 7), (3 / 4, 5 / 6), (7 / 8, 1 / 2)]  # FFV:296-297, LBF:278-282
        lo0, w0, lo1, w1 = 0.25, 1.75, 0.25, 1.75
    else:
        ref = [(5 / 2, 3 / 5), (7 / 3, 1 / 3), (5.0, -2 / 5)]   # KV:168-172
        lo0, w0, lo1, w1 = 2.2, 3.8, -0.9, 1.8
    for k in range(P):
        if k < len(ref):
            pts[k] = ref[k]
        else:
            u1 = _u01(g)
            u2 = _u01(g)
            pts[k] = (lo0 + w0 * u1, lo1 + w1 * u2)
    return pts

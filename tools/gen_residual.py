#!/usr/bin/env python3
"""Generate the force-free residual device code (committed output:
pde_engine_b200/csrc/residual_ff_gen.cuh).

The formula is the reference's (problems/force_free/validator.py:305-347, Omega = 0):
    A = u_rr + u_zz - u_r/rho,  B = u_r^2 + u_z^2,  L_T f = u_z f_rho - u_r f_z,
    M = [[L_T A, L_T B], [L_T^2 A, L_T^2 B]],   R = det M.
SymPy applies it to a generic u(rho, z); every Derivative becomes a symbol
d<idx> (partial derivative, idx(i,j) = n(n+1)/2 + j) and 1/rho becomes w.  Each
matrix entry is expanded into monomials; the device evaluates every monomial
once and accumulates both the signed sum p_k and the absolute sum a_k
(S = a0 a3 + a1 a2 is the round-off scale of R = p0 p3 - p1 p2).

The SAME schedule (which products are shared, which fused multiply-adds are formed, in which order they are
accumulated) is also emitted as a run-time residual program (pde_engine_b200/residual_programs.py, the input of
pde_compile_residual_program), together with the Kerr residual in the schedule of Residual<KERR> (validate.cuh):
tests/test_gpu_program.py shows that the interpreter and the CUDA specialisations agree bit for bit.

Run:  python tools/gen_residual.py   (needs sympy; the outputs are committed so
the build itself does not).
"""
from __future__ import annotations

import os
import sys

import sympy as sp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from pde_engine_b200.residual_compiler import ProgramBuilder   # noqa: E402


def jidx(i, j):
    return (i + j) * (i + j + 1) // 2 + j


def main():
    rho, z = sp.symbols("rho z", positive=True)
    u = sp.Function("u")(rho, z)
    u_rho, u_z = u.diff(rho), u.diff(z)
    A = u_rho.diff(rho) + u_z.diff(z) - u_rho / rho
    B = u_rho ** 2 + u_z ** 2

    def LT(f):
        return u_z * f.diff(rho) - u_rho * f.diff(z)

    LT_A, LT_B = LT(A), LT(B)
    entries = [LT_A, LT_B, LT(LT_A), LT(LT_B)]
    d = [sp.Symbol(f"d{g}") for g in range(15)]
    w = sp.Symbol("w")

    gens = d + [w]
    monos = []   # (part, coef, exponent tuple over gens)
    for k, e in enumerate(entries):
        e = e.doit()
        rep = {}
        for der in e.atoms(sp.Derivative):
            cnt = dict(der.variable_count)
            rep[der] = d[jidx(cnt.get(rho, 0), cnt.get(z, 0))]
        poly = sp.Poly(sp.expand(e.xreplace(rep).subs(rho, 1 / w)), *gens)
        for mon, coef in zip(poly.monoms(), poly.coeffs()):
            monos.append((k, int(coef), mon))
    # Every monomial is evaluated as  left * right  with two fused multiply-adds:
    #   p += (+-left) * right          (signed sum)
    #   a += |left| * |right|          (absolute sum; |.| is a free operand modifier)
    # `left` is a product chain over the variables in a fixed global order, shared
    # between monomials through memoised prefixes.
    freq = [0] * len(gens)
    for _, _, mon in monos:
        for v, e in enumerate(mon):
            freq[v] += e
    order = sorted(range(len(gens)), key=lambda v: -freq[v])
    prefix = {}          # tuple of var indices -> C name
    decl = []
    # the run-time program of the same schedule
    pb = ProgramBuilder(4, 1)
    inv_idx = {}
    for i_ in range(5):
        for j_ in range(5 - i_):
            inv_idx[jidx(i_, j_)] = (i_, j_)
    operand = {f"d{g}": pb.d(*inv_idx[g]) for g in range(15)}
    operand["w"] = pb.col(0)

    def name_of(v):
        return str(gens[v])

    def product(factors):
        if len(factors) == 1:
            return name_of(factors[0])
        key = tuple(factors)
        if key not in prefix:
            left = product(factors[:-1])
            nm = f"m{len(prefix)}"
            prefix[key] = nm
            decl.append(f"    const double {nm} = {left} * {name_of(factors[-1])};")
            operand[nm] = pb.mul(operand[left], operand[name_of(factors[-1])])
        return prefix[key]

    body = []
    for k in range(4):
        groups = {}
        for kk, coef, mon in monos:
            if kk != k:
                continue
            factors = [v for v in order for _ in range(mon[v])]
            left = product(factors[:-1])
            right = name_of(factors[-1])
            groups.setdefault(abs(coef), []).append((1 if coef > 0 else -1, left, right))
        pk, ak = [], []
        for gi, (c, items) in enumerate(sorted(groups.items())):
            ps, as_ = f"p{k}_{gi}", f"a{k}_{gi}"
            for n_, (sgn, left, right) in enumerate(items):
                sl = f"-{left}" if sgn < 0 else left
                if n_ == 0:
                    body.append(f"    double {ps} = {sl} * {right}, {as_} = fabs({left}) * fabs({right});")
                    pb.acc0(operand[left], operand[right], sgn < 0)
                else:
                    body.append(f"    {ps} = fma({sl}, {right}, {ps}); {as_} = fma(fabs({left}), fabs({right}), {as_});")
                    pb.acc(operand[left], operand[right], sgn < 0)
            operand[ps] = pb.sta()
            pk.append((c, ps))
            ak.append((c, as_))

        def combine(lst):
            out = None
            for c, nm in lst:
                if out is None:
                    out = nm if c == 1 else f"{float(c)!r} * {nm}"
                else:
                    out = f"fma({float(c)!r}, {nm}, {out})" if c != 1 else f"({out} + {nm})"
            return out
        body.append(f"    p[{k}] = {combine(pk)}; a[{k}] = {combine(ak)};")
        # the same combination in the program: c * nm | fma(c, nm, out) | (out + nm)
        for q, (c, nm) in enumerate(pk):
            if q == 0:
                if c == 1:
                    pb.lda(operand[nm])
                else:
                    pb.acc0(pb.const(float(c)), operand[nm])
            elif c != 1:
                pb.acc(pb.const(float(c)), operand[nm])
            else:
                pb.adda(operand[nm])
        operand[f"p{k}"] = pb.sta()

    lines = []
    lines.append("// GENERATED by tools/gen_residual.py -- do not edit.")
    lines.append("// force-free residual entries (problems/force_free/validator.py:305-347):")
    lines.append("// p[0..3] = LT_A, LT_B, L2T_A, L2T_B; a[k] = sum of |monomial| of entry k.")
    lines.append(f"// {len(monos)} monomials in the partial derivatives d[idx(i,j)] of u and w = 1/rho;")
    lines.append(f"// {len(prefix)} shared sub-products + 2 fused multiply-adds per monomial.")
    lines.append("#pragma once")
    lines.append("namespace pde {")
    lines.append("__device__ __forceinline__ void ff_residual_entries(const double* __restrict__ d, double w, double* __restrict__ p, double* __restrict__ a) {")
    for g in range(15):
        lines.append(f"    const double d{g} = d[{g}];")
    lines += decl
    lines += body
    lines.append("}")
    lines.append("}  // namespace pde")
    path = os.path.join(REPO, "pde_engine_b200", "csrc", "residual_ff_gen.cuh")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", path, "monomials", len(monos), "shared sub-products", len(prefix))
    # R = fma(p0, p3, -(p1 * p2))   (Residual<FORCE_FREE>::eval)
    pb.acc0(operand["p1"], operand["p2"], True)
    pb.acc(operand["p0"], operand["p3"])
    pb.out()
    ff_words, ff_file = pb.assemble()

    # Kerr (KV:77-91 expanded; Residual<KERR>::eval): R = (c1r u_r + c1 u_rr) + (c2x u_x + c2 u_xx), columns
    # c1 = G/(1-x^2), c1r = d_r c1, c2 = G/Delta, c2x = d_x c2; four products, three sums, nothing fused
    kb = ProgramBuilder(2, 4)
    t0 = kb.mul(kb.col(1), kb.d(1, 0))
    t1 = kb.mul(kb.col(0), kb.d(2, 0))
    t2 = kb.mul(kb.col(3), kb.d(0, 1))
    t3 = kb.mul(kb.col(2), kb.d(0, 2))
    kb.lda(t0); kb.adda(t1); s01 = kb.sta()
    kb.lda(t2); kb.adda(t3); s23 = kb.sta()
    kb.lda(s01); kb.adda(s23); kb.out()
    k_words, k_file = kb.assemble()

    out = os.path.join(REPO, "pde_engine_b200", "residual_programs.py")
    with open(out, "w") as f:
        f.write('"""GENERATED by tools/gen_residual.py -- do not edit.\n\n'
                'The two built-in residuals as run-time residual programs (pde_compile_residual_program), in the exact\n'
                'schedule of their CUDA specialisations (csrc/residual_ff_gen.cuh, Residual<KERR> in csrc/validate.cuh).\n"""\n')
        f.write(f"FORCE_FREE = dict(order=4, n_cols=1, consts={pb.consts!r}, n_file={ff_file},\n    words={ff_words!r})\n\n")
        f.write(f"KERR = dict(order=2, n_cols=4, consts={kb.consts!r}, n_file={k_file},\n    words={k_words!r})\n")
    print("wrote", out, "force-free:", len(ff_words), "words, file", ff_file, "| Kerr:", len(k_words), "words, file", k_file)


if __name__ == "__main__":
    main()

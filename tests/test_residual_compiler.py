"""Run-time residual programs (north_star item 3), host side: the assembler, the SymPy front end and the C-ABI
validation, checked on the CPU against the oracle's residuals (oracle/residuals.py: FFV:305-347, KV:77-91) with the
host restatement of the device interpreter (residual_compiler.interpret)."""
import numpy as np
import pytest
import sympy as sp

from oracle import jets as J
from oracle import parser as op
from oracle import residuals as Rz
from pde_engine_b200 import residual_programs as rp
from pde_engine_b200.residual_compiler import (R_MAX_FILE, ProgramBuilder, compile_residual, decode, interpret, n_coef,
                                               R_ACC, R_ACC0, R_MUL, R_OUT, R_STA, word)

FF_STRS = ["rho*z", "rho**3 + z**2", "rho**2/z", "z/(1 - rho)", "rho/(rho**2 + z**2)", "rho + exp(rho/z)", "rho**2*z",
           "rho**2*exp(-2*z)", "sqrt(rho**2 + z**2) - z", "sqrt(rho)*exp(z)/(rho + z)"]
KERR_STRS = ["r", "x", "a**2*x**2 + r**2", "r*x", "exp(-r)*(1 - x)", "sqrt(r)/(1 - x)", "1 - x", "r**2*x/(1 + x**2)"]


def _jets(problem, strs, P=48):
    sess = op.Session.for_problem(problem)
    pts = Rz.collocation_grid(problem, P)
    order = 4 if problem == "force_free" else 2
    us = [J.evaluate(op.compile_expr(s, sess).whole(), pts, order, sess.const_vals, sess.pow_vals) for s in strs]
    return pts, order, us


def test_builtin_programs_equal_the_oracle_residuals():
    """The generated programs of the two built-in residuals (same schedule as the CUDA specialisations) give the
    oracle's R and its scale S = sum of |monomial|."""
    pts, order, us = _jets("force_free", FF_STRS)
    w = (1.0 / pts[:, 0])[None, :]
    for s, u in zip(FF_STRS, us):
        d = J.derivatives(u, 4)
        with np.errstate(all="ignore"):
            R0, S0, _ = Rz.force_free_residual(u, pts[:, 0])
        R = interpret(rp.FORCE_FREE["words"], 4, 1, rp.FORCE_FREE["consts"], d, w)
        S = interpret(rp.FORCE_FREE["words"], 4, 1, rp.FORCE_FREE["consts"], d, w, magnitudes=True)
        assert np.all(np.abs(R - R0) <= 1e-13 * S0), s
        assert np.allclose(S, S0, rtol=1e-12), s
    pts, order, us = _jets("kerr_magnetosphere", KERR_STRS)
    cols = Rz.kerr_coeffs(pts, 1.0, 0.1).T
    for s, u in zip(KERR_STRS, us):
        d = J.derivatives(u, 2)
        R0, S0, _ = Rz.kerr_residual(u, pts)
        R = interpret(rp.KERR["words"], 2, 4, rp.KERR["consts"], d, cols)
        S = interpret(rp.KERR["words"], 2, 4, rp.KERR["consts"], d, cols, magnitudes=True)
        assert np.all(np.abs(R - R0) <= 1e-14 * S0), s
        assert np.allclose(S, S0, rtol=1e-13), s
    assert rp.FORCE_FREE["n_file"] <= 34          # fits two spill slots of 17 entries: no special kernel configuration


def test_front_end_compiles_the_reference_formulas():
    """compile_residual on the reference's own formulas -- det M of FFV:305-347 and `_lhs` of KV:77-91 applied to a
    generic Function('u') -- agrees with the oracle's residuals point by point."""
    rho, z = sp.symbols("rho z", positive=True)
    u = sp.Function("u")(rho, z)
    ur, uz = u.diff(rho), u.diff(z)
    A = ur.diff(rho) + uz.diff(z) - ur / rho
    B = ur ** 2 + uz ** 2
    LT = lambda f: uz * f.diff(rho) - ur * f.diff(z)      # noqa: E731
    M = sp.Matrix([[LT(A), LT(B)], [LT(LT(A)), LT(LT(B))]])
    cr = compile_residual(M.det(), u, (rho, z))
    assert cr.order == 4 and cr.n_cols <= 4 and cr.n_file <= R_MAX_FILE
    pts, _, us = _jets("force_free", FF_STRS)
    tab = cr.point_table(np.ascontiguousarray(pts.T))
    for s, uj in zip(FF_STRS, us):
        d = J.derivatives(uj, 4)
        with np.errstate(all="ignore"):
            R0, S0, _ = Rz.force_free_residual(uj, pts[:, 0])
        R = interpret(cr.words, cr.order, cr.n_cols, cr.consts, d, tab)
        S = interpret(cr.words, cr.order, cr.n_cols, cr.consts, d, tab, magnitudes=True)
        ok = np.isfinite(R0) & np.isfinite(S0)
        assert np.all(np.abs(R - R0)[ok] <= 1e-11 * np.maximum(S, S0)[ok]), s      # a different (fully expanded) evaluation order
        assert np.all(S[ok] >= 0.999 * np.abs(R[ok]))

    r, x, Ms, a = sp.symbols("r x M a", real=True)
    v = sp.Function("u")(r, x)
    Delta = r ** 2 - 2 * Ms * r + a ** 2
    G = 1 - 2 * Ms * r / (r ** 2 + a ** 2 * x ** 2)
    lhs = sp.diff(G / (1 - x ** 2) * v.diff(r), r) + sp.diff(G / Delta * v.diff(x), x)       # KV:82-91
    ck = compile_residual(lhs, v, (r, x), params={Ms: 1, a: sp.Rational(1, 10)})
    assert ck.order == 2 and ck.n_cols == 4
    pts, _, us = _jets("kerr_magnetosphere", KERR_STRS)
    tab = ck.point_table(np.ascontiguousarray(pts.T))
    for s, uj in zip(KERR_STRS, us):
        d = J.derivatives(uj, 2)
        R0, S0, _ = Rz.kerr_residual(uj, pts)
        R = interpret(ck.words, 2, ck.n_cols, ck.consts, d, tab)
        assert np.all(np.abs(R - R0) <= 1e-11 * S0), s
    # golden vectors of SURVEY 8c (reference `_lhs` at M = 1, a = 1/10): u = r and u = r*x at (5/2, 3/5)
    sess = op.Session.for_problem("kerr_magnetosphere")
    p0 = np.array([[2.5, 0.6]])
    for s, want in (("r", 0.49913682877163457), ("r*x", 0.30252620848450956)):
        uj = J.evaluate(op.compile_expr(s, sess).whole(), p0, 2, sess.const_vals, sess.pow_vals)
        got = interpret(ck.words, 2, ck.n_cols, ck.consts, J.derivatives(uj, 2), ck.point_table(p0.T.copy()))
        assert abs(got[0] - want) < 1e-12 * abs(want)


def test_toy_plugin_axisymmetric_laplace():
    """A third PDE: u_rr + u_r/rho + u_zz = 0.  Harmonic candidates give R = 0 to round-off, others do not."""
    rho, z = sp.symbols("rho z", positive=True)
    u = sp.Function("u")(rho, z)
    cr = compile_residual(u.diff(rho, 2) + u.diff(rho) / rho + u.diff(z, 2), u, (rho, z))
    assert (cr.order, cr.n_cols) == (2, 1) and cr.n_file <= 16
    sess = op.Session.for_problem("force_free")
    pts = Rz.collocation_grid("force_free", 48)
    tab = cr.point_table(np.ascontiguousarray(pts.T))
    for s, harmonic in (("z", True), ("rho**2 - 2*z**2", True), ("1/sqrt(rho**2 + z**2)", True), ("z/(rho**2 + z**2)**(3/2)", True),
                        ("rho**2", False), ("rho*z", False), ("exp(z)/rho", False)):
        uj = J.evaluate(op.compile_expr(s, sess).whole(), pts, 2, sess.const_vals, sess.pow_vals)
        d = J.derivatives(uj, 2)
        R = interpret(cr.words, 2, 1, cr.consts, d, tab)
        S = interpret(cr.words, 2, 1, cr.consts, d, tab, magnitudes=True)
        if harmonic:
            assert np.all(np.abs(R) <= 1e-12 * S + 1e-300), s
        else:
            assert np.median(np.abs(R) / S) > 1e-3, s


def test_assembler_liveness_and_limits():
    b = ProgramBuilder(2, 1)
    t = b.mul(b.d(1, 0), b.col(0))
    dead = b.mul(b.d(0, 1), b.d(0, 1))           # never used: not emitted
    b.acc0(t, b.d(1, 0))
    b.acc(b.const(3.0), b.d(2, 0), neg=True)
    s1 = b.sta()
    b.lda(s1)
    b.adda(b.d(0, 2))
    b.out()
    words, n_file = b.assemble()
    ops = [decode(w)[0] for w in words]
    assert ops == [R_MUL, R_ACC0, R_ACC, R_STA, 4, 6, R_OUT] and dead is not None
    nc = n_coef(2)
    assert n_file == nc + 1 + 1 + 1               # inputs + one temporary (the product's slot is reused by s1)
    assert decode(words[2]) == (R_ACC, nc + 1, 3, 0, True)
    with pytest.raises(ValueError):
        bb = ProgramBuilder(4, 0)
        regs = [bb.mul(bb.d(1, 0), bb.d(0, 1)) for _ in range(60)]
        bb.acc0(regs[0], regs[1])
        for r_ in regs:                           # 60 products alive at once: more than the file holds
            bb.acc(r_, r_)
        for r_ in regs:
            bb.acc(r_, r_)
        bb.out()
        bb.assemble()
    with pytest.raises(ValueError):
        rho, z = sp.symbols("rho z", positive=True)
        u = sp.Function("u")(rho, z)
        compile_residual(sp.exp(u.diff(rho)), u, (rho, z))       # not a polynomial in the derivatives


def test_abi_rejects_malformed_programs():
    from pde_engine_b200 import _lib, core
    ok = core.ResidualProgram.from_words(2, 1, [2.0], [word(R_ACC0, 1, 6), word(R_ACC, 7, 3), word(R_OUT)])
    assert (ok.order, ok.n_coef, ok.cols, ok.problem_id) == (2, 6, 1, core.PROBLEM_PROGRAM)
    for bad in ([word(R_ACC0, 1, 9), word(R_OUT)],                 # reads a temporary nobody wrote
                [word(R_ACC0, 1, 2), word(R_STA, 0, 0, 3), word(R_OUT)],     # writes an input
                [word(R_ACC, 1, 2), word(R_OUT)],                  # accumulates into nothing
                [word(R_ACC0, 1, 2)],                              # no OUT
                [word(R_ACC0, 1, 2), word(R_OUT), word(R_OUT)]):
        with pytest.raises(_lib.PdeError):
            core.ResidualProgram.from_words(2, 1, [2.0], bad)
    with pytest.raises(_lib.PdeError):
        core.ResidualProgram.from_words(3, 0, [], [word(R_OUT)])

"""Oracle parser / jets / residuals vs reference-derived golden vectors (CPU)."""
import numpy as np
import pytest
import sympy as sp

from conftest import uniques_by_depth
from oracle import bytecode as bc
from oracle import enumerate as oe
from oracle import jets as J
from oracle import parser as op
from oracle import residuals as Rz

# golden values computed from the reference formula at (4/5, 6/7): SURVEY 8c
SURVEY_R = {
    "rho*z": 384 / 35,
    "rho**3 + z**2": 6469632 / 8575,
    "rho**2/z": -2024782584832 / 13839609375,
    "z/(1 - rho)": -695068359375 / 2744,
    "rho**2 + rho*z": 7424 / 175,
    "rho/(rho**2 + z**2)": 0.6688667569155341,
    "rho + exp(rho/z)": -18898.767383500639,
    "rho*exp(rho/z)": -119417.05954571853,
}
SURVEY_JET_RHO2_OVER_Z = [0.7466666666666667, 1.8666666666666667, -0.87111111111111106, 2.3333333333333335,
                          -2.1777777777777776, 2.0325925925925925, 0, -2.7222222222222223, 5.0814814814814815,
                          -7.1140740740740744, 0, 0, 6.3518518518518521, -17.785185185185185, 33.199012345679016]
SURVEY_KERR = {   # KV _lhs at M=1, a=1/10, points (5/2,3/5), (7/3,1/3), (5,-2/5)
    "r": [0.49913682877163457, 0.41301237258376955, 0.09521981147411096],
    "x": [0.0012176444886115207, 0.0013317644264461499, -8.5265568003151096e-06],
    "a**2*x**2 + r**2": [3.125319839879233, 2.2528533063406169, 2.3816300611356378],
    "r*x": [0.30252620848450956, 0.14077824118963086, -0.038130557373645964],
    "exp(-r)*(1 - x)": [-0.0062043570088011982, -0.016426985457951997, 0.0058400714577483319],
    "sqrt(r)/(1 - x)": [8.2181211536773215, 2.0625010126232111, 0.068940535827255095],
}


def _ff_eval(s, pts):
    sess = op.Session.for_problem("force_free")
    c = op.compile_expr(s, sess)
    assert c.flags == 0, s
    u = J.evaluate(c.whole(), pts, 4, sess.const_vals, sess.pow_vals)
    return u, Rz.force_free_residual(u, pts[:, 0])


def test_survey_golden_residuals():
    pts = np.array([[4 / 5, 6 / 7]])
    for s, want in SURVEY_R.items():
        _, (R, S, _) = _ff_eval(s, pts)
        assert abs(R[0] - want) <= 1e-10 * S[0], s
    u, _ = _ff_eval("rho**2/z", pts)
    got = J.derivatives(u, 4)[:, 0]
    np.testing.assert_allclose(got, SURVEY_JET_RHO2_OVER_Z, rtol=1e-12, atol=1e-13)
    # known solutions: R = 0 with LT_A = L2T_A = 0 and the quoted (LT_B, L2T_B)
    for s, ltb, l2tb in [("rho**2*z", -1.6985861224489796, 7.6503248979591838),
                         ("rho**2*exp(-2*z)", 0.047849286395455479, 0.041362985383618267),
                         ("sqrt(rho**2 + z**2) - z", 0.31302355952730204, 0.44125751579422567)]:
        _, (R, S, parts) = _ff_eval(s, pts)
        assert abs(R[0]) <= 1e-13 * S[0]
        assert abs(parts[1, 0] - ltb) < 1e-12 and abs(parts[3, 0] - l2tb) < 1e-12
        assert abs(parts[0, 0]) < 1e-13 and abs(parts[2, 0]) < 1e-12


def test_survey_golden_kerr():
    pts = np.array([[5 / 2, 3 / 5], [7 / 3, 1 / 3], [5.0, -2 / 5]])
    sess = op.Session.for_problem("kerr_magnetosphere")
    for s, want in SURVEY_KERR.items():
        c = op.compile_expr(s, sess)
        u = J.evaluate(c.whole(), pts, 2, sess.const_vals, sess.pow_vals)
        R, S, _ = Rz.kerr_residual(u, pts)
        np.testing.assert_allclose(R, want, rtol=1e-11, atol=1e-15)


@pytest.mark.parametrize("problem", ["force_free", "kerr_magnetosphere"])
def test_residuals_match_reference_vectors(problem, resid_ff, resid_kerr):
    """Per-point residual of every golden record (the reference's own det_M / _lhs
    evaluated with evalf(50)) within 1e-10 * scale; u-jets within 1e-10 relative to
    the largest derivative of the same point wherever no cancellation is involved."""
    g = resid_ff if problem == "force_free" else resid_kerr
    pts = np.array(g["points"])
    order = 4 if problem == "force_free" else 2
    sess = op.Session.for_problem(problem)
    n_resid = n_jet = 0
    for rec in g["records"]:
        c = op.compile_expr(rec["s"], sess)
        assert c.flags == 0, rec["s"]
        u = J.evaluate(c.whole(), pts, order, sess.const_vals, sess.pow_vals)
        d = J.derivatives(u, order)
        if problem == "force_free":
            R, S, _ = Rz.force_free_residual(u, pts[:, 0])
        else:
            R, S, _ = Rz.kerr_residual(u, pts)
        for k in range(len(pts)):
            gj = rec["jets"][k]
            if any(v is None for v in gj) or not np.isfinite(d[:, k]).all():
                continue
            gj = np.array(gj)
            # value and first derivatives are cancellation free -> pure rtol
            np.testing.assert_allclose(d[:3, k], gj[:3], rtol=1e-10, atol=1e-10 * np.max(np.abs(gj[:3])) + 1e-300)
            n_jet += 1
            gR = rec["R"][k]
            if gR is None or not np.isfinite(R[k]):
                continue
            assert abs(R[k] - gR) <= 1e-10 * S[k] + 1e-300, (rec["s"], k, R[k], gR, S[k])
            n_resid += 1
    assert n_resid > 3000 and n_jet > 3000


def test_finiteness_agrees_with_reference(resid_ff):
    """Where the reference's value is non-real / non-finite the real FP64 jet must not
    be finite (those rows go to the CPU), and vice versa."""
    pts = np.array(resid_ff["points"])
    sess = op.Session.for_problem("force_free")
    for rec in resid_ff["records"]:
        c = op.compile_expr(rec["s"], sess)
        u = J.evaluate(c.whole(), pts, 4, sess.const_vals, sess.pow_vals)
        for k in range(len(pts)):
            ref_fin = all(v is not None for v in rec["jets"][k])
            ours = bool(np.isfinite(u[:, k]).all())
            assert ref_fin == ours, (rec["s"], k)


def test_structured_equals_monomials(resid_ff):
    """The line-by-line jet restatement of FFV:305-347 and the SymPy-expanded monomial
    form are two independent evaluations of the same four matrix entries."""
    pts = np.array(resid_ff["points"])
    sess = op.Session.for_problem("force_free")
    for rec in resid_ff["records"][::7]:
        c = op.compile_expr(rec["s"], sess)
        u = J.evaluate(c.whole(), pts, 4, sess.const_vals, sess.pow_vals)
        R, S, parts = Rz.force_free_residual(u, pts[:, 0])
        sv = Rz.force_free_parts_structured(u, pts[:, 0])
        d = J.derivatives(u, 4)
        mag = np.max(np.abs(d), axis=0)
        for k in range(4):
            ok = np.isfinite(parts[k]) & np.isfinite(sv[k])
            scale = np.maximum(np.abs(parts[k]), 1.0)[ok]
            assert np.all(np.abs(parts[k][ok] - sv[k][ok]) <= 1e-8 * scale * np.maximum(mag[ok] ** 3, 1.0)), rec["s"]


def test_splice_rule_equals_textual_splice(enum_ff):
    """The term-level splice reproduces what sympify makes of the reference's
    un-parenthesised string templates (LBF:170-195) -- depth 2 in full, depth 3 sampled;
    'rho * rho**2 + z**2' must be rho**3 + z**2."""
    E = uniques_by_depth(enum_ff)
    sess = op.Session.for_problem("force_free")
    assert op.to_sympy(op.compile_expr("(rho * rho**2 + z**2)", sess).whole(), sess) == sp.sympify("rho**3 + z**2")
    for depth, step in ((2, 1), (3, 11)):
        cands, triples = oe.candidates_for_depth(E, depth, with_triples=True)
        flat = [s for k in range(1, depth) for s in E[k]]
        comp = [op.compile_expr(s, sess) for s in flat]
        for s, (o, i, j) in list(zip(cands, triples))[::step]:
            code = op.splice(o, comp[i], comp[j] if j >= 0 else None)
            want = sp.sympify(s)
            assert op.to_sympy(code, sess) == want, s
            assert op.to_sympy(op.compile_expr(s, sess).whole(), sess) == want, s
            assert bc.stack_depth(code) >= 1


def test_hash_is_structural():
    a = bc.structural_hash(bytes([bc.OP_VAR0, bc.OP_VAR1, bc.OP_ADD]))
    b = bc.structural_hash(bytes([bc.OP_VAR1, bc.OP_VAR0, bc.OP_ADD]))
    c = bc.structural_hash(bytes([bc.OP_VAR0, bc.OP_VAR1, bc.OP_ADD, 0]))
    assert a != b and a != c
    assert bc.structural_hash(b"") == bc.HASH_SEED

"""Parity of the CUDA jet interpreter (through the C-ABI) against the oracle and the
reference-derived golden vectors.  Tolerance (BASELINE.json north_star): per-point
residuals within 1e-10 relative to the residual's scale S wherever finite; u-jets
within 1e-10 (relative to the largest coefficient magnitude of the jet)."""
import numpy as np
import pytest

from conftest import load_golden, uniques_by_depth
from oracle import jets as J
from oracle import parser as op
from oracle import residuals as Rz

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _setup(problem, P, cuda_device):
    import torch
    import pde_engine_b200 as pb
    from pde_engine_b200.grids import collocation_grid
    sess = pb.Session.for_problem(problem)
    prog = pb.ResidualProgram.for_problem(problem)
    pts = collocation_grid(problem, P)
    tab = prog.point_table(pts)
    return pb, sess, prog, pts, torch.from_numpy(pts).to(cuda_device), torch.from_numpy(tab).to(cuda_device)


def _oracle_eval(problem, strs, pts_soa):
    sess = op.Session.for_problem(problem)
    order = 4 if problem == "force_free" else 2
    pts = np.ascontiguousarray(pts_soa.T)
    out = []
    for s in strs:
        c = op.compile_expr(s, sess)
        if c.flags:
            out.append(None)
            continue
        u = J.evaluate(c.whole(), pts, order, sess.const_vals, sess.pow_vals)
        R, S, _ = (Rz.force_free_residual(u, pts[:, 0]) if problem == "force_free" else Rz.kerr_residual(u, pts))
        out.append((u, R, S))
    return out


_OPAQUE = None


def _exact_jet(problem, s, point, order):
    """Normalised Taylor coefficients of u at one point from SymPy's exact derivatives (50 digits):
    the arbiter when two float64 evaluation orders of an ill-conditioned jet disagree."""
    import math
    import sympy as sp
    global _OPAQUE
    if _OPAQUE is None:   # expression_operations.py:30-60 as sympify locals (GM:85-93)
        _OPAQUE = {"neg": lambda x: -x, "inv": lambda x: 1 / x, "square": lambda x: x ** 2,
                   "pow_3_2": lambda x: x ** sp.Rational(3, 2), "pow_neg_3_2": lambda x: x ** sp.Rational(-3, 2),
                   "exp_neg": lambda x: sp.exp(-x)}
    sess = op.Session.for_problem(problem)
    v0, v1 = (sp.Symbol(n, real=True) for n in sess.var_names)
    loc = dict(_OPAQUE); loc[sess.var_names[0]] = v0; loc[sess.var_names[1]] = v1
    for k, v in sess.named_consts.items():
        loc[k] = sp.nsimplify(v)
    e = op.to_sympy(op.compile_expr(s, sess).whole(), sess, loc)
    at = {v0: sp.Float(float(point[0]), 60), v1: sp.Float(float(point[1]), 60)}
    out = np.zeros(J.ncoef(order))
    for n in range(order + 1):
        for j in range(n + 1):
            d = sp.diff(e, v0, n - j, v1, j) if n else e
            out[J.idx(n - j, j)] = float(sp.re(d.subs(at).evalf(50))) / (math.factorial(n - j) * math.factorial(j))
    return out


def _compare_points(jets, resid, scale, oracle, strs, problem=None, pts=None, s_floor=1e-40, fin_agree=0.995):
    n_cmp = 0
    n_arbitrated = 0
    for i, o in enumerate(oracle):
        if o is None:
            continue
        u, R, S = o
        gj, gR, gS = jets[i], resid[i], scale[i]
        fin_o = np.isfinite(u).all(axis=0)
        fin_g = np.isfinite(gj).all(axis=0)
        # same finiteness pattern (the domain policy: NaN where SymPy goes complex)
        assert (fin_o == fin_g).mean() > fin_agree, strs[i]
        ok = fin_o & fin_g
        if not ok.any():
            continue
        mag = np.max(np.abs(u[:, ok]), axis=0)
        err = np.max(np.abs(gj[:, ok] - u[:, ok]), axis=0)
        # value + first derivatives: cancellation free -> tight; higher orders relative to the jet magnitude
        assert np.all(np.abs(gj[:3, ok] - u[:3, ok]) <= RTOL * np.maximum(np.abs(u[:3, ok]), 1e-3 * mag + 1e-300)), strs[i]
        bad = np.flatnonzero(err > 1e-8 * mag + 1e-300)
        if bad.size:
            # Ill-conditioned jets (a smooth function written through a pole, e.g. inv(z/(1 - rho**2 + z**2))
            # next to the pole of the inner quotient): the oracle's float64 recurrence is itself only
            # accurate to ~1e-8 there, so two correct evaluation orders differ.  Arbitrate with exact
            # derivatives: the device must be as accurate as the float64 oracle (up to a small factor).
            assert problem is not None and bad.size <= 4, (strs[i], bad.size)
            cols = np.flatnonzero(ok)[bad]
            for c in cols:
                ex = _exact_jet(problem, strs[i], pts[:, c], 4 if problem == "force_free" else 2)
                e_dev = np.max(np.abs(gj[:, c] - ex)); e_orc = np.max(np.abs(u[:, c] - ex))
                assert e_dev <= 4 * e_orc + 1e-9 * np.max(np.abs(ex)), (strs[i], c, e_dev, e_orc)
                n_arbitrated += 1
            assert n_arbitrated <= 40
        okr = ok & np.isfinite(R) & np.isfinite(S) & np.isfinite(gR) & np.isfinite(gS) & (S > 0)
        # a (numerically) constant u has derivatives, R and S at pure round-off level: nothing to compare
        magf = np.zeros(u.shape[1])
        magf[ok] = mag
        okr &= S > s_floor * np.maximum(magf, 1.0) ** 6
        assert np.all(np.abs(gR[okr] - R[okr]) <= RTOL * S[okr] * 10 + 1e-300), strs[i]
        assert np.all(np.abs(gS[okr] - S[okr]) <= 1e-6 * S[okr]), strs[i]   # S is only a scale; it inherits the jets' conditioning
        n_cmp += int(okr.sum())
    return n_cmp


@pytest.mark.parametrize("problem", ["force_free", "kerr_magnetosphere"])
def test_eval_points_matches_oracle(problem, cuda_device, enum_ff, enum_kerr):
    import torch
    g = enum_ff if problem == "force_free" else enum_kerr
    E = uniques_by_depth(g)
    strs = E[1] + E[2] + E[3][::(9 if problem == "force_free" else 40)]
    pb, sess, prog, pts, pts_t, tab_t = _setup(problem, 64, cuda_device)
    es = sess.compile(strs)
    code, ln = es.programs(128)
    jets, resid, scale = pb.eval_points(sess, prog, torch.from_numpy(code).to(cuda_device), torch.from_numpy(ln).to(cuda_device),
                                        pts_t, tab_t, None, spill_slots=8)
    torch.cuda.synchronize()
    oracle = _oracle_eval(problem, strs, pts)
    n = _compare_points(jets.cpu().numpy(), resid.cpu().numpy(), scale.cpu().numpy(), oracle, strs, problem, pts)
    assert n > 0.8 * 64 * len(strs) * 0.8


@pytest.mark.parametrize("problem", ["force_free", "kerr_magnetosphere"])
def test_residuals_match_reference_vectors(problem, cuda_device, resid_ff, resid_kerr):
    """CUDA residuals vs the reference's own det_M / _lhs values (evalf(50)) at the
    golden points -- the first 8 points of the product grid."""
    import torch
    g = resid_ff if problem == "force_free" else resid_kerr
    strs = [r["s"] for r in g["records"]]
    pb, sess, prog, pts, pts_t, tab_t = _setup(problem, 64, cuda_device)
    np.testing.assert_array_equal(pts[:, :8].T, np.array(g["points"]))
    es = sess.compile(strs)
    code, ln = es.programs(128)
    assert (ln > 0).all()
    jets, resid, scale = pb.eval_points(sess, prog, torch.from_numpy(code).to(cuda_device), torch.from_numpy(ln).to(cuda_device),
                                        pts_t, tab_t, None, spill_slots=8)
    resid, scale = resid.cpu().numpy(), scale.cpu().numpy()
    n = 0
    for i, rec in enumerate(g["records"]):
        for k in range(8):
            gR = rec["R"][k]
            if gR is None or not np.isfinite(resid[i, k]):
                continue
            assert abs(resid[i, k] - gR) <= RTOL * scale[i, k] * 10 + 1e-300, (rec["s"], k, resid[i, k], gR, scale[i, k])
            n += 1
    assert n > 3000


def test_validate_reduction_consistent_with_points(cuda_device, enum_ff):
    """pde_validate's per-candidate outputs == reducing pde_eval_points' per-point
    output on the host (same kernel, dump vs reduce mode)."""
    import torch
    E = uniques_by_depth(enum_ff)
    strs = E[2] + E[3][::29] + ["zoo*rho", "I*sqrt(rho)"]
    P = 256
    pb, sess, prog, pts, pts_t, tab_t = _setup("force_free", P, cuda_device)
    es = sess.compile(strs)
    code, ln = es.programs(128)
    code_t, len_t = torch.from_numpy(code).to(cuda_device), torch.from_numpy(ln).to(cuda_device)
    tau = 1e-10
    out = pb.validate(sess, prog, code_t, len_t, pts_t, tab_t, None, tau=tau, min_finite=8, vote_frac=0.5, n_ref=3, spill_slots=8)
    _, resid, scale = pb.eval_points(sess, prog, code_t, len_t, pts_t, tab_t, None, spill_slots=8, want_jets=False)
    torch.cuda.synchronize()
    resid, scale = resid.cpu().numpy(), scale.cpu().numpy()
    o = {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}
    bits = o["survivor_bits"].view(np.uint32)
    for i in range(len(strs)):
        surv = (bits[i >> 5] >> (i & 31)) & 1
        if ln[i] == 0:
            assert o["n_finite"][i] == -1 and surv == 1
            continue
        fin = np.isfinite(resid[i]) & np.isfinite(scale[i]) & (scale[i] > 0)
        assert o["n_finite"][i] == fin.sum()
        votes = (np.abs(resid[i][fin]) > tau * scale[i][fin]).sum()
        assert o["n_votes"][i] == votes
        if fin.any():
            ratio = np.abs(resid[i][fin]) / scale[i][fin]
            assert abs(o["ratio_max"][i] - ratio.max()) <= 1e-12 * ratio.max()    # device uses a Newton reciprocal
            assert o["resid_max"][i] == np.abs(resid[i][fin]).max()
        reject = fin.sum() >= 8 and votes > 0 and votes >= 0.5 * fin.sum()
        assert surv == (0 if reject else 1)
        for k in range(3):
            a, b = o["ref_rs"][i, k], (resid[i, k], scale[i, k])
            assert (a[0] == b[0] or (np.isnan(a[0]) and np.isnan(b[0]))) and (a[1] == b[1] or (np.isnan(a[1]) and np.isnan(b[1])))


def test_filter_keeps_every_reference_valid_row(cuda_device):
    """Golden verdicts: rows 1-85 of the reference's committed run DB (58 valid, 27
    invalid) + its validator cache + the 6 known solutions of the sequential report.
    The GPU filter must keep every reference-valid row (is_valid identity is
    mandatory) and should reject the reference-invalid ones."""
    from pde_engine_b200.validator import GpuBatchValidator
    fx = load_golden("ref_fixtures.json")
    rows = [r for r in fx["ff_run_db"] if r["id"] <= 85 and r["reason"] != "constant-only (skipped)"]
    strs = [r["expression"] for r in rows]
    want = [bool(r["is_valid"]) for r in rows]
    for s, v, reason in fx["ff_validator_cache"]:
        if not reason.startswith("Error"):
            strs.append(s)
            want.append(bool(v))
    for r in fx["ff_sequential_report_valid"]:
        strs.append(r["expression"])
        want.append(True)
    gv = GpuBatchValidator(None, "force_free", P=4096)
    bv = gv.prefilter(strs)
    want = np.array(want)
    assert bv.survivor[want].all(), [s for s, w, k in zip(strs, want, bv.survivor) if w and not k]
    # the filter is useful: the reference-invalid rows are (almost all) rejected on device
    assert bv.rejected[~want].mean() > 0.9, [s for s, w, k in zip(strs, want, bv.survivor) if (not w) and k]
    # valid rows have residuals at round-off level relative to the scale
    ok = want & (bv.n_finite > 0)
    assert np.median(bv.ratio_max[ok]) < 1e-12


def test_kerr_primitives_rejected(cuda_device):
    """The committed Kerr run DB rejects all 9 primitives except constants with
    'PDE residual != 0 (fast point check)' (KV:265-271); so does the device."""
    from pde_engine_b200.validator import GpuBatchValidator
    fx = load_golden("ref_fixtures.json")
    rows = fx["kerr_run_db"]
    gv = GpuBatchValidator(None, "kerr_magnetosphere", P=4096)
    strs = [r["expression"] for r in rows]
    bv = gv.prefilter(strs)
    for r, surv, nf in zip(rows, bv.survivor, bv.n_finite):
        if r["reason"] and r["reason"].startswith("PDE residual != 0"):
            assert not surv, r["expression"]
        if r["is_valid"]:
            assert surv


def test_full_size_properties(cuda_device, enum_ff):
    """BASELINE-size property checks (P = 4096, all 143 461 depth-4 uniques):
    linearity of the residual scale (u -> c*u leaves |R|/S invariant), idempotence
    (same input, same bits), and every known solution survives."""
    import torch
    E = uniques_by_depth(enum_ff)
    strs = E[4]
    pb, sess, prog, pts, pts_t, tab_t = _setup("force_free", 4096, cuda_device)
    es = sess.compile(strs)
    code, ln = es.programs(128)
    code_t, len_t = torch.from_numpy(code).to(cuda_device), torch.from_numpy(ln).to(cuda_device)
    a = pb.validate(sess, prog, code_t, len_t, pts_t, tab_t, None, spill_slots=4)
    b = pb.validate(sess, prog, code_t, len_t, pts_t, tab_t, None, spill_slots=4)
    torch.cuda.synchronize()
    for k in ("ratio_max", "resid_max", "n_finite", "n_votes", "survivor_bits"):
        x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
        assert np.array_equal(x, y, equal_nan=(x.dtype.kind == "f")), k
    nf = a["n_finite"].cpu().numpy()
    assert (nf >= -3).all() and (nf <= 4096).all()
    frac_evaluated = (nf >= 0).mean()
    assert frac_evaluated > 0.99
    # Bent solution variants found at depth 4 survive
    bits = a["survivor_bits"].cpu().numpy().view(np.uint32)
    for s in ("square(rho*exp(-z))", "rho**2/pow_3_2(rho**2 + z**2)", "-z/sqrt(rho**2 + z**2) + 1"):
        if s in strs:
            i = strs.index(s)
            assert (bits[i >> 5] >> (i & 31)) & 1, s


UNIVARIATE = {
    "force_free": ["sqrt(exp(rho))", "exp(-sqrt(rho))", "1/exp(rho**2)", "sqrt(rho)**(3/2)", "exp(exp(-z))", "(z**2)**(-3/2)",
                   "sqrt(1 + z**2)", "exp(1/(1 + rho))", "(1 + exp(rho))**2", "exp(sqrt(2))", "1/sqrt(exp(1))",
                   "sqrt(exp(rho))*exp(1/z)", "exp(exp(rho))/(1 + exp(-z))", "sqrt(exp(rho) + exp(sqrt(rho)))",
                   "exp(sqrt(rho*rho + 1))**2", "sqrt(exp(z))/(1 + sqrt(exp(z)))",
                   "z/exp(exp(z/rho))", "rho/exp(exp(z/rho))"],      # coordinate / T: quotient recurrence (relative derivatives ~1e13)
    "kerr_magnetosphere": ["sqrt(exp(r))", "exp(-sqrt(r))", "exp(exp(x))", "1/exp(x**2)", "sqrt(exp(r))*exp(x)", "(1 + exp(r))**2"],
}


@pytest.mark.parametrize("problem", ["force_free", "kerr_magnetosphere"])
def test_single_axis_bodies(problem, cuda_device):
    """Sub-expressions of one coordinate run through the single-axis sqrt / square / composition bodies
    (translate pass 3): on-axis coefficients against the oracle, off-axis coefficients exactly zero, and products
    of a rho-part with a z-part (general bodies fed by single-axis results) against the oracle too."""
    import torch
    strs = UNIVARIATE[problem]
    pb, sess, prog, pts, pts_t, tab_t = _setup(problem, 64, cuda_device)
    es = sess.compile(strs)
    assert not es.flags().any()
    code, ln = es.programs(128)
    jets, resid, scale = pb.eval_points(sess, prog, torch.from_numpy(code).to(cuda_device), torch.from_numpy(ln).to(cuda_device),
                                        pts_t, tab_t, None, spill_slots=8)
    torch.cuda.synchronize()
    jets = jets.cpu().numpy()
    oracle = _oracle_eval(problem, strs, pts)
    n = _compare_points(jets, resid.cpu().numpy(), scale.cpu().numpy(), oracle, strs, problem, pts)
    assert n > 0
    order = 4 if problem == "force_free" else 2
    v0, v1 = sess.var_names if hasattr(sess, "var_names") else (("rho", "z") if problem == "force_free" else ("r", "x"))
    for i, s in enumerate(strs):
        has0, has1 = (v0 in s), (v1 in s.replace("exp", "").replace("sqrt", "")) if v1 == "x" else (v1 in s)
        if has0 and has1:
            continue
        for a in range(order + 1):
            for b in range(order + 1 - a):
                off_axis = (b > 0) if has0 else (a > 0) if has1 else (a + b > 0)
                if off_axis:
                    assert (jets[i, J.idx(a, b)] == 0).all(), (s, a, b)       # structural zeros stay exact zeros


def _random_expr(rng, depth, vars_, in_exp=False):
    """Random expression string over the reference's vocabulary after normalisation: + - * /,
    rational powers, sqrt, exp, Abs, small rational constants.  No exp inside an exp: with relative derivatives of 1e13
    (`2/exp(exp(z/rho))`) 1/T by composition loses 8 digits against the oracle's quotient recurrence (values of 1e-51
    whose residual underflows anyway)."""
    if depth == 0 or rng.random() < 0.15:
        r = rng.random()
        if r < 0.4:
            return vars_[0]
        if r < 0.8:
            return vars_[1]
        return rng.choice(["1", "2", "3", "1/2", "1/3", "3/2", "5"])
    k = rng.random()
    if in_exp and 0.8 <= k < 0.9:
        k = 0.75
    a = _random_expr(rng, depth - 1, vars_, in_exp or (0.8 <= k < 0.9))
    if k < 0.5:
        b = _random_expr(rng, depth - 1, vars_, in_exp)
        op = rng.choice(["+", "-", "*", "/"])
        if op == "-" and a == b:            # a literal zero: `0**3` is finite in the oracle and NaN on the device
            op = "+"
        return f"({a} {op} {b})"
    if k < 0.7:
        return f"({a})**({rng.choice(['2', '3', '-1', '-2', '1/2', '3/2', '-3/2', '-1/2', '5/2'])})"
    if k < 0.8:
        return f"sqrt({a})"
    if k < 0.9:
        return f"exp({rng.choice(['', '-'])}({a}))"
    if k < 0.95:
        return f"Abs({a})"
    return f"-({a})"


@pytest.mark.parametrize("problem,seed", [("force_free", 11), ("force_free", 12), ("kerr_magnetosphere", 13)])
def test_random_expressions_match_oracle(problem, seed, cuda_device):
    """Fuzz: 400 random depth <= 4 expression strings through the product compiler + interpreter against the
    oracle's parser + float64 jets (same tolerances as the golden-vector tests)."""
    import random
    import torch
    rng = random.Random(seed)
    pb, sess, prog, pts, pts_t, tab_t = _setup(problem, 64, cuda_device)
    vars_ = ("rho", "z") if problem == "force_free" else ("r", "x")
    strs = []
    while len(strs) < 400:
        s = _random_expr(rng, rng.choice([2, 3, 4]), vars_)
        if any(v in s for v in vars_):
            strs.append(s)
    es = sess.compile(strs)
    flags = es.flags()
    code, ln = es.programs(128)
    jets, resid, scale = pb.eval_points(sess, prog, torch.from_numpy(code).to(cuda_device), torch.from_numpy(ln).to(cuda_device),
                                        pts_t, tab_t, None, spill_slots=8)
    torch.cuda.synchronize()
    oracle = _oracle_eval(problem, strs, pts)
    # product and oracle agree on what is compilable
    assert [o is None for o in oracle] == [bool(f) for f in flags]
    # strings that cancel to a constant (`Abs(sqrt(z**2)) - z`, `x/(x + x)`) have jets at round-off level: noise on both sides
    for i, o in enumerate(oracle):
        if o is not None:
            u = o[0]
            fin = np.isfinite(u).all(axis=0)
            if fin.any() and np.abs(u[1:, fin]).max() <= 1e-9 * max(np.abs(u[0, fin]).max(), 1.0):
                oracle[i] = None
    # s_floor: random strings such as `exp(-z/2) - rho*(z/rho)` depend on one coordinate only up to round-off; their
    # R and S are products of 1e-16-sized derivatives (1e-28 and below) -- noise in both implementations
    n = _compare_points(jets.cpu().numpy(), resid.cpu().numpy(), scale.cpu().numpy(), oracle, strs, problem, pts, s_floor=1e-20,
                        fin_agree=0.9)   # overflow corners (`z/exp(exp(z/rho))`: 1/inf vs inf*0) may differ at a few points
    assert n > 64 * 100

"""The round-off majorant calculus (oracle/majorant.py), pinned on the CPU against the reference-derived golden
vectors: the float64 oracle and the reference's 50-digit values are two evaluations of the same expression, so
they must agree within the bounds the calculus promises; and every reference-valid row must stay below the
decision threshold at EVERY finite point (no vote at all, not merely a minority of votes)."""
import json
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import jets as J
from oracle import majorant as Mj
from oracle import parser as op
from oracle import residuals as Rz

TAU = 1e-10


@pytest.mark.parametrize("problem", ["force_free", "kerr_magnetosphere"])
def test_bounds_hold_against_the_reference_values(problem, resid_ff, resid_kerr):
    g = resid_ff if problem == "force_free" else resid_kerr
    pts = np.array(g["points"])
    order = 4 if problem == "force_free" else 2
    sess = op.Session.for_problem(problem)
    mi = J.multi_indices(order)
    fact = np.array([math.factorial(i) * math.factorial(j) for i, j in mi], dtype=np.float64)
    deg = np.array([i + j for i, j in mi])
    t0 = Mj.T0_DEFAULT
    n_j = n_r = 0
    worst_j = worst_r = 0.0
    for rec in g["records"]:
        c = op.compile_expr(rec["s"], sess)
        with np.errstate(all="ignore"):
            u, V, D, W = Mj.evaluate(c.whole(), pts, order, sess.const_vals, sess.pow_vals)
            if problem == "force_free":
                R, S, _ = Rz.force_free_residual(u, pts[:, 0])
                St = Mj.force_free_scale(u, pts[:, 0], W, TAU)
            else:
                R, S, _ = Rz.kerr_residual(u, pts)
                St = Mj.kerr_scale(u, pts, W, TAU)
        assert np.all(St[np.isfinite(St) & np.isfinite(S)] >= S[np.isfinite(St) & np.isfinite(S)] * (1 - 1e-12))      # S~ majorises S
        for k in range(len(pts)):
            gj = rec["jets"][k]
            if any(v is None for v in gj) or not np.isfinite(u[:, k]).all() or not np.isfinite(W[k]):
                continue
            cg = np.array(gj) / fact                       # the reference's partial derivatives -> Taylor coefficients
            bound = 2.0 * Mj.EPS * W[k] / t0 ** deg + 1e-300
            q = np.abs(u[:, k] - cg) / bound
            assert np.all(q <= 1.0), (rec["s"], k, q.max())
            worst_j = max(worst_j, float(q.max()))
            n_j += 1
            gR = rec["R"][k]
            if gR is None or not np.isfinite(R[k]) or not np.isfinite(St[k]) or St[k] <= 0:
                continue
            assert abs(R[k] - gR) <= 2.0 * TAU * St[k], (rec["s"], k, R[k], gR, St[k])
            worst_r = max(worst_r, abs(R[k] - gR) / (2.0 * TAU * St[k]))
            n_r += 1
    print(f"{problem}: {n_j} jets within the W bound (worst {worst_j:.3g} of it), {n_r} residuals within 2 tau S~ (worst {worst_r:.3g} of it)")
    assert n_j > 2500 and n_r > 2000


def _ff_votes(s, pts, sess):
    c = op.compile_expr(s, sess)
    if c.flags:
        return None
    with np.errstate(all="ignore"):
        u, V, D, W = Mj.evaluate(c.whole(), pts, 4, sess.const_vals, sess.pow_vals)
        R, S, _ = Rz.force_free_residual(u, pts[:, 0])
        St = Mj.force_free_scale(u, pts[:, 0], W, TAU)
        fin = np.isfinite(R) & np.isfinite(St) & (St > 0)
        fin_plain = np.isfinite(R) & np.isfinite(S) & (S > 0)
        return (int(fin.sum()), int((fin & (np.abs(R) > TAU * St)).sum()),
                int(fin_plain.sum()), int((fin_plain & (np.abs(R) > TAU * S)).sum()))


def test_reference_valid_rows_never_vote():
    """Every row the reference validates (committed run DB ids 1-85, its validator cache, the sequential report, the
    depth-2/3 verdict fixtures generated from the unmodified reference) has |R| <= tau S~ at every finite point of a
    64-point grid, and the reference-invalid rows are still (almost all) rejected by the majority rule."""
    sess = op.Session.for_problem("force_free")
    pts = Rz.collocation_grid("force_free", 64)
    fx = load_golden("ref_fixtures.json")
    rows = [(r["expression"], bool(r["is_valid"])) for r in fx["ff_run_db"] if r["id"] <= 85 and r["reason"] != "constant-only (skipped)"]
    rows += [(s, bool(v)) for s, v, reason in fx["ff_validator_cache"] if not reason.startswith("Error")]
    rows += [(r["expression"], True) for r in fx["ff_sequential_report_valid"]]
    for f in ("verdicts_force_free_d2.json", "verdicts_force_free_d3.json"):
        for r in json.load(open(os.path.join(GOLDEN, f)))["records"]:
            reason = r.get("reason", "")
            if "is_valid" in r and not reason.startswith("Error") and not reason.startswith("constant-only"):
                rows.append((r["s"], bool(r["is_valid"])))
    rows = list(dict(rows).items())[::1 if len(rows) < 1500 else 3]      # the full depth-3 verdict set is sampled here (CPU time); the GPU test takes all of it
    n_valid = n_invalid = n_rejected = 0
    for s, valid in rows:
        v = _ff_votes(s, pts, sess)
        if v is None:
            continue
        nf, nv, _, _ = v
        if valid:
            assert nv == 0, (s, v)
            n_valid += 1
        else:
            n_invalid += 1
            n_rejected += nf >= 8 and nv >= 0.5 * nf
    assert n_valid > 150 and n_rejected >= 0.9 * n_invalid, (n_valid, n_invalid, n_rejected)


ILL_CONDITIONED_TRUE_SOLUTIONS = [
    # cancellation inside u's own jet that the normaliser cannot see through the opaque names (LB:73); the
    # reference validates each as the function on the right
    "z*inv(z)/rho",                                   # 1/rho: u_z = O(1e-16) instead of 0 -- rejected by round 1's rule
    "rho*inv(rho/z)",                                 # z
    "rho**2*z + exp(rho*z) - exp_neg(neg(rho*z))",    # rho**2*z (X-point) + an exact zero written two ways
    "rho**2*exp(-2*z) + square(rho/z) - rho**2/z**2",           # bent solution + 0
    "sqrt(rho**2 + z**2) - z + exp(exp(rho)) - exp_neg(neg(exp(rho)))",   # parabolic + 0 of magnitude up to 1600
    "1 - z/sqrt(rho**2 + z**2) + pow_3_2(rho + z) - sqrt(rho + z)*(rho + z)",   # radial + 0
    "rho**2/pow_3_2(rho**2 + z**2) + inv(rho - z) - 1/(rho - z)",      # dipolar + 0 with a pole on the diagonal of the grid
]


def test_ill_conditioned_true_solutions_never_vote():
    """known solutions plus cancelling sub-expressions, also on a grid skewed to rho / z ~ 1e-3: the plain
    scale S misses the cancellation (votes on the plain scale are reported), the majorant scale never votes."""
    sess = op.Session.for_problem("force_free")
    grids = [Rz.collocation_grid("force_free", 64)]
    skew = grids[0].copy()
    skew[:, 0] = 1e-3 * skew[:, 0]                    # rho / z ~ 1e-3
    grids.append(skew)
    plain_votes = 0
    for pts in grids:
        for s in ILL_CONDITIONED_TRUE_SOLUTIONS:
            nf, nv, nfp, nvp = _ff_votes(s, pts, sess)
            assert nv == 0, (s, nf, nv)
            plain_votes += nvp
    assert plain_votes > 0         # the hole is real: the plain scale does vote on some of these

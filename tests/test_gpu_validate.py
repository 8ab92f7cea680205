"""Parity of the CUDA jet interpreter (through the C-ABI) against the oracle and the
reference-derived golden vectors.

Tolerances (BASELINE.json north_star: per-point residuals agree with the reference's float64
evaluation to rtol 1e-10 wherever finite), with no statistical allowance and no arbitration:

  residual   |R_dev - R_ref| <= 1e-10 * S      at EVERY finite point where the plain scale S (sum of
                                               |monomial|) is not itself pure round-off: S >= 1e-12 * S~
             |R_dev - R_ref| <= 2e-10 * S~     at EVERY finite point, no exception (S~ = the decision scale
                                               with the round-off majorant, include/pde_b200.h; for
                                               `z*inv(z)/rho` S is 1e-30 and both R's are noise)
  jets       |c_dev - c_ref| <= 1e-10 * mag    (mag = largest coefficient of the jet) at >= 99.9 % of the points, and
             |c_dev - c_ref| <= 2 eps W / t0^|g|   at EVERY finite point: two float64 evaluation orders of the
                                               same expression differ by at most the bound the device carries
The second line of each pair is the majorant theory itself under test (oracle/majorant.py)."""
import math

import numpy as np
import pytest

from conftest import load_golden, uniques_by_depth
from oracle import jets as J
from oracle import majorant as Mj
from oracle import parser as op
from oracle import residuals as Rz

pytestmark = pytest.mark.gpu

RTOL = 1e-10
TAU = 1e-10
NOISE = 1e-12        # S below NOISE * S~: the plain scale is pure round-off, nothing to compare on it


def _setup(problem, P, cuda_device):
    import torch
    import pde_engine_b200 as pb
    from pde_engine_b200.grids import collocation_grid
    sess = pb.Session.for_problem(problem)
    prog = pb.ResidualProgram.for_problem(problem)
    pts = collocation_grid(problem, P)
    tab = prog.point_table(pts)
    return pb, sess, prog, pts, torch.from_numpy(pts).to(cuda_device), torch.from_numpy(tab).to(cuda_device)


def _device_eval(pb, sess, prog, strs, pts_t, tab_t, cuda_device, prim=None, code=None):
    """Full per-point device output for a list of strings: dict of numpy arrays + the compile flags."""
    import torch
    if code is None:
        es = sess.compile(strs)
        c, ln = es.programs(128)
        flags = es.flags()
        code_t, len_t = torch.from_numpy(c).to(cuda_device), torch.from_numpy(ln).to(cuda_device)
    else:
        code_t, len_t = code
        flags = np.zeros(code_t.shape[0], np.uint8)
    jets, R, S, St, maj = pb.eval_points(sess, prog, code_t, len_t, pts_t, tab_t, prim, spill_slots=8, want_maj=True, tau=TAU)
    torch.cuda.synchronize()
    return dict(jets=jets.cpu().numpy(), R=R.cpu().numpy(), S=S.cpu().numpy(), St=St.cpu().numpy(), maj=maj.cpu().numpy()), flags


def _oracle_eval(problem, strs, pts_soa, prims=None, prim_maj=None):
    """(u, R, S, S~, V, D, W) per string (None where the oracle's compiler flags it), programs may be bytes."""
    sess = op.Session.for_problem(problem)
    order = 4 if problem == "force_free" else 2
    pts = np.ascontiguousarray(pts_soa.T)
    out = []
    for s in strs:
        if isinstance(s, str):
            c = op.compile_expr(s, sess)
            if c.flags:
                out.append(None)
                continue
            prog = c.whole()
        else:
            prog = s
        with np.errstate(all="ignore"):
            u, V, D, W = Mj.evaluate(prog, pts, order, sess.const_vals, sess.pow_vals, prims or (), prim_maj or ())
            if problem == "force_free":
                R, S, _ = Rz.force_free_residual(u, pts[:, 0])
                St = Mj.force_free_scale(u, pts[:, 0], W, TAU)
            else:
                R, S, _ = Rz.kerr_residual(u, pts)
                St = Mj.kerr_scale(u, pts, W, TAU)
        out.append((u, R, S, St, V, D, W))
    return out


def _compare_points(dev, oracle, strs, order, fin_agree=0.995, report=None):
    """The four assertions of the module docstring on every string; returns the number of residual points compared."""
    deg = np.array([i + j for i, j in J.multi_indices(order)])
    t0 = Mj.T0_DEFAULT
    n_R = n_J = n_J_tight = 0
    w_ratio = []
    st_ratio = []
    n_fin_pts = n_fin_diff = 0
    worst = dict(R_over_S=0.0, R_over_St=0.0, J_over_mag=0.0, J_over_bound=0.0)
    for i, o in enumerate(oracle):
        if o is None:
            continue
        u, R, S, St, V, D, W = o
        gj, gR, gS, gSt, gW = dev["jets"][i], dev["R"][i], dev["S"][i], dev["St"][i], dev["maj"][i, 2].astype(np.float64)
        with np.errstate(all="ignore"):
            fin_o = np.isfinite(u).all(axis=0)
            fin_g = np.isfinite(gj).all(axis=0)
            # same finiteness pattern (the domain policy: NaN where SymPy goes complex).  The two arithmetics overflow
            # in different INTERMEDIATES (the device forms F's scalar Taylor coefficients x0**(k - j) before the
            # composition, the oracle runs the quotient recurrence on the jet: `exp(big)**(-3/2)` is inf * 0 = NaN on
            # the device and 0 in numpy), so single points of a string may differ; globally they are rare (below)
            assert (fin_o == fin_g).mean() > min(fin_agree, 0.9), strs[i]
            n_fin_pts += fin_o.size
            n_fin_diff += int((fin_o != fin_g).sum())
            ok = fin_o & fin_g & np.isfinite(gW)
            if not ok.any():
                continue
            mag = np.max(np.abs(u), axis=0)
            dj = np.abs(gj - u)
            # jets, every point: within the round-off majorant the DEVICE carries
            bound = 2.0 * Mj.EPS * gW[None, :] / (t0 ** deg)[:, None]
            viol = ok[None, :] & (dj > bound)
            assert not viol.any(), (strs[i], np.argwhere(viol)[:3].tolist())
            rel = np.max(dj, axis=0) / np.where(mag > 0, mag, 1.0)
            n_J += int(ok.sum())
            n_J_tight += int((ok & (rel <= RTOL)).sum())
            worst["J_over_mag"] = max(worst["J_over_mag"], float(np.max(rel[ok])))
            with np.errstate(invalid="ignore"):
                q = np.where(bound > 0, dj / bound, 0.0)[:, ok]
            worst["J_over_bound"] = max(worst["J_over_bound"], float(np.max(q)))
            # residual, every point: within tau * S~; where S is not noise: within 1e-10 * S
            okr = ok & np.isfinite(R) & np.isfinite(gR) & np.isfinite(gSt) & (gSt > 0)
            dR = np.abs(gR - R)
            assert np.all(dR[okr] <= 2.0 * TAU * gSt[okr]), strs[i]
            sharp = okr & np.isfinite(S) & (S > 0) & (S >= NOISE * gSt)
            assert np.all(dR[sharp] <= RTOL * S[sharp]), (strs[i], float(np.max(dR[sharp] / S[sharp])))
            assert np.all(np.abs(gS[sharp] - S[sharp]) <= 1e-6 * S[sharp]), strs[i]   # S is only a scale; it inherits the jets' conditioning
            if okr.any():
                worst["R_over_St"] = max(worst["R_over_St"], float(np.max(dR[okr] / gSt[okr])))
            if sharp.any():
                worst["R_over_S"] = max(worst["R_over_S"], float(np.max(dR[sharp] / S[sharp])))
            n_R += int(sharp.sum())
            # the device's majorants are the oracle's rules in float32: never materially below them, rarely far above
            fw = ok & np.isfinite(W) & (W > 1e-25) & (W < 1e30)
            if fw.any():
                # float32 rules: a pole distance `a - D` computed next to the radius of convergence cancels, so single points
                # may sit a few per cent off the float64 rules in either direction (they carry a huge W and do not vote);
                # the tightness statement is therefore global (below), the per-string one only excludes a gross underestimate
                assert np.all(gW[fw] >= 0.9 * W[fw]), strs[i]
                w_ratio.append(gW[fw] / W[fw])
            fs = fw & np.isfinite(St) & np.isfinite(gSt) & (St > 0)
            if fs.any():                      # the decision scale: the same polynomial of the same inflated partials; S~ is of
                # degree <= 6 in theta, so the device's safety factors (0.2 % on theta, 1e-4 per majorant rule) and the float32
                # cancellation next to a radius of convergence show 6-fold: never materially below the oracle's (per string),
                # close to it on the bulk of the points (global, below)
                assert np.all(gSt[fs] >= 0.9 * St[fs]), strs[i]
                st_ratio.append(gSt[fs] / St[fs])
    assert n_J_tight >= 0.999 * n_J, (n_J_tight, n_J)
    assert n_fin_diff <= (1.0 - fin_agree) * 0.1 * n_fin_pts, (n_fin_diff, n_fin_pts)      # e.g. <= 0.05 % of the points at fin_agree = 0.995
    if w_ratio:
        wr = np.concatenate(w_ratio)
        # 1e-4 safety factor per rule + MUFU approximations: the device's W is the oracle's to ~1 % on the bulk of the points
        assert np.median(np.abs(wr - 1.0)) < 2e-2 and np.mean(wr < 0.98) < 1e-3 and np.mean(wr > 1.5) < 1e-2, \
            (float(np.median(np.abs(wr - 1.0))), float(np.mean(wr < 0.98)), float(np.mean(wr > 1.5)))
    if st_ratio:
        sr = np.concatenate(st_ratio)
        assert np.median(sr) < 1.1 and np.mean(sr > 2.0) < 2e-2, (float(np.median(sr)), float(np.mean(sr > 2.0)))
    if report is not None:
        report.update(worst, n_R=n_R, n_J=n_J, n_J_tight=n_J_tight)
    print(f"parity: {n_R} residual points, worst |dR|/S {worst['R_over_S']:.2e}, |dR|/S~ {worst['R_over_St']:.2e}; "
          f"{n_J} jet points ({n_J - n_J_tight} beyond 1e-10*mag, worst {worst['J_over_mag']:.2e}), worst |dc|/bound {worst['J_over_bound']:.2e}")
    return n_R


@pytest.mark.parametrize("problem", ["force_free", "kerr_magnetosphere"])
def test_eval_points_matches_oracle(problem, cuda_device, enum_ff, enum_kerr):
    g = enum_ff if problem == "force_free" else enum_kerr
    E = uniques_by_depth(g)
    strs = E[1] + E[2] + E[3][::(9 if problem == "force_free" else 40)]
    pb, sess, prog, pts, pts_t, tab_t = _setup(problem, 64, cuda_device)
    dev, _ = _device_eval(pb, sess, prog, strs, pts_t, tab_t, cuda_device)
    oracle = _oracle_eval(problem, strs, pts)
    n = _compare_points(dev, oracle, strs, 4 if problem == "force_free" else 2)
    assert n > 0.8 * 64 * len(strs) * 0.8


@pytest.mark.parametrize("problem", ["force_free", "kerr_magnetosphere"])
def test_residuals_match_reference_vectors(problem, cuda_device, resid_ff, resid_kerr):
    """CUDA residuals AND jets of every order vs the reference's own det_M / _lhs / repeated-diff values
    (evalf(50)) at the golden points -- the first 8 points of the product grid: |dR| <= 1e-10 * S, every partial
    derivative of orders 0..4 within 1e-10 of the jet's magnitude."""
    g = resid_ff if problem == "force_free" else resid_kerr
    strs = [r["s"] for r in g["records"]]
    order = 4 if problem == "force_free" else 2
    pb, sess, prog, pts, pts_t, tab_t = _setup(problem, 64, cuda_device)
    np.testing.assert_array_equal(pts[:, :8].T, np.array(g["points"]))
    dev, flags = _device_eval(pb, sess, prog, strs, pts_t, tab_t, cuda_device)
    assert not flags.any()
    fact = np.array([math.factorial(i) * math.factorial(j) for i, j in J.multi_indices(order)], dtype=np.float64)
    n = n_jet = 0
    worst_R = worst_J = 0.0
    for i, rec in enumerate(g["records"]):
        for k in range(8):
            gj = rec["jets"][k]
            dj = dev["jets"][i, :, k] * fact          # normalised Taylor coefficients -> partial derivatives
            if all(v is not None for v in gj) and np.isfinite(dj).all():
                gj = np.array(gj)
                mag = np.max(np.abs(gj))
                err = np.max(np.abs(dj - gj))
                assert err <= RTOL * mag + 1e-300, (rec["s"], k, err, mag)
                worst_J = max(worst_J, err / mag if mag > 0 else 0.0)
                n_jet += 1
            gR = rec["R"][k]
            if gR is None or not np.isfinite(dev["R"][i, k]):
                continue
            S = dev["S"][i, k]
            assert abs(dev["R"][i, k] - gR) <= RTOL * S + 1e-300, (rec["s"], k, dev["R"][i, k], gR, S)
            if S > 0:
                worst_R = max(worst_R, abs(dev["R"][i, k] - gR) / S)
            n += 1
    print(f"golden {problem}: {n} residuals, worst |dR|/S {worst_R:.2e}; {n_jet} jets (all orders), worst |dd|/mag {worst_J:.2e}")
    assert n > 2500 and n_jet > 3000


def test_validate_reduction_consistent_with_points(cuda_device, enum_ff):
    """pde_validate's per-candidate outputs == reducing pde_eval_points' per-point output on the host (same
    kernel, dump vs reduce mode), in the one-pass mode (majorants on every point) and in the two-pass mode
    (proposals on the plain isotropic scale over all points, confirmation with majorants on the first 128)."""
    import torch
    E = uniques_by_depth(enum_ff)
    strs = E[2] + E[3][::29] + ["zoo*rho", "I*sqrt(rho)", "z*inv(z)/rho"]
    P = 256
    pb, sess, prog, pts, pts_t, tab_t = _setup("force_free", P, cuda_device)
    es = sess.compile(strs)
    code, ln = es.programs(128)
    code_t, len_t = torch.from_numpy(code).to(cuda_device), torch.from_numpy(ln).to(cuda_device)
    tau = 1e-10
    _, resid, _, scale, _ = pb.eval_points(sess, prog, code_t, len_t, pts_t, tab_t, None, spill_slots=8, want_jets=False, want_maj=True, tau=tau)
    torch.cuda.synchronize()
    resid, scale = resid.cpu().numpy(), scale.cpu().numpy()

    def host_rule(i, npts):
        r, sc = resid[i, :npts], scale[i, :npts]
        fin = np.isfinite(r) & np.isfinite(sc) & (sc > 0)
        votes = int((np.abs(r[fin]) > tau * sc[fin]).sum())
        return fin, votes, bool(fin.sum() >= 8 and votes > 0 and votes >= 0.5 * fin.sum())

    # one pass: everything is the reduction of the per-point (R, S~)
    out = pb.validate(sess, prog, code_t, len_t, pts_t, tab_t, None, tau=tau, min_finite=8, vote_frac=0.5, confirm_points=0, n_ref=3, spill_slots=8)
    torch.cuda.synchronize()
    o = {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}
    bits = o["survivor_bits"].view(np.uint32)
    for i in range(len(strs)):
        surv = (bits[i >> 5] >> (i & 31)) & 1
        if ln[i] == 0:
            assert o["n_finite"][i] == -1 and surv == 1
            continue
        fin, votes, reject = host_rule(i, P)
        assert o["n_finite"][i] == fin.sum()
        assert o["n_votes"][i] == votes
        if fin.any():
            ratio = np.abs(resid[i][fin]) / scale[i][fin]
            assert abs(o["ratio_max"][i] - ratio.max()) <= 1e-12 * ratio.max()    # device uses a Newton reciprocal
            assert o["resid_max"][i] == np.abs(resid[i][fin]).max()
        assert surv == (0 if reject else 1)
        for k in range(3):
            a, b = o["ref_rs"][i, k], (resid[i, k], scale[i, k])
            assert (a[0] == b[0] or (np.isnan(a[0]) and np.isnan(b[0]))) and (a[1] == b[1] or (np.isnan(a[1]) and np.isnan(b[1])))
    one_pass_survivor = [int((bits[i >> 5] >> (i & 31)) & 1) for i in range(len(strs))]

    # two passes: a rejection needs the proposal AND the confirmation on the first 128 points
    out2 = pb.validate(sess, prog, code_t, len_t, pts_t, tab_t, None, tau=tau, min_finite=8, vote_frac=0.5, confirm_points=128, n_ref=3, spill_slots=8)
    torch.cuda.synchronize()
    o2 = {k: (v.cpu().numpy() if v is not None else None) for k, v in out2.items()}
    bits2 = o2["survivor_bits"].view(np.uint32)
    n_rej = 0
    for i in range(len(strs)):
        surv = int((bits2[i >> 5] >> (i & 31)) & 1)
        if ln[i] == 0:
            assert surv == 1
            continue
        nf1, nv1 = int(o2["n_finite"][i]), int(o2["n_votes"][i])
        proposed = nf1 >= 8 and nv1 > 0 and nv1 >= 0.5 * nf1
        cf = o2["confirm"][i]
        if not proposed:
            assert surv == 1 and cf[0] == -1 and cf[1] == -1        # never re-examined
            continue
        fin, votes, reject = host_rule(i, 128)
        assert (int(cf[0]), int(cf[1])) == (int(fin.sum()), votes), strs[i]
        assert surv == (0 if reject else 1), strs[i]
        n_rej += 1 - surv
    assert n_rej > 20
    i = strs.index("z*inv(z)/rho")       # the reference validates it as 1/rho; its R and S are both round-off
    assert one_pass_survivor[i] == 1 and (bits2[i >> 5] >> (i & 31)) & 1


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


def test_ill_conditioned_true_solutions(cuda_device):
    """exact solutions written so that float64 loses most of its digits INSIDE u's own jet -- cancelling
    sub-expressions, large offsets, exponent shifts -- on the product grid and on a grid skewed to rho/z ~ 1e-3.
    Their residual is noise of the size of the plain scale S (the round-1 rule rejects several of them); the majorant
    rule must let every one of them survive, in both filter modes."""
    import torch
    import pde_engine_b200 as pb
    from pde_engine_b200.grids import collocation_grid
    sols = ["rho**2*z", "rho**2*exp(-2*z)", "sqrt(rho**2 + z**2) - z", "rho**2/(rho**2 + z**2)**(3/2)", "1 - z/sqrt(rho**2 + z**2)", "1/rho"]
    disguised = []
    for u in sols:
        disguised += [f"({u} + exp(z)) - exp(z)", f"({u} + 1000000) - 1000000", f"({u})*(rho + 1000)/(rho + 1000)",
                      f"({u} + rho/z) - rho/z", f"(({u}) + 1/(rho - 1)) - 1/(rho - 1)"]
    disguised += ["z*inv(z)/rho", "rho**2*exp(-2*z + 30)/exp(30)", "(rho + z) - rho", "exp(z + rho)/exp(rho)*0 + rho**2*z",
                  "sqrt((rho**2 + z**2)**2)/sqrt(rho**2 + z**2) - z"]
    sess = pb.Session.for_problem("force_free")
    prog = pb.ResidualProgram.for_problem("force_free")
    es = sess.compile(disguised)
    assert not es.flags().any()
    c, ln = es.programs(128)
    code_t, len_t = torch.from_numpy(c).to(cuda_device), torch.from_numpy(ln).to(cuda_device)
    P = 1024
    grids = {"product": collocation_grid("force_free", P)}
    skew = collocation_grid("force_free", P).copy()
    skew[0] = 1e-3 * (1.0 + skew[0])                      # rho in [1.25e-3, 3e-3], z in [0.25, 2]
    grids["skewed rho/z ~ 1e-3"] = skew
    n_plain_rejects = 0
    for name, pts in grids.items():
        pts_t = torch.from_numpy(pts).to(cuda_device)
        tab_t = torch.from_numpy(prog.point_table(pts)).to(cuda_device)
        for cp in (0, 128):                                # one pass with majorants / two passes
            out = pb.validate(sess, prog, code_t, len_t, pts_t, tab_t, None, tau=TAU, min_finite=8, vote_frac=0.5,
                              confirm_points=cp, n_ref=0, spill_slots=4)
            bits = out["survivor_bits"].cpu().numpy().view(np.uint32)
            surv = np.array([(bits[i >> 5] >> (i & 31)) & 1 for i in range(len(disguised))], bool)
            assert surv.all(), (name, cp, [s for s, k in zip(disguised, surv) if not k])
        # what the plain scale S (no round-off majorant) would have decided
        _, R, S, St, _ = pb.eval_points(sess, prog, code_t, len_t, pts_t, tab_t, None, spill_slots=4, want_jets=False, want_maj=True, tau=TAU)
        R, S, St = R.cpu().numpy(), S.cpu().numpy(), St.cpu().numpy()
        with np.errstate(all="ignore"):
            fin = np.isfinite(R) & np.isfinite(S) & (S > 0)
            votes = (fin & (np.abs(R) > TAU * S)).sum(axis=1)
            n_plain_rejects += int(((fin.sum(axis=1) >= 8) & (votes >= 0.5 * fin.sum(axis=1))).sum())
            finm = np.isfinite(R) & np.isfinite(St) & (St > 0)
            assert not (finm & (np.abs(R) > TAU * St)).any()          # no point of an exact solution votes
    assert n_plain_rejects >= 1, n_plain_rejects           # the cases are hard: the rule without majorants flips some of them
    print(f"ill-conditioned true solutions: {len(disguised)} x {len(grids)} grids survive; the plain-scale rule would reject {n_plain_rejects}")


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
    strs = UNIVARIATE[problem]
    pb, sess, prog, pts, pts_t, tab_t = _setup(problem, 64, cuda_device)
    dev, flags = _device_eval(pb, sess, prog, strs, pts_t, tab_t, cuda_device)
    assert not flags.any()
    jets = dev["jets"]
    oracle = _oracle_eval(problem, strs, pts)
    order = 4 if problem == "force_free" else 2
    n = _compare_points(dev, oracle, strs, order)
    assert n > 0
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


def test_integer_power_at_a_zero_of_its_base(cuda_device):
    """x ** n, n = 3, 4, ...: Taylor coefficients C(n, j) x^(n-j) by products only, so the jet stays finite where the
    base vanishes (round 1 divided by x: `(2*z - 1)**3` was NaN at the reference's third test point, z = 1/2)."""
    strs = ["(2*z - 1)**3", "rho*(2*z - 1)**4", "(rho - 7/8)**3 + z", "(z - 1/2)**5*exp(rho)", "(2*z - 1)**3/rho + (rho - 7/8)**4"]
    pb, sess, prog, pts, pts_t, tab_t = _setup("force_free", 64, cuda_device)
    assert tuple(pts[:, 2]) == (7 / 8, 1 / 2)
    dev, flags = _device_eval(pb, sess, prog, strs, pts_t, tab_t, cuda_device)
    assert not flags.any()
    assert np.isfinite(dev["jets"][:, :, 2]).all() and np.isfinite(dev["R"][:, 2]).all()
    oracle = _oracle_eval("force_free", strs, pts)
    for i, o in enumerate(oracle):
        np.testing.assert_allclose(dev["jets"][i, :, 2], o[0][:, 2], rtol=1e-12, atol=1e-13)
    _compare_points(dev, oracle, strs, 4)


def _random_expr(rng, depth, vars_, in_exp=False):
    """Random expression string over the reference's vocabulary after normalisation: + - * /,
    rational powers, sqrt, exp, Abs, small rational constants.  No exp inside an exp: doubly exponential values overflow
    to inf / underflow to 0 at most points, which tests nothing."""
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
    dev, flags = _device_eval(pb, sess, prog, strs, pts_t, tab_t, cuda_device)
    oracle = _oracle_eval(problem, strs, pts)
    # product and oracle agree on what is compilable
    assert [o is None for o in oracle] == [bool(f) for f in flags]
    # no exclusions: strings that cancel to a constant (`Abs(sqrt(z**2)) - z`, `x/(x + x)`) or depend on one coordinate only up
    # to round-off (`exp(-z/2) - rho*(z/rho)`) are covered by the majorant bounds like everything else
    n = _compare_points(dev, oracle, strs, 4 if problem == "force_free" else 2,
                        fin_agree=0.9)   # overflow corners (`z/exp(exp(z/rho))`: 1/inf vs inf*0) may differ at a few points
    assert n > 64 * 100

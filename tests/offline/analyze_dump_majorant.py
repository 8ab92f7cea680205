#!/usr/bin/env python3
"""Offline: does the majorant theory hold between the two float64 evaluations we have (device, numpy oracle)?
   jets:  |c_dev - c_orc| <= 2 eps W / t0^|g|          residual: |R_dev - R_orc| <= 2 tau S~
   and device (V, D, W, S~) vs the oracle's restatement."""
import gzip, json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from oracle import jets as J, parser as op, residuals as Rz, synth as osyn, majorant as Mj
from pde_engine_b200.grids import collocation_grid

TAU, T0 = 1e-10, Mj.T0_DEFAULT

def load(name):
    return json.load(gzip.open(os.path.join(REPO, "tests", "golden", name), "rt"))

def cmp(name, progs, d, key, problem, order, prims=None, prim_maj=None):
    sess = op.Session.for_problem(problem)
    pts = np.ascontiguousarray(collocation_grid(problem, 64).T)
    jets, R, St, maj = d[key + "_jets"], d[key + "_R"], d[key + "_St"], d[key + "_maj"]
    mi = J.multi_indices(order)
    deg = np.array([i + j for i, j in mi])
    worstJ = worstR = 0.0; nJ = nR = 0; badJ = []; badR = []; majdiff = []
    for i, s in enumerate(progs):
        if isinstance(s, str):
            c = op.compile_expr(s, sess)
            if c.flags: continue
            prog = c.whole()
        else:
            prog = s
        with np.errstate(all="ignore"):
            u, V, D, W = Mj.evaluate(prog, pts, order, sess.const_vals, sess.pow_vals, prims or (), prim_maj or (), t0=T0)
            if problem == "force_free":
                oR, oS, _ = Rz.force_free_residual(u, pts[:, 0]); oSt = Mj.force_free_scale(u, pts[:, 0], W, TAU, T0)
            else:
                oR, oS, _ = Rz.kerr_residual(u, pts); oSt = Mj.kerr_scale(u, pts, W, TAU, T0)
            ok = np.isfinite(u).all(axis=0) & np.isfinite(jets[i]).all(axis=0) & np.isfinite(W)
            bound = 2 * Mj.EPS * W[None, :] / (T0 ** deg)[:, None]
            ej = np.abs(jets[i] - u) / bound
            ej = np.where(bound > 0, ej, np.where(np.abs(jets[i] - u) == 0, 0.0, np.inf))
            e = np.max(ej, axis=0)
            for k in np.flatnonzero(ok & (e > 1)):
                badJ.append((e[k], s if isinstance(s, str) else i, k))
            nJ += int(ok.sum()); worstJ = max(worstJ, np.max(e[ok]) if ok.any() else 0)
            okr = ok & np.isfinite(oR) & np.isfinite(R[i]) & np.isfinite(oSt) & (oSt > 0)
            er = np.abs(R[i] - oR) / (2 * TAU * oSt)
            for k in np.flatnonzero(okr & (er > 1)):
                badR.append((er[k], s if isinstance(s, str) else i, k))
            nR += int(okr.sum()); worstR = max(worstR, np.max(er[okr]) if okr.any() else 0)
            # device majorants vs oracle
            fin = np.isfinite(W) & np.isfinite(maj[i, 2]) & (W > 0)
            if fin.any():
                majdiff.append(np.max(np.abs(maj[i, 2, fin] / W[fin] - 1)))
            fs = np.isfinite(oSt) & np.isfinite(St[i]) & (oSt > 0)
            if fs.any():
                majdiff.append(np.max(np.abs(St[i][fs] / oSt[fs] - 1)))
    print(f"== {name}: jets {nJ} points, worst |dc|/(2 eps W/t0^n) = {worstJ:.3g}, violations {len(badJ)}; residual {nR} points, worst |dR|/(2 tau S~) = {worstR:.3g}, violations {len(badR)}; device-vs-oracle (W, S~) max rel diff {max(majdiff):.3g}")
    for t in sorted(badJ, reverse=True)[:6]: print("   J", t)
    for t in sorted(badR, reverse=True)[:6]: print("   R", t)

def main():
    tag = sys.argv[1]
    d = np.load(os.path.join(REPO, "gpurun_out", f"dev_dump_{tag}.npz"))
    for problem, short, order, step in (("force_free", "ff", 4, 3), ("kerr_magnetosphere", "kerr", 2, 10)):
        e = load("enum_force_free_d4.json.gz" if short == "ff" else "enum_kerr_magnetosphere_d3.json.gz")
        E = {int(k): e["depths"][k]["uniques"] for k in e["depths"]}
        cmp(f"{problem} e123", E[1] + E[2] + E[3][::step], d, f"{short}_e123", problem, order)
        g = load(f"resid_{problem}.json.gz")
        cmp(f"{problem} golden", [r["s"] for r in g["records"]], d, f"{short}_golden", problem, order)
    sess = op.Session.for_problem("force_free")
    pts = np.ascontiguousarray(collocation_grid("force_free", 64).T)
    pm = [Mj.evaluate(op.compile_expr(s, sess).whole(), pts, 4, sess.const_vals, sess.pow_vals, t0=T0) for s in osyn.PRIM_EXPRS]
    cmp("synth", osyn.trees(osyn.SEED_TREES, 0, 2000, 5), d, "synth", "force_free", 4, [m[0] for m in pm], [(m[2], m[3]) for m in pm])

if __name__ == "__main__":
    main()

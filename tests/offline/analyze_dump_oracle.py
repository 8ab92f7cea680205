#!/usr/bin/env python3
"""Offline: device dump (e123 + synth sets) vs the float64 oracle; list points beyond 1e-10*S / 1e-10*mag."""
import gzip, json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from oracle import jets as J, parser as op, residuals as Rz, synth as osyn
from pde_engine_b200.grids import collocation_grid

def load(name):
    return json.load(gzip.open(os.path.join(REPO, "tests", "golden", name), "rt"))

SFLOOR = float(os.environ.get('SFLOOR', '1e-40'))

def cmp(name, strs_or_progs, jets, R, S, problem, order, prims=None):
    sess = op.Session.for_problem(problem)
    pts = np.ascontiguousarray(collocation_grid(problem, 64).T)
    nR = nJ = 0
    badR, badJ = [], []
    for i, s in enumerate(strs_or_progs):
        if isinstance(s, str):
            c = op.compile_expr(s, sess)
            if c.flags: continue
            prog = c.whole()
        else:
            prog = s
        with np.errstate(all="ignore"):
            u = J.evaluate(prog, pts, order, sess.const_vals, sess.pow_vals, prims) if prims is not None else J.evaluate(prog, pts, order, sess.const_vals, sess.pow_vals)
            oR, oS, _ = (Rz.force_free_residual(u, pts[:, 0]) if problem == "force_free" else Rz.kerr_residual(u, pts))
        ok = np.isfinite(u).all(axis=0) & np.isfinite(jets[i]).all(axis=0)
        if not ok.any(): continue
        mag = np.max(np.abs(u), axis=0)
        ej = np.max(np.abs(jets[i] - u), axis=0) / np.where(mag > 0, mag, 1)
        for k in np.flatnonzero(ok & (ej > 1e-10)):
            badJ.append((ej[k], i, k))
        nJ += int(ok.sum())
        okr = ok & np.isfinite(oR) & np.isfinite(oS) & np.isfinite(R[i]) & np.isfinite(S[i]) & (oS > 0)
        magf = np.where(ok, mag, 0.0)
        okr &= oS > SFLOOR * np.maximum(magf, 1.0) ** 6
        er = np.abs(R[i] - oR) / np.where(oS > 0, oS, 1)
        for k in np.flatnonzero(okr & (er > 1e-10)):
            badR.append((er[k], i, k, oS[k], mag[k]))
        nR += int(okr.sum())
    badR.sort(reverse=True); badJ.sort(reverse=True)
    print(f"== {name}: jets {nJ} points, {len(badJ)} beyond 1e-10*mag; residuals {nR} points, {len(badR)} beyond 1e-10*S")
    for t in badJ[:8]:
        print("   J %.3e  %s @%d" % (t[0], strs_or_progs[t[1]] if isinstance(strs_or_progs[t[1]], str) else t[1], t[2]))
    for t in badR[:8]:
        print("   R %.3e  %s @%d  S=%.3e mag=%.3e" % (t[0], strs_or_progs[t[1]] if isinstance(strs_or_progs[t[1]], str) else t[1], t[2], t[3], t[4]))
    return badJ, badR

def main():
    tag = sys.argv[1]
    d = np.load(os.path.join(REPO, "gpurun_out", f"dev_dump_{tag}.npz"))
    for problem, short, order, step in (("force_free", "ff", 4, 3), ("kerr_magnetosphere", "kerr", 2, 10)):
        e = load("enum_force_free_d4.json.gz" if short == "ff" else "enum_kerr_magnetosphere_d3.json.gz")
        E = {int(k): e["depths"][k]["uniques"] for k in e["depths"]}
        strs = E[1] + E[2] + E[3][::step]
        cmp(f"{problem} e123", strs, d[f"{short}_e123_jets"], d[f"{short}_e123_R"], d[f"{short}_e123_S"], problem, order)
    sess = op.Session.for_problem("force_free")
    pts = np.ascontiguousarray(collocation_grid("force_free", 64).T)
    oprim = [J.evaluate(op.compile_expr(s, sess).whole(), pts, 4, sess.const_vals, sess.pow_vals) for s in osyn.PRIM_EXPRS]
    progs = osyn.trees(osyn.SEED_TREES, 0, 2000, 5)
    cmp("synth", progs, d["synth_jets"], d["synth_R"], d["synth_S"], "force_free", 4, oprim)

if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Build container, offline: compare a device dump (tools/dump_device.py, brought back in gpurun_out/)
with the reference-derived golden vectors and with the oracle; print the worst ratios."""
import gzip, json, math, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from oracle import jets as J, parser as op, residuals as Rz

def load(name):
    return json.load(gzip.open(os.path.join(REPO, "tests", "golden", name), "rt"))

def fact_vec(order):
    f = np.zeros(J.ncoef(order))
    for n in range(order + 1):
        for j in range(n + 1):
            f[J.idx(n - j, j)] = math.factorial(n - j) * math.factorial(j)
    return f

def main():
    tag = sys.argv[1]
    d = np.load(os.path.join(REPO, "gpurun_out", f"dev_dump_{tag}.npz"))
    for problem, short, order in (("force_free", "ff", 4), ("kerr_magnetosphere", "kerr", 2)):
        g = load(f"resid_{problem}.json.gz")
        jets, R, S = d[f"{short}_golden_jets"], d[f"{short}_golden_R"], d[f"{short}_golden_S"]
        fv = fact_vec(order)
        sess = op.Session.for_problem(problem)
        pts = np.array(g["points"])
        worstR, worstJ = [], []
        by_order = {n: [] for n in range(order + 1)}
        orc_by_order = {n: [] for n in range(order + 1)}
        for i, rec in enumerate(g["records"]):
            c = op.compile_expr(rec["s"], sess)
            u = J.evaluate(c.whole(), pts, order, sess.const_vals, sess.pow_vals)
            od = J.derivatives(u, order)
            if problem == "force_free":
                oR, oS, _ = Rz.force_free_residual(u, pts[:, 0])
            else:
                oR, oS, _ = Rz.kerr_residual(u, pts)
            for k in range(8):
                gj = rec["jets"][k]
                if any(v is None for v in gj):
                    continue
                gj = np.array(gj)
                dj = jets[i, :, k] * fv
                if not np.isfinite(dj).all():
                    continue
                mag = np.max(np.abs(gj))
                if mag == 0: continue
                for n in range(order + 1):
                    sl = slice(n * (n + 1) // 2, (n + 1) * (n + 2) // 2)
                    e = np.max(np.abs(dj[sl] - gj[sl])) / mag
                    eo = np.max(np.abs(od[sl, k] - gj[sl])) / mag
                    by_order[n].append(e); orc_by_order[n].append(eo)
                    if n >= 2: worstJ.append((e, eo, rec["s"], k, n))
                gR = rec["R"][k]
                if gR is None or not np.isfinite(R[i, k]) or not (S[i, k] > 0):
                    continue
                worstR.append((abs(R[i, k] - gR) / S[i, k], abs(oR[k] - gR) / oS[k] if oS[k] > 0 else float('nan'), rec["s"], k, S[i, k] / oS[k] if oS[k] > 0 else float('nan')))
        worstR.sort(key=lambda t: -t[0]); worstJ.sort(key=lambda t: -t[0])
        print(f"== {problem}: {len(worstR)} residual points; |dR|/S device: max {worstR[0][0]:.3e}; > 1e-10: {sum(1 for t in worstR if t[0] > 1e-10)}; > 1e-11: {sum(1 for t in worstR if t[0] > 1e-11)}; > 1e-12: {sum(1 for t in worstR if t[0] > 1e-12)}")
        print("   oracle max", max(t[1] for t in worstR if t[1] == t[1]))
        for t in worstR[:12]:
            print("   R dev %.3e orc %.3e  S_dev/S_orc %.4f  %s @%d" % (t[0], t[1], t[4], t[2], t[3]))
        for n in range(order + 1):
            a = np.array(by_order[n]); b = np.array(orc_by_order[n])
            print(f"   jets order {n}: device max {a.max():.3e} (>1e-10: {(a > 1e-10).sum()}, >1e-12: {(a>1e-12).sum()}) oracle max {b.max():.3e} (>1e-10: {(b > 1e-10).sum()}) of {len(a)}")
        for t in worstJ[:12]:
            print("   J dev %.3e orc %.3e  %s @%d order %d" % t)

if __name__ == "__main__":
    main()

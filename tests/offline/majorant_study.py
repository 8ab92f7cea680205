#!/usr/bin/env python3
"""Offline prototype study of the majorant decision rule on reference verdicts (CPU, numpy oracle)."""
import json, os, sys, gzip
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from oracle import jets as J, parser as op, residuals as Rz, majorant as Mj
from pde_engine_b200.grids import collocation_grid

P = int(sys.argv[1]) if len(sys.argv) > 1 else 256
tau = 1e-10
t0 = float(sys.argv[2]) if len(sys.argv) > 2 else 0.125
sess = op.Session.for_problem("force_free")
pts = np.ascontiguousarray(collocation_grid("force_free", P).T)

def study(s):
    c = op.compile_expr(s, sess)
    if c.flags: return None
    u, V, D, W = Mj.evaluate(c.whole(), pts, 4, sess.const_vals, sess.pow_vals, t0=t0)
    R, S, _ = Rz.force_free_residual(u, pts[:, 0])
    St = Mj.force_free_scale(u, pts[:, 0], W, tau, t0)
    with np.errstate(all="ignore"):
        fin_old = np.isfinite(R) & np.isfinite(S) & (S > 0)
        fin_new = np.isfinite(R) & np.isfinite(St) & (St > 0)
        v_old = (fin_old & (np.abs(R) > tau * S)).sum()
        v_new = (fin_new & (np.abs(R) > tau * St)).sum()
        infl = np.nanmedian(np.where(fin_new & fin_old, St / S, np.nan)) if (fin_new & fin_old).any() else np.nan
        rmax = np.nanmax(np.where(fin_new, np.abs(R) / St, np.nan)) if fin_new.any() else np.nan
    return dict(nf_old=int(fin_old.sum()), nf_new=int(fin_new.sum()), v_old=int(v_old), v_new=int(v_new), infl=infl, rmax=rmax)

rows = []
fx = json.load(open(os.path.join(REPO, "tests/golden/ref_fixtures.json")))
for r in fx["ff_run_db"]:
    if r["id"] <= 85 and r["reason"] != "constant-only (skipped)":
        rows.append((r["expression"], bool(r["is_valid"])))
for f in ("verdicts_force_free_d2.json", "verdicts_force_free_d3.json"):
    for r in json.load(open(os.path.join(REPO, "tests/golden", f)))["records"]:
        if "is_valid" in r and not r.get("reason", "").startswith("Error") and not r.get("reason","").startswith("constant-only"):
            rows.append((r["s"], bool(r["is_valid"])))
log = "/tmp/pde_ref_work/verdicts_force_free_d3.jsonl"
if os.path.exists(log):
    for line in open(log):
        r = json.loads(line)
        if "is_valid" in r and not r.get("reason", "").startswith("Error") and not r.get("reason","").startswith("constant-only"):
            rows.append((r["s"], bool(r["is_valid"])))
rows += [("z*inv(z)/rho", True), ("rho*inv(rho/z)", True)]
seen = set(); uniq = []
for s, v in rows:
    if s not in seen:
        seen.add(s); uniq.append((s, v))
print(len(uniq), "rows;", sum(v for _, v in uniq), "valid")
def rej(nf, v): return nf >= 8 and v > 0 and v >= 0.5 * nf
stats = dict(valid=0, valid_votes_old=0, valid_votes_new=0, valid_rej_old=0, valid_rej_new=0, invalid=0, inv_rej_old=0, inv_rej_new=0)
infl = []
for s, v in uniq:
    r = study(s)
    if r is None: continue
    if v:
        stats["valid"] += 1
        stats["valid_votes_old"] += r["v_old"] > 0
        stats["valid_votes_new"] += r["v_new"] > 0
        if rej(r["nf_old"], r["v_old"]): stats["valid_rej_old"] += 1; print("  FALSE REJECT old:", s, r)
        if rej(r["nf_new"], r["v_new"]): stats["valid_rej_new"] += 1; print("  FALSE REJECT new:", s, r)
        if r["v_new"] > 0: print("  valid with votes (new):", s, r)
    else:
        stats["invalid"] += 1
        stats["inv_rej_old"] += rej(r["nf_old"], r["v_old"])
        stats["inv_rej_new"] += rej(r["nf_new"], r["v_new"])
        if rej(r["nf_old"], r["v_old"]) and not rej(r["nf_new"], r["v_new"]):
            print("  lost reject:", s, r)
    if r["infl"] == r["infl"]: infl.append(r["infl"])
print(stats)
print("inflation S~/S median of medians %.3g, p90 %.3g, max %.3g" % (np.median(infl), np.percentile(infl, 90), np.max(infl)))

#!/usr/bin/env python3
"""Offline: turn the per-candidate |R|/S~ statistics of tools/filter_margin.py (gpurun_out/margin_*.npz) into the
committed evidence file of the filter's decision rule -- profiles/filter_margin_<problem>_d<depth>.json:

  * histogram of the per-candidate MEDIAN of |R|/S~ over the finite points (decades) and of the VOTE FRACTION
    (share of finite points with |R| > tau S~, tau = 1e-10), for rejected candidates and for survivors;
  * the same two histograms on the round-1 rule's plain scale S (no round-off majorant), for comparison;
  * the BORDERLINE list: candidates with a vote fraction in [0.3, 0.7] or a median ratio in [1e-13, 1e-7] (the
    strings, their statistics and the kernel's verdict) -- what a change of tau, of the vote fraction or of the
    kernel's arithmetic could flip;
  * every candidate with a reference verdict (tests/golden/verdicts_*.json, the unmodified reference validator run
    offline): reference-valid rows must have NO voting point at all under the majorant rule (sound by construction:
    for an exact solution |R| <= tau S~ in any float64 evaluation order), and the report lists the largest vote
    fraction among them.

usage: python tools/filter_margin_report.py force_free 4 [t0]
"""
import gzip
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hist_decades(x, lo=-20, hi=2):
    x = np.asarray(x, float)
    out = {"nan_or_no_finite_point": int(np.isnan(x).sum()), "exactly_zero": int((x == 0).sum())}
    with np.errstate(divide="ignore"):
        lg = np.log10(x[(x > 0) & np.isfinite(x)])
    edges = np.arange(lo, hi + 1)
    h, _ = np.histogram(np.clip(lg, lo, hi - 1e-9), bins=edges)
    out["decades"] = {f"1e{int(a)}..1e{int(a) + 1}": int(c) for a, c in zip(edges[:-1], h) if c}
    return out


def hist_frac(x):
    x = np.asarray(x, float)
    edges = np.linspace(0, 1, 11)
    h, _ = np.histogram(np.clip(x, 0, 1 - 1e-12), bins=edges)
    return {"exactly_0": int((x == 0).sum()), "exactly_1": int((x == 1).sum()),
            "bins": {f"{a:.1f}..{b:.1f}": int(c) for a, b, c in zip(edges[:-1], edges[1:], h)}}


def main():
    problem = sys.argv[1] if len(sys.argv) > 1 else "force_free"
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    t0 = sys.argv[3] if len(sys.argv) > 3 else "0.0625"
    d = np.load(os.path.join(REPO, "gpurun_out", f"margin_{problem}_d{depth}_t{t0}.npz"))
    f = "enum_force_free_d4.json.gz" if problem == "force_free" else "enum_kerr_magnetosphere_d3.json.gz"
    strs = json.load(gzip.open(os.path.join(REPO, "tests", "golden", f), "rt"))["depths"][str(depth)]["uniques"]
    n = len(strs)
    qs = list(d["qs"])
    med = d["quant"][:, qs.index(0.5)]
    taus = list(d["taus"])
    vf = d["frac"][:, taus.index(1e-10)]
    surv = d["kernel_survivor"].astype(bool)
    nf = d["n_finite"]
    rep = {"problem": problem, "depth": depth, "n": n, "points": 4096, "tau": float(d["tau"]), "t0": float(d["t0"]),
           "rule": "reject iff n_finite >= 8 and votes >= 0.5 n_finite, a point votes iff |R| > tau S~ (S~: decision scale with the "
                   "round-off majorant, include/pde_b200.h); two-pass mode: the rejection must also stand on the confirmation sub-grid",
           "kernel": {"survivors": int(surv.sum()), "rejected": int((~surv).sum()), "not_evaluated": int((d["kernel_n_finite"] < 0).sum())},
           "median_ratio": {"rejected": hist_decades(med[~surv]), "survivors": hist_decades(med[surv])},
           "vote_fraction": {"rejected": hist_frac(vf[~surv]), "survivors": hist_frac(vf[surv & (nf > 0)])},
           "vote_fraction_other_tau": {f"{t:g}": hist_frac(d["frac"][:, k][nf > 0]) for k, t in enumerate(taus)},
           "plain_scale_S_round1_rule": {"vote_fraction": hist_frac(d["frac_sharp"][d["n_finite_sharp"] > 0]),
                                         "would_reject": int(((d["n_finite_sharp"] >= 8) & (d["frac_sharp"] >= 0.5)).sum())}}
    border = ((vf >= 0.3) & (vf <= 0.7) & (nf >= 8)) | ((med >= 1e-13) & (med <= 1e-7))
    rep["borderline"] = {"criterion": "vote fraction in [0.3, 0.7] (n_finite >= 8) or median |R|/S~ in [1e-13, 1e-7]",
                         "count": int(border.sum()), "rejected_among_them": int((border & ~surv).sum()),
                         "candidates": [dict(i=int(i), s=strs[i], n_finite=int(nf[i]), vote_fraction=round(float(vf[i]), 4),
                                             median_ratio=float(med[i]), q05=float(d["quant"][i, 0]), q95=float(d["quant"][i, -1]),
                                             survivor=bool(surv[i])) for i in np.flatnonzero(border)[:400]]}
    vpath = os.path.join(REPO, "tests", "golden", f"verdicts_{problem}_d{depth}.json")
    if os.path.exists(vpath):
        recs = json.load(open(vpath))["records"]
        idx = {s: i for i, s in enumerate(strs)}
        val = [idx[r["s"]] for r in recs if r.get("is_valid") and r["s"] in idx]
        inv = [idx[r["s"]] for r in recs if r.get("is_valid") is False and r["s"] in idx and r.get("reason") != "constant-only (skipped)"]
        val, inv = np.array(val, int), np.array(inv, int)
        rep["reference_verdicts"] = {
            "source": os.path.relpath(vpath, REPO), "valid": len(val), "invalid": len(inv),
            "valid_rejected_by_kernel": int((~surv[val]).sum()) if len(val) else 0,
            "valid_max_vote_fraction": float(vf[val].max()) if len(val) else None,
            "valid_max_median_ratio": float(np.nanmax(med[val])) if len(val) else None,
            "valid_vote_fraction_on_plain_scale_S_max": float(d["frac_sharp"][val].max()) if len(val) else None,
            # what the round-1 rule (votes on the plain scale S, no round-off majorant) would have done to them
            "valid_rejected_by_round1_rule": [strs[i] for i in val if d["n_finite_sharp"][i] >= 8 and d["frac_sharp"][i] >= 0.5],
            "timeouts_of_the_reference": sum(1 for r in recs if r.get("timeout")),
            "timeouts_rejected_by_kernel": int((~surv[[idx[r["s"]] for r in recs if r.get("timeout") and r["s"] in idx]]).sum()) if any(r.get("timeout") for r in recs) else 0,
            "invalid_rejected_by_kernel": int((~surv[inv]).sum()) if len(inv) else 0,
            "invalid_median_ratio": hist_decades(med[inv]) if len(inv) else None}
    out = os.path.join(REPO, "profiles", f"filter_margin_{problem}_d{depth}.json")
    json.dump(rep, open(out, "w"), indent=1)
    print("wrote", out, "borderline", rep["borderline"]["count"], rep.get("reference_verdicts", {}).get("valid_max_vote_fraction"))


if __name__ == "__main__":
    main()

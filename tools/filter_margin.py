#!/usr/bin/env python3
"""GPU box: per-candidate distribution of |R|/S over the 4096-point grid for a whole unique set
(default: the 143 461 force-free depth-4 uniques) -- the evidence behind the filter's decision rule
(SURVEY 7 "log the ratio histogram at depth 4/5").

Per candidate: n_finite, quantiles (5, 10, 25, 50, 75, 90, 95 %) of |R|/S~ over the finite points (S~ = the
decision scale with the round-off majorant, include/pde_b200.h), the fraction of finite points with
|R| > tau * S~ for tau = 1e-12, 1e-10, 1e-8, 1e-6, the same vote fraction on the plain scale S (the round-1
rule), and the product kernel's own outputs (n_votes, survivor).  Written to gpurun_out/margin_<problem>_d<depth>.npz; the
histograms / borderline lists are made offline by tools/filter_margin_report.py.
"""
from __future__ import annotations

import gzip
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import pde_engine_b200 as pb  # noqa: E402
from pde_engine_b200.grids import collocation_grid  # noqa: E402
from pde_engine_b200.validator import GpuBatchValidator  # noqa: E402

QS = (0.05, 0.10, 0.25, 0.50, 0.75, 0.90, 0.95)
TAUS = (1e-12, 1e-10, 1e-8, 1e-6)


def main():
    problem = sys.argv[1] if len(sys.argv) > 1 else "force_free"
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    P = 4096
    dev = torch.device("cuda", 0)
    f = "enum_force_free_d4.json.gz" if problem == "force_free" else "enum_kerr_magnetosphere_d3.json.gz"
    g = json.load(gzip.open(os.path.join(REPO, "tests", "golden", f), "rt"))
    strs = g["depths"][str(depth)]["uniques"]
    n = len(strs)
    t0 = float(sys.argv[3]) if len(sys.argv) > 3 else pb.core.T0_DEFAULT
    gv = GpuBatchValidator(None, problem, P=P, t0=t0)
    bv = gv.prefilter(strs)
    sess, prog = gv.session, gv.program
    quant = np.full((n, len(QS)), np.nan)
    frac = np.zeros((n, len(TAUS)))
    nfin = np.zeros(n, np.int32)
    nfin_sharp = np.zeros(n, np.int32)       # round-1 rule, for comparison: finite points / vote fraction on the plain scale S
    frac_sharp = np.zeros(n)
    q_t = torch.tensor(QS, dtype=torch.float64, device=dev)
    CH = 8192
    for lo in range(0, n, CH):
        hi = min(lo + CH, n)
        es = sess.compile(strs[lo:hi])
        code, ln = es.programs(128)
        _, R, S0, S, _ = pb.eval_points(sess, prog, torch.from_numpy(code).to(dev), torch.from_numpy(ln).to(dev), gv.pts, gv.table, None,
                                        spill_slots=8, want_jets=False, want_maj=True, tau=gv.tau, t0=gv.t0)
        aR = R.abs()
        fin0 = torch.isfinite(aR) & torch.isfinite(S0) & (S0 > 0)
        nfin_sharp[lo:hi] = fin0.sum(dim=1).cpu().numpy()
        frac_sharp[lo:hi] = ((fin0 & (aR > 1e-10 * S0)).sum(dim=1).double() / fin0.sum(dim=1).clamp(min=1).double()).cpu().numpy()
        fin = torch.isfinite(aR) & torch.isfinite(S) & (S > 0)
        ratio = torch.where(fin, aR / S, torch.full_like(aR, float("nan")))
        nf = fin.sum(dim=1)
        nfin[lo:hi] = nf.cpu().numpy()
        has = nf > 0
        if has.any():
            qv = torch.nanquantile(ratio[has], q_t, dim=1).T        # [m, len(QS)]
            tmp = np.full((hi - lo, len(QS)), np.nan)
            tmp[has.cpu().numpy()] = qv.cpu().numpy()
            quant[lo:hi] = tmp
        for k, tau in enumerate(TAUS):
            v = (fin & (aR > tau * S)).sum(dim=1).double() / nf.clamp(min=1).double()
            frac[lo:hi, k] = v.cpu().numpy()
        del R, S, S0, aR, fin, fin0, ratio
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    path = os.path.join(REPO, "gpurun_out", f"margin_{problem}_d{depth}_t{t0:g}.npz")
    np.savez_compressed(path, quant=quant, frac=frac, n_finite=nfin, qs=np.array(QS), taus=np.array(TAUS),
                        n_finite_sharp=nfin_sharp, frac_sharp=frac_sharp, t0=gv.t0, tau=gv.tau,
                        kernel_n_finite=bv.n_finite, kernel_n_votes=bv.n_votes, kernel_survivor=bv.survivor,
                        kernel_ratio_max=bv.ratio_max)
    print("wrote", path, "n", n, "survivors", int(bv.survivor.sum()))


if __name__ == "__main__":
    main()

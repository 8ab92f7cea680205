#!/usr/bin/env python3
"""GPU box: dump the device's per-point jets / residuals / scales for the parity sets into
gpurun_out/ so they can be compared OFFLINE (build container: oracle + exact SymPy arithmetic) --
the GPU box has neither the reference nor time for 50-digit arithmetic.

  gpurun_out/dev_dump_<tag>.npz : per set  <set>_jets [n, NC, 64], <set>_R, <set>_S [n, 64], <set>_strs
  sets: golden (tests/golden/resid_*.json.gz records), e123 (E1 + E2 + E3[::step]), fuzz seeds, synth
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


def load(name):
    return json.load(gzip.open(os.path.join(REPO, "tests", "golden", name), "rt"))


def dump(problem, strs, P=64):
    dev = torch.device("cuda", 0)
    sess = pb.Session.for_problem(problem)
    prog = pb.ResidualProgram.for_problem(problem)
    pts = collocation_grid(problem, P)
    pts_t = torch.from_numpy(pts).to(dev)
    tab_t = torch.from_numpy(prog.point_table(pts)).to(dev)
    es = sess.compile(strs)
    code, ln = es.programs(128)
    jets, R, S, St, maj = pb.eval_points(sess, prog, torch.from_numpy(code).to(dev), torch.from_numpy(ln).to(dev), pts_t, tab_t, None,
                                         spill_slots=8, want_maj=True)
    torch.cuda.synchronize()
    return jets.cpu().numpy(), R.cpu().numpy(), S.cpu().numpy(), ln, St.cpu().numpy(), maj.cpu().numpy()


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "cur"
    out = {}
    for problem, short in (("force_free", "ff"), ("kerr_magnetosphere", "kerr")):
        g = load(f"resid_{problem}.json.gz")
        strs = [r["s"] for r in g["records"]]
        j, R, S, ln, St, mj = dump(problem, strs)
        out[f"{short}_golden_jets"], out[f"{short}_golden_R"], out[f"{short}_golden_S"], out[f"{short}_golden_len"] = j, R, S, ln
        out[f"{short}_golden_St"], out[f"{short}_golden_maj"] = St, mj
        e = load("enum_force_free_d4.json.gz" if short == "ff" else "enum_kerr_magnetosphere_d3.json.gz")
        E = {int(d): e["depths"][d]["uniques"] for d in e["depths"]}
        strs = E[1] + E[2] + E[3][::(3 if short == "ff" else 10)]
        j, R, S, ln, St, mj = dump(problem, strs)
        out[f"{short}_e123_jets"], out[f"{short}_e123_R"], out[f"{short}_e123_S"], out[f"{short}_e123_len"] = j, R, S, ln
        out[f"{short}_e123_St"], out[f"{short}_e123_maj"] = St, mj
    # synthetic trees (PRIM leaves)
    from pde_engine_b200.synthetic import primitive_jets, SEED_TREES
    dev = torch.device("cuda", 0)
    sess = pb.Session.for_problem("force_free")
    prog = pb.ResidualProgram.for_problem("force_free")
    pts = collocation_grid("force_free", 64)
    pts_t = torch.from_numpy(pts).to(dev)
    tab_t = torch.from_numpy(prog.point_table(pts)).to(dev)
    prim_t = primitive_jets(sess, prog, pts_t, tab_t)
    d = pb.synth_trees(SEED_TREES, 0, 2000, 5, 48)
    j, R, S, St, mj = pb.eval_points(sess, prog, d["code"], d["len"], pts_t, tab_t, prim_t, spill_slots=2, want_maj=True)
    torch.cuda.synchronize()
    out["synth_jets"], out["synth_R"], out["synth_S"] = j.cpu().numpy(), R.cpu().numpy(), S.cpu().numpy()
    out["synth_St"], out["synth_maj"] = St.cpu().numpy(), mj.cpu().numpy()
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(REPO, "gpurun_out", f"dev_dump_{tag}.npz"), **out)
    print("wrote dev_dump", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

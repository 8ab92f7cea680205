#!/usr/bin/env python3
"""A/B: does the ORDER of the candidates matter to validate_kernel?  The same synthetic depth-5 batch (a) as generated
(random order), (b) sorted by (length, leading program bytes) so that the warps resident on an SM run near-identical
micro-op streams.  python tools/sorted_order_ab.py [n_trees]"""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import pde_engine_b200 as pb
from pde_engine_b200.grids import collocation_grid
from pde_engine_b200.synthetic import SEED_TREES, primitive_jets

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda", 0)
sess = pb.Session.for_problem("force_free")
prog = pb.ResidualProgram.for_problem("force_free")
pts = collocation_grid("force_free", 4096)
pts_t = torch.from_numpy(pts).to(dev)
tab_t = torch.from_numpy(prog.point_table(pts)).to(dev)
prim_t = primitive_jets(sess, prog, pts_t, tab_t)
trees = pb.synth_trees(SEED_TREES, 0, n, 5, 48, device=dev)
code, ln = trees["code"], trees["len"]


def run(code, ln, tag):
    out = None
    best = 1e9
    for rep in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = pb.validate(sess, prog, code, ln, pts_t, tab_t, prim_t, tau=1e-10, min_finite=8, vote_frac=0.5, n_ref=3, spill_slots=2, out=out)
        b.record()
        torch.cuda.synchronize()
        if rep:
            best = min(best, a.elapsed_time(b))
    print(f"{tag}: {best:.2f} ms", flush=True)
    return out


o1 = run(code, ln, "as generated")
# keys: the leading 8 program bytes, big-endian (lexicographic), then sorted stably by length
def be64(lo, hi):
    w = code[:, lo:hi].to(torch.int64)
    k = torch.zeros(n, dtype=torch.int64, device=dev)
    for j in range(hi - lo):
        k = (k << 8) | w[:, j]
    return k
# the op SIGNATURE: leaves collapsed (coordinates -> 1, PRIM -> 8, constants -> 0x80): what decides the micro-op kinds
raw = code
sig = code.clone()
sig[(code == 2)] = 1
sig[(code >= 0x08) & (code < 0x10)] = 8
sig[(code >= 0x80)] = 0x80


def be64sig(lo, hi):
    w = sig[:, lo:hi].to(torch.int64)
    k = torch.zeros(n, dtype=torch.int64, device=dev)
    for j in range(hi - lo):
        k = (k << 8) | w[:, j]
    return k


t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, keys in (("sorted by op signature, bytes 0-27", [be64sig(21, 28), be64sig(14, 21), be64sig(7, 14), be64sig(0, 7)]),
                   ("sorted by length only", [ln.to(torch.int64)]),
                   ("sorted by op signature 0-41, then length", [ln.to(torch.int64), be64sig(35, 42), be64sig(28, 35), be64sig(21, 28), be64sig(14, 21), be64sig(7, 14), be64sig(0, 7)][::-1][::-1]),("sorted by leading 7 bytes", [be64(0, 7)]),
                   ("sorted by bytes 0-13", [be64(7, 14), be64(0, 7)]),
                   ("sorted by length, then bytes 0-13", [be64(7, 14), be64(0, 7), ln.to(torch.int64)]),
                   ("sorted by bytes 0-27", [be64(21, 28), be64(14, 21), be64(7, 14), be64(0, 7)])):
    t0.record()
    order = torch.arange(n, device=dev)
    for k in keys:                      # least significant key first, stable sorts
        order = order[torch.sort(k[order], stable=True).indices]
    t1.record()
    torch.cuda.synchronize()
    sort_ms = t0.elapsed_time(t1)
    c2, l2 = code[order].contiguous(), ln[order].contiguous()
    o2 = run(c2, l2, f"{name} (sort {sort_ms:.2f} ms)")
    # same verdicts, permuted
    b1 = o1["n_votes"][order]
    assert torch.equal(b1, o2["n_votes"]) and torch.equal(o1["n_finite"][order], o2["n_finite"])

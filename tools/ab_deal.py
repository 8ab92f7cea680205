#!/usr/bin/env python3
"""A/B of the chunk dealing of validate_kernel (PDE_B200_STATIC_DEAL=1: round-robin) at 10^6 and 10^7 synthetic trees."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pde_engine_b200 as pb
from pde_engine_b200.grids import collocation_grid
from pde_engine_b200.synthetic import SEED_TREES, primitive_jets
dev = torch.device("cuda", 0)
sess = pb.Session.for_problem("force_free"); prog = pb.ResidualProgram.for_problem("force_free")
pts = collocation_grid("force_free", 4096); pts_t = torch.from_numpy(pts).to(dev); tab_t = torch.from_numpy(prog.point_table(pts)).to(dev)
prim_t = primitive_jets(sess, prog, pts_t, tab_t)
for n in (1000000, 10000000):
    trees = pb.synth_trees(SEED_TREES, 0, n, 5, 48, device=dev)
    out = None
    for _ in range(2):
        out = pb.validate(sess, prog, trees["code"], trees["len"], pts_t, tab_t, prim_t, confirm_points=128, n_ref=3, spill_slots=2, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        out = pb.validate(sess, prog, trees["code"], trees["len"], pts_t, tab_t, prim_t, confirm_points=128, n_ref=3, spill_slots=2, out=out)
    b.record(); torch.cuda.synchronize()
    print(os.environ.get("MODE", "?"), n, round(a.elapsed_time(b) / 3, 2), "ms per step", flush=True)
    del trees, out

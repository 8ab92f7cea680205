#!/usr/bin/env python3
"""Stage 2 on the real depth-4 unique set (143 461 force-free strings): kernel time per configuration.
Development / profiling driver:  python tools/depth4_kernel.py [spill_slots] [reps]"""
import gzip, json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import pde_engine_b200 as pb
from pde_engine_b200.grids import collocation_grid

ss = int(sys.argv[1]) if len(sys.argv) > 1 else 3
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with gzip.open(os.path.join(REPO, "tests", "golden", "enum_force_free_d4.json.gz"), "rt") as f:
    strs = json.load(f)["depths"]["4"]["uniques"]
dev = torch.device("cuda", 0)
sess = pb.Session.for_problem("force_free")
prog = pb.ResidualProgram.for_problem("force_free")
pts = collocation_grid("force_free", 4096)
pts_t = torch.from_numpy(pts).to(dev)
tab_t = torch.from_numpy(prog.point_table(pts)).to(dev)
es = sess.compile(strs)
code, ln = es.programs(128)
c, l = torch.from_numpy(code).to(dev), torch.from_numpy(ln).to(dev)
out = None
for r in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = pb.validate(sess, prog, c, l, pts_t, tab_t, None, spill_slots=ss, out=out)
    b.record()
    torch.cuda.synchronize()
    print(f"spill_slots {ss} rep {r}: {a.elapsed_time(b):.2f} ms", flush=True)
nf = out["n_finite"].cpu().numpy()
print("overflowed (needs more spill slots):", int((nf == -3).sum()), "not evaluable:", int((nf < 0).sum()))
# one rank's shard at N = 8 (the first 1/8 of the strings): how close to 1/8 of the full time?  (kernel tail)
n8 = (len(strs) // 8) // 32 * 32
o8 = None
best = 1e9
for r in range(reps + 2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    o8 = pb.validate(sess, prog, c[:n8], l[:n8], pts_t, tab_t, None, spill_slots=ss, out=o8)
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print(f"shard 1/8 ({n8} candidates): {best:.2f} ms", flush=True)

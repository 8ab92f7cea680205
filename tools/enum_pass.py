#!/usr/bin/env python3
"""Stage 1 at depth 5 from the real depth 1-4 force-free unique sets (11.78 M candidates): pass time.
Development / profiling driver:  python tools/enum_pass.py [L] [reps]"""
import gzip, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import pde_engine_b200 as pb

L = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
with gzip.open(os.path.join(REPO, "tests", "golden", "enum_force_free_d4.json.gz"), "rt") as f:
    gd = json.load(f)["depths"]
flat, db = [], [0]
for d in ("1", "2", "3", "4"):
    flat += gd[d]["uniques"]
    db.append(len(flat))
sess = pb.Session.for_problem("force_free")
es = sess.compile(flat)
n5 = pb.enumerate_count(es, db, 5, True)
cand = pb.enumerate_candidates(es, db, 5, True, 0, n5, L)
torch.cuda.synchronize()
for r in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    cand = pb.enumerate_candidates(es, db, 5, True, 0, n5, L)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"L {L} rep {r}: {ms:.3f} ms  {n5} candidates  {n5 * (L + 21) / ms / 1e6:.0f} GB/s (incl. output allocation)", flush=True)
print("nonempty rows:", int((cand["len"] > 0).sum()), "mean len:", float(cand["len"].float().mean()))

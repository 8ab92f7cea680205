#!/usr/bin/env python3
"""Stage 2 on the real Kerr depth <= 3 unique set (16 482 strings, order-2 jets): kernel time.
Development driver:  python tools/kerr_kernel.py [reps]"""
import gzip, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch
import pde_engine_b200 as pb
from pde_engine_b200.grids import collocation_grid

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
with gzip.open(os.path.join(REPO, "tests", "golden", "enum_kerr_magnetosphere_d3.json.gz"), "rt") as f:
    gk = json.load(f)["depths"]
strs = [s for d in sorted(gk, key=int) for s in gk[d]["uniques"]] * 8          # 8 copies: a longer kernel to time
dev = torch.device("cuda", 0)
sess = pb.Session.for_problem("kerr_magnetosphere")
prog = pb.ResidualProgram.for_problem("kerr_magnetosphere")
pts = collocation_grid("kerr_magnetosphere", 4096)
pts_t = torch.from_numpy(pts).to(dev)
tab_t = torch.from_numpy(prog.point_table(pts)).to(dev)
es = sess.compile(strs)
code, ln = es.programs(128)
c, l = torch.from_numpy(code).to(dev), torch.from_numpy(ln).to(dev)
out = None
for r in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = pb.validate(sess, prog, c, l, pts_t, tab_t, None, spill_slots=2, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"rep {r}: {ms:.2f} ms  {len(strs) * 4096 / ms / 1e6:.1f} G evals/s", flush=True)
import numpy as np
print("survivors:", int(np.unpackbits(out["survivor_bits"].cpu().numpy().view(np.uint8)).sum()), "n_finite sum:", int(out["n_finite"].clamp(min=0).sum()))

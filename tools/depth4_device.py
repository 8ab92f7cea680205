#!/usr/bin/env python3
"""Depth 4 without strings: GpuBatchValidator.filter_enumerated on the 258 285 raw force-free depth-4 candidates
(stage 1 -> stage 2 on the device), and the cost of each of W contiguous windows of the index space (what the ranks
of a sharded run would each do).  Development / profiling driver:  python tools/depth4_device.py [W]"""
import gzip, json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import pde_engine_b200 as pb
from pde_engine_b200.distributed import shard_range
from pde_engine_b200.validator import GpuBatchValidator

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
with gzip.open(os.path.join(REPO, "tests", "golden", "enum_force_free_d4.json.gz"), "rt") as f:
    gd = json.load(f)["depths"]
flat, db = [], [0]
for d in ("1", "2", "3"):
    flat += gd[d]["uniques"]
    db.append(len(flat))
gv = GpuBatchValidator(None, "force_free", P=4096, group=None)
for rep in range(4):
    t0 = time.perf_counter()
    surv = gv.filter_enumerated(flat, db, 4, True, 128)
    print(f"filter_enumerated rep {rep}: {1e3 * (time.perf_counter() - t0):.2f} ms, n = {len(surv)}, rejected {int((~surv).sum())}", flush=True)
es = gv.session.compile(flat)
n = pb.enumerate_count(es, db, 4, True)
csr = pb.enumerate_candidates_csr(es, db, 4, True, 0, n, 128)
first, nu = pb.dedup_csr(csr["pool"], csr["off"], csr["len"], csr["hash"])
ln = csr["len"].cpu().numpy().astype(np.int64)
f = first.cpu().numpy().astype(bool)
print(f"n {n}, distinct programs {nu}, not spliced {int((ln == 0).sum())}, survivors among first occurrences {int((surv & f).sum())}")
for lo, cnt in [shard_range(n, r, W) for r in range(W)]:
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gv._enum_filter_local(es, db, 4, True, 128, lo, cnt)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    sl = slice(lo, lo + cnt)
    print(f"window [{lo}, {lo + cnt}): {1e3 * best:.2f} ms; sum of program bytes {int(ln[sl].sum())}, first occurrences {int(f[sl].sum())}", flush=True)

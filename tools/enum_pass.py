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
# one eighth of the index space (what one of 8 ranks enumerates): should cost ~1/8 of the full pass
import ctypes as C
from pde_engine_b200 import _lib as _l
dbc = (C.c_int32 * len(db))(*db)
for first, count, tag in ((0, n5, "full"), (3 * (n5 // 8), n5 // 8, "window 3/8..4/8")):
    out = {k: v[:max(count, 1)] for k, v in cand.items()}
    best = 1e9
    for r in range(reps + 2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _l.check(_l.lib.pde_enumerate(es._h, dbc, 5, 1, first, count, L, C.c_void_p(out["triple"].data_ptr()),
                                      C.c_void_p(out["code"].data_ptr()), C.c_void_p(out["len"].data_ptr()),
                                      C.c_void_p(out["hash"].data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"{tag}: {best:.3f} ms for {count} candidates (count pass cached on the handle)", flush=True)
# CSR form: programs packed in one byte pool (16-byte aligned)
csr = pb.enumerate_candidates_csr(es, db, 5, True, 0, n5, L)
torch.cuda.synchronize()
pool_bytes = csr["pool"].numel()
best = 1e9
for r in range(reps + 2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _l.check(_l.lib.pde_enumerate_csr(es._h, dbc, 5, 1, 0, n5, L, C.c_void_p(csr["triple"].data_ptr()), C.c_void_p(csr["off"].data_ptr()),
                                      C.c_void_p(csr["pool"].data_ptr()), C.c_void_p(csr["len"].data_ptr()),
                                      C.c_void_p(csr["hash"].data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    b.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
real = pool_bytes + n5 * (4 + 1 + 8 + 12)
print(f"CSR full: {best:.3f} ms; pool {pool_bytes / n5:.1f} B + 25 B per candidate = {real / n5:.1f} B; "
      f"{real / best / 1e6:.0f} GB/s on real bytes, {n5 * 69 / best / 1e6:.0f} GB/s on SURVEY's 69 B", flush=True)

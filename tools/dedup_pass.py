import gzip, json, os, sys, time
sys.path.insert(0, "/root/repo")
import torch
import pde_engine_b200 as pb
with gzip.open("/root/repo/tests/golden/enum_force_free_d4.json.gz", "rt") as f:
    gd = json.load(f)["depths"]
flat, db = [], [0]
for d in ("1", "2", "3", "4"):
    flat += gd[d]["uniques"]; db.append(len(flat))
sess = pb.Session.for_problem("force_free")
es = sess.compile(flat)
n5 = pb.enumerate_count(es, db, 5, True)
cand = pb.enumerate_candidates(es, db, 5, True, 0, n5, 128)
for r in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    first, nu = pb.dedup(cand["code"], cand["len"], cand["hash"])
    print(f"dedup call {r}: {(time.perf_counter() - t0) * 1e3:.2f} ms  unique {nu}")

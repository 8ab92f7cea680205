import gzip, json, os, sys, time, gc
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import pde_engine_b200 as pb
from pde_engine_b200 import core
from pde_engine_b200.validator import GpuBatchValidator
REPO="/root/repo"
with gzip.open(os.path.join(REPO, "tests", "golden", "enum_force_free_d4.json.gz"), "rt") as f:
    gd = json.load(f)["depths"]
flat, db = [], [0]
for d in ("1", "2", "3"):
    flat += gd[d]["uniques"]; db.append(len(flat))
gv = GpuBatchValidator(None, "force_free", P=4096, group=None)
P = time.perf_counter
for rep in range(60):
    t = [P()]
    exprs = gv.session.compile(flat); t.append(P())
    n = core.enumerate_count(exprs, db, 4, True); t.append(P())
    bits = gv._enum_filter_local(exprs, db, 4, True, 128, 0, n); t.append(P())
    words = bits.cpu().numpy().view(np.uint32); t.append(P())
    k = np.arange(n)
    s = ((words[k >> 5] >> (k & 31).astype(np.uint32)) & 1).astype(bool); t.append(P())
    del exprs, bits; t.append(P())
    tot = 1e3 * (t[-1] - t[0])
    if rep < 3 or tot > 40:
        print(f"rep {rep}: {tot:.2f} ms: compile {1e3*(t[1]-t[0]):.2f} count {1e3*(t[2]-t[1]):.2f} local {1e3*(t[3]-t[2]):.2f} d2h {1e3*(t[4]-t[3]):.2f} unpack {1e3*(t[5]-t[4]):.2f} free {1e3*(t[6]-t[5]):.2f}", flush=True)
print("done")

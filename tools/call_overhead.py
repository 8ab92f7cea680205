import sys, time
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import pde_engine_b200 as pb
from pde_engine_b200.grids import collocation_grid
dev = torch.device("cuda", 0)
sess = pb.Session.for_problem("force_free"); prog = pb.ResidualProgram.for_problem("force_free")
pts = collocation_grid("force_free", 4096)
pts_t = torch.from_numpy(pts).to(dev); tab_t = torch.from_numpy(prog.point_table(pts)).to(dev)
es = sess.compile(["rho**2*z", "rho*z", "exp(rho/z)", "rho/z + 1"] * 500)
code, ln = es.programs(128)
c, l = torch.from_numpy(code).to(dev), torch.from_numpy(ln).to(dev)
for n in (4, 2000):
    out = pb.validate(sess, prog, c[:n], l[:n], pts_t, tab_t, None, spill_slots=2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        out = pb.validate(sess, prog, c[:n], l[:n], pts_t, tab_t, None, spill_slots=2, out=out)
    torch.cuda.synchronize()
    print(f"n={n}: {(time.perf_counter() - t0) / 50 * 1e6:.0f} us per pde_validate call")
from pde_engine_b200.validator import GpuBatchValidator
gv = GpuBatchValidator(None, "force_free", P=4096)
strs = ["rho**2*z", "rho*z", "exp(rho/z)", "rho/z + 1"] * 500
gv.prefilter(strs); t0 = time.perf_counter()
for _ in range(10): gv.prefilter(strs)
print(f"prefilter(2000 strings): {(time.perf_counter() - t0) / 10 * 1e3:.2f} ms")

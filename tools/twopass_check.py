#!/usr/bin/env python3
"""GPU box: the two-pass filter (proposals on all points, confirmation with majorants on a sub-grid) against the
one-pass majorant rule on the whole grid: survivor counts, disagreements, kernel times (CUDA events)."""
import gzip, json, os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import pde_engine_b200 as pb
from pde_engine_b200.grids import collocation_grid


def bits_to_bool(bits, n):
    b = bits.cpu().numpy().view(np.uint32)
    k = np.arange(n)
    return ((b[k >> 5] >> (k & 31).astype(np.uint32)) & 1).astype(bool)


def main():
    dev = torch.device("cuda", 0)
    for problem, f, depth in (("force_free", "enum_force_free_d4.json.gz", 4), ("force_free", "enum_force_free_d4.json.gz", 3),
                              ("kerr_magnetosphere", "enum_kerr_magnetosphere_d3.json.gz", 3)):
        g = json.load(gzip.open(os.path.join(REPO, "tests", "golden", f), "rt"))
        strs = g["depths"][str(depth)]["uniques"]
        sess = pb.Session.for_problem(problem)
        prog = pb.ResidualProgram.for_problem(problem)
        pts = collocation_grid(problem, 4096)
        pts_t = torch.from_numpy(pts).to(dev)
        tab_t = torch.from_numpy(prog.point_table(pts)).to(dev)
        code, ln = sess.compile(strs).programs(128)
        code_t, len_t = torch.from_numpy(code).to(dev), torch.from_numpy(ln).to(dev)
        res = {}
        for cp in (0, 128, 256, 512):
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                o = pb.validate(sess, prog, code_t, len_t, pts_t, tab_t, None, confirm_points=cp, spill_slots=2)
                e1.record()
                torch.cuda.synchronize()
            res[cp] = (bits_to_bool(o["survivor_bits"], len(strs)), e0.elapsed_time(e1))
        base = res[0][0]
        print(f"{problem} d{depth}: n {len(strs)}; one-pass majorant survivors {int(base.sum())} in {res[0][1]:.2f} ms")
        for cp in (128, 256, 512):
            sv, ms = res[cp]
            print(f"   confirm_points {cp}: survivors {int(sv.sum())} in {ms:.2f} ms; survive here but rejected by the one-pass rule: "
                  f"{int((sv & ~base).sum())}; rejected here but survive the one-pass rule: {int((~sv & base).sum())}")


if __name__ == "__main__":
    main()

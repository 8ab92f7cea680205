#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric):
candidate x point FP64 jet evaluations per second on the synthetic depth-5 batch
(10^7 bytecode trees x 4096 collocation points, force-free residual).

    python bench.py --gpus N --steps K --warmup W            (ours)
    python bench.py --impl reference --gpus N ...            (CPU arm: the oracle port
                                                              of the path on all host cores)

A step = one pass of stage 2 (pde_validate) over the rank's resident batch of
trees; at N > 1 every rank owns its own 10^7 trees (weak scaling, no data-path
collective) and each step ends with the gather of survivor bitmasks + hashes
to rank 0.  Timed with CUDA events, barrier + synchronize on both sides, max
over ranks.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

METRIC = "candidate_x_point_fp64_jet_evals_per_sec"
UNIT = "evals/s"
NOMINAL_FP64_TFLOPS = 37.2      # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz
# validate_kernel's DRAM traffic from the committed ncu --set full capture (read + written bytes / trees of that launch)
TRAFFIC_BYTES_PER_TREE = 55.0
TRAFFIC_SOURCE = "profiles/r2_validate_full_metrics.csv (ncu --set full of pass 1, 200 000 trees: dram__bytes_read.sum 10.99 MB, dram__bytes_write.sum 0; pass 2 adds 8.5 + 2.6 MB)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trees", type=int, default=10_000_000, help="trees per GPU")
    ap.add_argument("--points", type=int, default=4096)
    ap.add_argument("--depth", type=int, default=5)
    ap.add_argument("--L", type=int, default=48)
    ap.add_argument("--cpu-sample", type=int, default=3600, help="trees in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ref-wall", type=float, default=45.0,
                    help="seconds given to the reference's own validator pool (cpu_baseline_reference; 0 = skip)")
    ap.add_argument("--confirm-points", type=int, default=128,
                    help="points of the confirmation pass with round-off majorants (0: one pass with majorants on all points)")
    return ap.parse_args()


# ------------------------------------------------------------------ CPU arm
def _cpu_setup(points):
    import numpy as np
    from oracle import jets as J, parser as op, residuals as Rz, synth as osyn
    pts = Rz.collocation_grid("force_free", points)
    osess = op.Session.for_problem("force_free")
    prim = [J.evaluate(op.compile_expr(s, osess).whole(), pts, 4, osess.const_vals, osess.pow_vals) for s in osyn.PRIM_EXPRS]
    return pts, osess, prim


_CPU = {}


def _cpu_worker(args):
    """Oracle port of the path: jets + residual + vote for trees [first, first+count)."""
    import numpy as np
    from oracle import jets as J, residuals as Rz, synth as osyn
    seed, first, count, depth, points = args
    if _CPU.get("points") != points:
        _CPU["points"] = points
        _CPU["ctx"] = _cpu_setup(points)
    pts, osess, prim = _CPU["ctx"]
    surv = 0
    for i in range(count):
        code = osyn.tree(seed, first + i, depth)
        u = J.evaluate(code, pts, 4, osess.const_vals, osess.pow_vals, prim)
        R, S, _ = Rz.force_free_residual(u, pts[:, 0])
        with np.errstate(all="ignore"):
            fin = np.isfinite(R) & np.isfinite(S) & (S > 0)
            votes = int((np.abs(R[fin]) > 1e-10 * S[fin]).sum())
        nf = int(fin.sum())
        surv += 0 if (nf >= 8 and votes > 0 and votes >= 0.5 * nf) else 1
    return surv


def cpu_baseline(sample_trees, points, depth, procs=1):
    from oracle import synth as osyn
    t0 = time.perf_counter()
    if procs <= 1:
        _cpu_worker((osyn.SEED_TREES, 0, sample_trees, depth, points))
    else:
        import multiprocessing as mp
        per = max(1, sample_trees // procs)
        with mp.get_context("fork").Pool(procs) as pool:
            pool.map(_cpu_worker, [(osyn.SEED_TREES, k * per, per, depth, points) for k in range(procs)])
        sample_trees = per * procs
    dt = time.perf_counter() - t0
    return sample_trees * points / dt, dt, sample_trees


def run_reference(args):
    """Reference arm: the CPU implementation of the path (oracle port -- the reference is
    pure Python/SymPy and cannot travel to the GPU box) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores * 300, 600)
    from oracle import synth as osyn
    import multiprocessing as mp
    per = max(1, per_step // cores)
    jobs = lambda k0: [(osyn.SEED_TREES, (k0 * cores + k) * per, per, args.depth, args.points) for k in range(cores)]
    with mp.get_context("fork").Pool(cores) as pool:
        for w in range(args.warmup):
            pool.map(_cpu_worker, jobs(w))
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(_cpu_worker, jobs(args.warmup + s))
        dt = time.perf_counter() - t0
    n = per * cores
    value = n * args.points * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"synthetic depth-{args.depth} trees x {args.points} points, force-free residual",
                   "trees_per_step": n, "points": args.points, "note": "bounded sample of the 10^7-tree workload per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n} trees x {args.points} points per step, oracle (numpy) port of jets+residual+vote, {cores} processes"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if args.ref_wall > 0:
        line["cpu_baseline_reference"] = cpu_baseline_reference(args.ref_wall)
    _emit(line)


def cpu_baseline_reference(wall_s):
    """The reference's OWN validator pool (SymPy) on this box: `--validators os.cpu_count()` from a copy of
    baseline/_ref with the one-line worker import repair, bounded to wall_s seconds (tools/ref_cpu_baseline.py)."""
    try:
        from tools import ref_cpu_baseline as rb
        if not rb.available():
            return {"unavailable": "baseline/_ref (the reference's code, made by tools/refcopy.py) is not in this snapshot"}
        r = rb.run_reference_validators("force_free", 3, os.cpu_count() or 1, wall_s)
        if r["rows_validated"] == 0:          # the pool did not come up inside the window (seen once on a busy host): one more try
            first_try = {k: r[k] for k in ("validator_processes_started", "rows_inserted", "wall_s", "log_tail")}
            r = rb.run_reference_validators("force_free", 3, os.cpu_count() or 1, wall_s)
            r["first_try"] = first_try
        r["sample"] = (f"{r['rows_validated']} rows of the reference's own depth <= 3 force-free run validated by its "
                       f"{r['validators']} validator processes in {r['wall_s']} s (1 test point per row, FFV:296-297)")
        return r
    except Exception as e:
        return {"error": repr(e)[:300]}


# ------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, ln in self.lines:
            if t < t0 or t > t1 + 0.3:
                continue
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ ours
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import pde_engine_b200 as pb
    from pde_engine_b200 import flops as flopmod
    from pde_engine_b200.distributed import gather_survivors
    from pde_engine_b200.grids import collocation_grid
    from pde_engine_b200.synthetic import SEED_TREES, primitive_jets

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, P, L = args.trees, args.points, args.L

    sess = pb.Session.for_problem("force_free")
    prog = pb.ResidualProgram.for_problem("force_free")
    pts = collocation_grid("force_free", P)
    pts_t = torch.from_numpy(pts).to(dev)
    tab_t = torch.from_numpy(prog.point_table(pts)).to(dev)
    prim_t = primitive_jets(sess, prog, pts_t, tab_t)
    trees = pb.synth_trees(SEED_TREES, rank * n, n, args.depth, L, device=dev)   # stage-1 stand-in: resident in HBM
    flops_per_point = flopmod.batch_flops_per_point(trees["code"], "force_free", 4)
    out = None

    def step():
        nonlocal out
        out = pb.validate(sess, prog, trees["code"], trees["len"], pts_t, tab_t, prim_t,
                          tau=1e-10, min_finite=8, vote_frac=0.5, confirm_points=args.confirm_points, n_ref=3, spill_slots=2, out=out)
        if world > 1:
            gather_survivors(out["survivor_bits"], trees["hash"], n)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = pb.fp64_peak(20000)
    fp64_peak_3op = pb.fp64_peak_3op(20000)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = pb.launch_count()
    kern_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    e0.record()
    for s in range(args.steps):
        kern_ev[s][0].record()
        out = pb.validate(sess, prog, trees["code"], trees["len"], pts_t, tab_t, prim_t,
                          tau=1e-10, min_finite=8, vote_frac=0.5, confirm_points=args.confirm_points, n_ref=3, spill_slots=2, out=out)
        kern_ev[s][1].record()
        if world > 1:
            gather_survivors(out["survivor_bits"], trees["hash"], n)
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = pb.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    kms = torch.tensor([sum(a.elapsed_time(b) for a, b in kern_ev) / args.steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(ms.item()), float(kms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    value = world * n * P * args.steps / (total_ms * 1e-3)

    # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        code_h = trees["code"].cpu().pin_memory()
        len_h = trees["len"].cpu().pin_memory()
        bits_h = torch.empty((n + 31) // 32, dtype=torch.int32).pin_memory()
        nfin_h = torch.empty(n, dtype=torch.int32).pin_memory()
        ratio_h = torch.empty(n, dtype=torch.float64).pin_memory()
        code_d, len_d = torch.empty_like(trees["code"]), torch.empty_like(trees["len"])

        def e2e_step():
            code_d.copy_(code_h, non_blocking=True)
            len_d.copy_(len_h, non_blocking=True)
            o = pb.validate(sess, prog, code_d, len_d, pts_t, tab_t, prim_t, tau=1e-10, min_finite=8, vote_frac=0.5,
                            n_ref=3, spill_slots=2, out=out)
            bits_h.copy_(o["survivor_bits"], non_blocking=True)
            nfin_h.copy_(o["n_finite"], non_blocking=True)
            ratio_h.copy_(o["ratio_max"], non_blocking=True)
            if world > 1:
                gather_survivors(o["survivor_bits"], trees["hash"], n)

        e2e_step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            e2e_step()
        b.record()
        barrier()
        ems = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e = {"value": world * n * P * args.steps / (float(ems.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(code_h.numel() + len_h.numel()),
               "d2h_bytes_per_step": int(bits_h.numel() * 4 + nfin_h.numel() * 4 + ratio_h.numel() * 8)}

    # ---- strong scaling of the same metric: the headline batch size IN TOTAL, 1/N of it per GPU ----
    strong = None
    if world > 1:
        ns_ = (n // world) // 32 * 32
        code_s, len_s, hash_s = trees["code"][:ns_], trees["len"][:ns_], trees["hash"][:ns_]
        out_s = None

        def strong_step():
            nonlocal out_s
            out_s = pb.validate(sess, prog, code_s, len_s, pts_t, tab_t, prim_t, tau=1e-10, min_finite=8, vote_frac=0.5,
                                confirm_points=args.confirm_points, n_ref=3, spill_slots=2, out=out_s)
            gather_survivors(out_s["survivor_bits"], hash_s, ns_)

        strong_step()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            strong_step()
        b.record()
        barrier()
        sms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        strong = {"scaling": "strong", "total_trees": ns_ * world, "trees_per_gpu": ns_, "ms_per_step": float(sms.item()) / args.steps,
                  "value": world * ns_ * P * args.steps / (float(sms.item()) * 1e-3), "unit": UNIT,
                  "note": "same kernel, same grid, the headline batch size in total: every rank validates the first 1/N of its resident trees and the step ends with the same survivor gather"}

    # ---- second BASELINE metric: depth-4 validation wall time ----
    import gzip
    depth4 = None
    try:
        with gzip.open(os.path.join(REPO, "tests", "golden", "enum_force_free_d4.json.gz"), "rt") as f:
            uniq4 = json.load(f)["depths"]["4"]["uniques"]
        from pde_engine_b200.distributed import shard_range
        first4, cnt4 = shard_range(len(uniq4), rank, world)
        mine = uniq4[first4:first4 + cnt4]
        # untimed warm-up of the exact code path on a small slice (first use of this kernel
        # configuration loads its module: a one-off cost of the process, not of the validation)
        esw = sess.compile(mine[:256])
        cw, lw = esw.programs(128)
        pb.validate(sess, prog, torch.from_numpy(cw).to(dev), torch.from_numpy(lw).to(dev), pts_t, tab_t, None,
                    tau=1e-10, min_finite=8, vote_frac=0.5, n_ref=3, spill_slots=2)
        if world > 1:     # same for the gather: NCCL sets up its buffers for a message size on first use
            gather_survivors(torch.zeros((cnt4 + 31) // 32, dtype=torch.int32, device=dev),
                             torch.zeros(cnt4, dtype=torch.int64, device=dev), cnt4)
        barrier()
        # (1) the public API path: GpuBatchValidator.prefilter = host compile pipelined with the device (chunks).
        # At N > 1 it is driven exactly as the engine would drive it: rank 0 holds ALL the strings and calls
        # prefilter(), ranks > 0 sit in serve(); the strings travel as one byte blob, every rank compiles and
        # validates its contiguous shard, the verdict rows are gathered on rank 0 (the only exchange).
        from pde_engine_b200.validator import GpuBatchValidator
        gv4 = GpuBatchValidator(None, "force_free", P=P, device=dev)
        tp0 = tp1 = 0.0
        if rank == 0:
            gv4.prefilter(uniq4)          # warm-up at full size: the validator's pinned staging buffers are grown once and reused
            gv4.prefilter(uniq4)
            walls4 = []
            try:
                for _ in range(5):            # median of 5: a single call varies by +-5 ms with the host's scheduling of the compile threads
                    tp0 = time.perf_counter()
                    bv4 = gv4.prefilter(uniq4)
                    walls4.append(time.perf_counter() - tp0)     # rank 0 returns after the gather: the wall of the whole job
            finally:
                gv4.shutdown()
            tp0, tp1 = 0.0, float(np.median(walls4))
        else:
            gv4.serve()
        barrier()
        # (2) the same work step by step, for the breakdown
        t0 = time.perf_counter()
        es4 = sess.compile(mine)                                    # host compiler: strings -> bytecode
        code4, len4 = es4.programs(128)
        t1 = time.perf_counter()
        c4 = torch.from_numpy(code4).to(dev, non_blocking=True)
        l4 = torch.from_numpy(len4).to(dev, non_blocking=True)
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        o4 = pb.validate(sess, prog, c4, l4, pts_t, tab_t, None, tau=1e-10, min_finite=8, vote_frac=0.5, n_ref=3, spill_slots=2)
        k1.record()
        bits4 = o4["survivor_bits"].cpu()
        nf4 = o4["n_finite"].cpu()
        barrier()
        t2 = time.perf_counter()
        w = torch.tensor([(tp1 - tp0) * 1e3, (t1 - t0) * 1e3, k0.elapsed_time(k1), (t2 - t0) * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        nsurv = int(sum(bin(int(x) & 0xffffffff).count("1") for x in bits4.tolist()))
        tot = torch.tensor([nsurv, int((nf4 < 0).sum())], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        if rank == 0:
            assert int(tot[0]) == int(bv4.survivor.sum()), "the API path (sharded, pipelined) and the one-shot shards disagree"
        depth4 = {"input": "143461 force-free depth-4 unique strings (tests/golden/enum_force_free_d4.json.gz)",
                  "n": len(uniq4), "points": P, "wall_ms_host_strings_to_survivor_bits": float(w[0]),
                  "wall_ms_runs": [round(1e3 * x, 2) for x in walls4] if rank == 0 else None,
                  "unpipelined_wall_ms": float(w[3]), "host_compile_ms": float(w[1]), "kernel_ms": float(w[2]),
                  "survivors_for_cpu_confirmation": int(tot[0]), "not_device_evaluable": int(tot[1]),
                  "host_threads": min(os.cpu_count() or 1, 16),
                  "note": "span A = GpuBatchValidator.prefilter on rank 0 (the public batch entry, all strings on rank 0's host): at N > 1 the strings go to the ranks in serve() as one byte blob, every rank compiles (multi-threaded C++ parser, cores / N threads) and validates its contiguous shard (H2D, kernel, D2H of all per-candidate outputs), the verdict rows are gathered on rank 0; host_compile_ms / kernel_ms / unpipelined_wall_ms are the same work done step by step on each rank's shard, max over ranks"}
    except Exception as e:   # the fixture is optional for the headline metric
        depth4 = {"error": repr(e)[:200]}

    # ---- the same depth WITHOUT strings: stage 1 -> stage 2 on the device (the generator's last-depth path) ----
    # Rank 0 passes the 3 786 operand strings (depth <= 3 uniques) to GpuBatchValidator.filter_enumerated; every rank
    # enumerates its window of the 258 285 raw depth-4 candidates into CSR rows and validates them where they were
    # written; the survivor words are gathered on rank 0.  No candidate ever exists as a string.
    depth4_dev = None
    try:
        with gzip.open(os.path.join(REPO, "tests", "golden", "enum_force_free_d4.json.gz"), "rt") as f:
            gd4 = json.load(f)["depths"]
        flat4, db4 = [], [0]
        for d_ in ("1", "2", "3"):
            flat4 += gd4[d_]["uniques"]
            db4.append(len(flat4))
        gv5 = GpuBatchValidator(None, "force_free", P=P, device=dev)
        walls = []
        if rank == 0:
            try:
                for _ in range(2):
                    gv5.filter_enumerated(flat4, db4, 4, True, 128)
                for _ in range(5):
                    tq = time.perf_counter()
                    surv5 = gv5.filter_enumerated(flat4, db4, 4, True, 128)
                    walls.append((time.perf_counter() - tq) * 1e3)
            finally:
                gv5.shutdown()             # (always: the other ranks sit in serve())
            es5 = gv5.session.compile(flat4)
            c5 = pb.enumerate_candidates_csr(es5, db4, 4, True, 0, len(surv5), 128, device=dev)
            f5, nu5 = pb.dedup_csr(c5["pool"], c5["off"], c5["len"], c5["hash"])
            f5 = f5.cpu().numpy().astype(bool)
            depth4_dev = {"input": f"{len(flat4)} force-free uniques of depth <= 3 (operands); the {len(surv5)} raw depth-4 candidates are enumerated on the device",
                          "n_candidates": int(len(surv5)), "distinct_programs": int(nu5), "points": P,
                          "wall_ms_operand_strings_to_survivor_flags": float(np.median(walls)), "wall_ms_runs": [round(x, 3) for x in walls],
                          "survivors_for_the_normaliser": int((surv5 & f5).sum()), "rejected_on_device": int((~surv5).sum()),
                          "note": "GpuBatchValidator.filter_enumerated from rank 0 (ranks > 0 in serve()): operand strings broadcast (0.1 MB), every rank compiles them, enumerates its shard_range window in CSR form (pde_enumerate_csr), drops exact duplicates inside the window (pde_dedup_csr) and validates the rest (pde_validate_csr); one gather of survivor words.  Host wall clock on rank 0 around the call, median of 5"}
        else:
            gv5.serve()
    except Exception as e:
        depth4_dev = {"error": repr(e)[:200]}
    barrier()

    # ---- BASELINE configs[3]: kerr_magnetosphere depth 3 (order-2 jets, 16 174 uniques), same span A ----
    kerr3 = None
    if rank == 0:
        try:
            with gzip.open(os.path.join(REPO, "tests", "golden", "enum_kerr_magnetosphere_d3.json.gz"), "rt") as f:
                gk = json.load(f)["depths"]
            uk = [s_ for d in sorted(gk, key=int) for s_ in gk[d]["uniques"]]
            ksess = pb.Session.for_problem("kerr_magnetosphere")
            kprog = pb.ResidualProgram.for_problem("kerr_magnetosphere")
            kpts = collocation_grid("kerr_magnetosphere", P)
            kpts_t = torch.from_numpy(kpts).to(dev)
            ktab_t = torch.from_numpy(kprog.point_table(kpts)).to(dev)
            # untimed warm-up of the exact code path at full size (first use of a kernel configuration loads its module:
            # a one-off cost of the process -- a 256-string warm-up once left 50 ms of it inside the timed call)
            ew = ksess.compile(uk)
            cw, lw = ew.programs(128)
            pb.validate(ksess, kprog, torch.from_numpy(cw).to(dev), torch.from_numpy(lw).to(dev), kpts_t, ktab_t, None,
                        tau=1e-10, min_finite=8, vote_frac=0.5, n_ref=3, spill_slots=2)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ek = ksess.compile(uk)
            ck, lk = ek.programs(128)
            t1 = time.perf_counter()
            ckd, lkd = torch.from_numpy(ck).to(dev), torch.from_numpy(lk).to(dev)
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            ok_ = pb.validate(ksess, kprog, ckd, lkd, kpts_t, ktab_t, None, tau=1e-10, min_finite=8, vote_frac=0.5, n_ref=3, spill_slots=2)
            k1.record()
            kb = ok_["survivor_bits"].cpu()
            knf = ok_["n_finite"].cpu()
            t2 = time.perf_counter()
            kerr3 = {"input": f"{len(uk)} kerr_magnetosphere uniques of depth <= 3 (tests/golden/enum_kerr_magnetosphere_d3.json.gz)",
                     "n": len(uk), "points": P, "wall_ms_host_strings_to_survivor_bits": (t2 - t0) * 1e3,
                     "host_compile_ms": (t1 - t0) * 1e3, "kernel_ms": k0.elapsed_time(k1),
                     "survivors_for_cpu_confirmation": int(sum(bin(int(x) & 0xffffffff).count("1") for x in kb.tolist())),
                     "not_device_evaluable": int((knf < 0).sum())}
        except Exception as e:
            kerr3 = {"error": repr(e)[:200]}

    # ---- function fingerprints of the depth-4 uniques and of the survivors ----
    fp_info = None
    if rank == 0 and world == 1 and isinstance(depth4, dict) and "error" not in depth4:
        try:
            import numpy as np
            from pde_engine_b200.fingerprint import GpuFingerprinter
            fpr = GpuFingerprinter("force_free", P=64, device=dev)
            fpr.fingerprint(uniq4[:256])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f4 = fpr.fingerprint(uniq4)
            t1 = time.perf_counter()
            surv = np.asarray(bv4.survivor, bool)
            ks = f4.key[surv]
            fp_info = {"rows": len(uniq4), "points": 64, "mantissa_bits": 26, "wall_ms_host_strings_to_keys": (t1 - t0) * 1e3,
                       "functions": int(len(np.unique(f4.key[f4.key != 0]))), "unknown": int((f4.key == 0).sum()),
                       "survivors": int(surv.sum()), "survivor_functions": int(len(np.unique(ks[ks != 0]))),
                       "survivor_unknown": int((ks == 0).sum()),
                       "note": "opt-in (run_discovery share_confirmations): the CPU confirms one representative per survivor function"}
        except Exception as e:
            fp_info = {"error": repr(e)[:200]}

    # ---- stage 1 (HBM bound): depth-5 enumeration from the depth 1-4 unique sets ----
    enum_info = None
    if rank == 0 and world == 1 and isinstance(depth4, dict) and "error" not in depth4:
        try:
            with gzip.open(os.path.join(REPO, "tests", "golden", "enum_force_free_d4.json.gz"), "rt") as f:
                gd = json.load(f)["depths"]
            flat, db = [], [0]
            for d in ("1", "2", "3", "4"):
                flat += gd[d]["uniques"]
                db.append(len(flat))
            es5 = sess.compile(flat)
            Le = 128
            n5 = pb.enumerate_count(es5, db, 5, True)
            cand = pb.enumerate_candidates(es5, db, 5, True, 0, n5, Le)        # warm-up (allocates outputs)
            torch.cuda.synchronize()
            l0 = pb.launch_count()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            import ctypes as C
            from pde_engine_b200 import _lib as _l
            dbc = (C.c_int32 * len(db))(*db)
            a.record()
            for _ in range(3):
                _l.check(_l.lib.pde_enumerate(es5._h, dbc, 5, 1, 0, n5, Le, C.c_void_p(cand["triple"].data_ptr()),
                                              C.c_void_p(cand["code"].data_ptr()), C.c_void_p(cand["len"].data_ptr()),
                                              C.c_void_p(cand["hash"].data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            b.record()
            torch.cuda.synchronize()
            ems = a.elapsed_time(b) / 3
            # the same pass in CSR form (programs packed in one 16-byte-aligned byte pool)
            csr = pb.enumerate_candidates_csr(es5, db, 5, True, 0, n5, Le)
            torch.cuda.synchronize()
            ca, cb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ca.record()
            for _ in range(3):
                _l.check(_l.lib.pde_enumerate_csr(es5._h, dbc, 5, 1, 0, n5, Le, C.c_void_p(csr["triple"].data_ptr()),
                                                  C.c_void_p(csr["off"].data_ptr()), C.c_void_p(csr["pool"].data_ptr()),
                                                  C.c_void_p(csr["len"].data_ptr()), C.c_void_p(csr["hash"].data_ptr()),
                                                  C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            cb.record()
            torch.cuda.synchronize()
            csr_ms = ca.elapsed_time(cb) / 3
            csr_bytes = int(csr["pool"].numel()) + n5 * (4 + 1 + 8 + 12)
            del csr
            first, nuniq = pb.dedup(cand["code"], cand["len"], cand["hash"])          # warm-up
            torch.cuda.synchronize()
            td0 = time.perf_counter()
            first, nuniq = pb.dedup(cand["code"], cand["len"], cand["hash"])          # syncs internally (returns the count)
            dedup_ms = (time.perf_counter() - td0) * 1e3
            bytes_per = Le + 1 + 8 + 12
            hbm_peak = None
            try:
                hbm_peak = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
            except Exception:
                pass
            gbs = n5 * bytes_per / (ems * 1e-3) / 1e9
            # the enumerator only WRITES: pure-write HBM rate for context (fill of 2 GiB, best of 5)
            wbuf = torch.empty(2 << 30, dtype=torch.uint8, device=dev)
            wbuf.fill_(1)
            wbest = 0.0
            for _ in range(5):
                wa, wb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                wa.record(); wbuf.fill_(2); wb.record(); torch.cuda.synchronize()
                wbest = max(wbest, wbuf.numel() / (wa.elapsed_time(wb) * 1e-3) / 1e9)
            del wbuf
            enum_info = {"depth": 5, "n_candidates": int(n5), "distinct_programs": int(nuniq), "L": Le,
                         "ms_per_pass": ems, "launches_per_pass": int((pb.launch_count() - l0) // 3),
                         "algorithmic_bytes_per_candidate": bytes_per, "achieved_GBps": gbs,
                         "peak_GBps": hbm_peak if hbm_peak else 6650.0,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "fallback 6.65 TB/s (B200_PROFILING.md)",
                         "frac": gbs / (hbm_peak if hbm_peak else 6650.0),
                         "write_only_peak_GBps": wbest, "frac_of_write_only_peak": gbs / wbest,
                         "note": "the pass writes 149 B per candidate and reads ~0 (operands are L2 resident): the copy-rate peak counts read + write bytes, a write-only stream (torch fill) reaches write_only_peak_GBps on this GPU",
                         "candidates_per_s": n5 / (ems * 1e-3),
                         # SURVEY 8d's algorithmic figure: 69 B per candidate (L = 48 rows)
                         "frac_on_survey_69B": n5 * 69 / (ems * 1e-3) / 1e9 / (hbm_peak if hbm_peak else 6650.0),
                         "csr": {"ms_per_pass": csr_ms, "bytes_per_candidate": csr_bytes / n5,
                                 "achieved_GBps_real_bytes": csr_bytes / (csr_ms * 1e-3) / 1e9,
                                 "frac_of_write_only_peak_real_bytes": csr_bytes / (csr_ms * 1e-3) / 1e9 / wbest,
                                 "GBps_on_survey_69B": n5 * 69 / (csr_ms * 1e-3) / 1e9,
                                 "candidates_per_s": n5 / (csr_ms * 1e-3),
                                 "note": "pde_enumerate_csr: a third of the bytes at ~0.86 of the time -- the pass is instruction-issue bound (ncu: 77-79 % of the issue slots active, ~1 150 thread instructions per candidate: decode, splice, hash), not HBM bound"},
                         # exact-duplicate removal (reported separately): hash-table insert + byte-wise
                         # confirming lookup; wall time of the call incl. its table allocation and the count read-back
                         "dedup_ms": dedup_ms, "dedup_candidates_per_s": n5 / (dedup_ms * 1e-3),
                         "dedup_duplicates_dropped": int(n5 - nuniq)}
            del cand, first
        except Exception as e:
            enum_info = {"error": repr(e)[:200]}

    if rank == 0:
        achieved = flops_per_point * P / (kernel_ms * 1e-3) / 1e12
        nf = out["n_finite"]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"synthetic depth-{args.depth}: {n} bytecode trees per GPU x {P} collocation points, force-free residual (order-4 jets)",
                       "trees_per_gpu": n, "points": P, "L": L, "seed": hex(SEED_TREES),
                       "l2": (f"inputs ({n * (L + 1) / 1e6:.0f} MB of bytecode per GPU, streamed once per step) exceed the 126 MB L2; no flush needed"
                              if n * (L + 1) > 126e6 else
                              f"inputs are {n * (L + 1) / 1e6:.0f} MB (< L2): reduced --trees run, not the headline configuration"),
                       "parallelism": f"{world} x independent candidate shards, final gather only"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak,
                         # DRAM bytes per launch, from the ncu --set full capture of this kernel (not measured in this
                         # run: a profiler cannot run inside the bench): dram__bytes_read.sum + dram__bytes_write.sum per
                         # 200 k trees, scaled per tree
                         "traffic": TRAFFIC_BYTES_PER_TREE * n, "traffic_unit": "bytes per launch (HBM idle: kernel is FP64-pipe bound)",
                         "traffic_source": TRAFFIC_SOURCE,
                         "peak_source": "measured: pde_fp64_peak register-resident DFMA chains on this GPU (MEASURED_PEAKS.json has no FP64 entry)",
                         "frac_of_nominal": achieved / NOMINAL_FP64_TFLOPS, "nominal_peak": NOMINAL_FP64_TFLOPS,
                         # context: a DFMA that reads three different register pairs (acc += a_i * b_j, the operand
                         # pattern of every jet convolution) issues every 3 cycles, not 2 -- measured on this GPU
                         "peak_3operand_dfma": fp64_peak_3op, "frac_of_3operand_peak": achieved / fp64_peak_3op,
                         "flops_per_candidate_point": flops_per_point / n, "kernel_ms": kernel_ms,
                         "kernel": "validate_kernel<force_free, reduce>"},
            "survivor_fraction": float((out["survivor_bits"].view(torch.uint8).cpu().numpy().view("uint8")
                                        .reshape(-1, 1) >> np.arange(8) & 1).sum() / n),
            "evaluated_fraction": float((nf >= 0).float().mean().item()),
            "survivor_note": "about half of the random depth-5 trees 'survive': they are NaN on most of the grid or depend on one "
                             "coordinate only (every monomial of the determinant vanishes), i.e. undecidable for a numerical filter; "
                             "the kernel has no early exit, every tree is evaluated at every point (evaluated_fraction)",
            "strong_scaling": strong,
            "depth4_validation": depth4,
            "depth4_device_resident": depth4_dev,
            "kerr_depth3_validation": kerr3,
            "function_fingerprints_depth4": fp_info,
            "enumerator": enum_info,
        }
        if not args.no_cpu_baseline and world == 1:
            v, dt, ns = cpu_baseline(args.cpu_sample, P, args.depth, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"first {ns} trees of the same batch x {P} points, oracle (numpy) port, {dt:.1f} s"}
            if args.ref_wall > 0:
                line["cpu_baseline_reference"] = cpu_baseline_reference(args.ref_wall)
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def _emit(line: dict) -> None:
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    global _RESULT_FD
    args = parse_args()
    # Libraries chat on stdout (NCCL prints its version line there): everything but the result line goes to stderr.
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

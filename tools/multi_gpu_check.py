#!/usr/bin/env python3
"""Multi-GPU parity of the sharded hot path, run under torchrun with N >= 2 GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multi_gpu_check.py

Every rank enumerates + validates its contiguous window of the 258 285 force-free depth-4 candidates
(reference order), rank 0 gathers survivor bitmasks + hashes (the only exchange) and compares them with its own
single-GPU pass over the whole index space: bit-exact."""
import gzip, json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np
import torch
import torch.distributed as dist
import pde_engine_b200 as pb
from pde_engine_b200.distributed import enumerated_program_fetcher, gather_survivors, merge_survivors, shard_range
from pde_engine_b200.grids import collocation_grid

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
with gzip.open(os.path.join(REPO, "tests", "golden", "enum_force_free_d4.json.gz"), "rt") as f:
    gd = json.load(f)["depths"]
flat, db = [], [0]
for d in ("1", "2", "3"):
    flat += gd[d]["uniques"]
    db.append(len(flat))
sess = pb.Session.for_problem("force_free")
prog = pb.ResidualProgram.for_problem("force_free")
pts = collocation_grid("force_free", 4096)
pts_t = torch.from_numpy(pts).to(dev)
tab_t = torch.from_numpy(prog.point_table(pts)).to(dev)
es = sess.compile(flat)
n = pb.enumerate_count(es, db, 4, True)
assert n == 258285, n


def run(first, count):
    c = pb.enumerate_candidates(es, db, 4, True, first, count, 128, device=dev)
    o = pb.validate(sess, prog, c["code"], c["len"], pts_t, tab_t, None, spill_slots=2)
    return c, o


first, count = shard_range(n, rank, world)
cand, out = run(first, count)
g = gather_survivors(out["survivor_bits"], cand["hash"], count, lens=cand["len"])
if rank == 0:
    idx, hs = merge_survivors(g, fetch_programs=enumerated_program_fetcher(es, db, 4, True, 128))
    cand_all, out_all = run(0, n)
    bits = out_all["survivor_bits"].cpu().numpy().view(np.uint32)
    k = np.arange(n)
    surv = ((bits[k >> 5] >> (k & 31).astype(np.uint32)) & 1).astype(bool)
    h = cand_all["hash"].cpu().numpy()
    # single-GPU reference of merge_survivors: survivors in order, first occurrence of every PROGRAM (pde_dedup's
    # byte-confirmed first-occurrence flags restricted to the survivors; uncompiled candidates are never merged)
    code_all, len_all = cand_all["code"].cpu().numpy(), cand_all["len"].cpu().numpy()
    seen, want_idx = set(), []
    for i in np.nonzero(surv)[0]:
        key = bytes(code_all[i, :len_all[i]])
        if len_all[i] == 0 or key not in seen:
            seen.add(key)
            want_idx.append(int(i))
    assert idx == want_idx, (len(idx), len(want_idx))
    assert hs == [int(h[i]) for i in want_idx]
    # per-shard bitmasks concatenate to the single-GPU bitmask
    cat = np.concatenate([b.cpu().numpy().view(np.uint32) for b in g[0]])
    assert np.array_equal(cat[:len(bits)], bits)
    print(f"multi-GPU parity ok: {world} ranks, {n} depth-4 candidates, {int(surv.sum())} survivors, "
          f"{len(idx)} distinct surviving programs; sharded == single-GPU bit for bit", flush=True)
dist.barrier()

# the same through the public API: GpuBatchValidator.filter_enumerated driven from rank 0, ranks > 0 in serve()
import time
from pde_engine_b200.validator import GpuBatchValidator
gv = GpuBatchValidator(None, "force_free", P=4096, device=dev)
if rank == 0:
    try:
        for _ in range(2):
            got = gv.filter_enumerated(flat, db, 4, True, 128)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            got = gv.filter_enumerated(flat, db, 4, True, 128)
            ts.append(1e3 * (time.perf_counter() - t0))
    finally:
        gv.shutdown()
    alone = GpuBatchValidator(None, "force_free", P=4096, device=dev, group=None)
    want = alone.filter_enumerated(flat, db, 4, True, 128)
    t0 = time.perf_counter()
    want = alone.filter_enumerated(flat, db, 4, True, 128)
    t1 = 1e3 * (time.perf_counter() - t0)
    es2 = alone.session.compile(flat)
    c2 = pb.enumerate_candidates_csr(es2, db, 4, True, 0, n, 128, device=dev)
    f2 = pb.dedup_csr(c2["pool"], c2["off"], c2["len"], c2["hash"])[0].cpu().numpy().astype(bool)
    assert np.array_equal(got, want), int((got != want).sum())      # (every rank dedups the whole depth: same rows evaluated)
    print(f"filter_enumerated: {world} ranks == 1 rank on all {n} candidates ({int(f2.sum())} first occurrences evaluated); "
          f"wall {sorted(ts)[len(ts) // 2]:.2f} ms on {world} GPUs (runs {[round(x, 2) for x in ts]}) vs {t1:.2f} ms on one", flush=True)
else:
    gv.serve()
dist.barrier()
dist.destroy_process_group()

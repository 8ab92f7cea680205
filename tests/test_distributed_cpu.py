"""Host-side sharding logic with world_size 2 over gloo on the CPU."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO


def test_shard_ranges_tile_the_index_space():
    from pde_engine_b200.distributed import shard_range
    for n in (0, 1, 31, 32, 33, 1000, 258285, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(world):
                first, count = shard_range(n, r, world)
                assert first == pos and count >= 0
                assert first % 32 == 0 or first == n
                pos += count
            assert pos == n
    # balanced to within one bitmask word
    sizes = [shard_range(10_000_000, r, 8)[1] for r in range(8)]
    assert max(sizes) - min(sizes) <= 32


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pde_engine_b200.distributed import gather_survivors, merge_survivors, shard_range
    first, count = shard_range(n_total, rank, world)
    rng = np.random.default_rng(1234)
    surv_all = rng.random(n_total) < 0.3
    hash_all = rng.integers(1, 2 ** 62, size=n_total, dtype=np.int64)
    len_all = rng.integers(1, 40, size=n_total).astype(np.uint8)
    prog_all = [bytes(rng.integers(1, 255, size=int(k)).astype(np.uint8)) for k in len_all]
    dup, col, un = n_total - 5, n_total - 9, (7, n_total - 11, n_total - 13)
    hash_all[dup] = hash_all[3]; len_all[dup] = len_all[3]; prog_all[dup] = prog_all[3]       # a cross-shard exact duplicate
    hash_all[col] = hash_all[5]; len_all[col] = len_all[5]                                    # a 64-bit COLLISION: same hash and length, other bytes
    prog_all[col] = bytes((b ^ 1) or 2 for b in prog_all[5])
    for u in un:                                                                              # not compiled on the device: constant hash, len 0
        hash_all[u] = 0x1234; len_all[u] = 0; prog_all[u] = b""
    for i in (3, 5, dup, col) + un:
        surv_all[i] = True
    surv = surv_all[first:first + count]
    words = np.zeros((count + 31) // 32, dtype=np.uint32)
    for i in np.nonzero(surv)[0]:
        words[i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    g = gather_survivors(torch.from_numpy(words.view(np.int32)), torch.from_numpy(hash_all[first:first + count].copy()), count,
                         lens=torch.from_numpy(len_all[first:first + count].copy()))
    if rank == 0:
        idx, hs = merge_survivors(g, fetch_programs=lambda ids: [prog_all[i] for i in ids])
        want = [i for i in np.nonzero(surv_all)[0] if i != dup]   # only the byte-identical duplicate goes (lowest index stays)
        idx_nofetch, _ = merge_survivors(g)                      # without programs nothing is merged
        q.put((idx == [int(i) for i in want] and all(u in idx for u in un) and col in idx and
               idx_nofetch == [int(i) for i in np.nonzero(surv_all)[0]],
               [int(h) for h in hs] == [int(hash_all[i]) for i in want]))
    else:
        assert g is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_survivors_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1003, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok == (True, True)


def test_grids_match_oracle():
    from oracle.residuals import collocation_grid as ogrid
    from pde_engine_b200.grids import collocation_grid
    for prob, oname in (("force_free", "force_free"), ("kerr_magnetosphere", "kerr")):
        a = collocation_grid(prob, 256)
        b = ogrid(oname, 256)
        assert np.array_equal(a.T, b)
    with pytest.raises(ValueError):
        collocation_grid("nope", 64)


def _fake_local(self, strs, compile_threads=None, blob=None, n=None):
    """Stand-in for the device filter (no GPU in the CPU suite): a deterministic verdict row per string.  A worker
    rank gets its shard as the byte blob rank 0 broadcast (strs is None), rank 0 gets both."""
    from pde_engine_b200.validator import BatchVerdict
    if blob is not None:
        from_blob = blob.decode().split("\0")[:-1]
        assert strs is None or list(strs) == from_blob
        assert n is None or n == len(from_blob)
        strs = from_blob
    n = len(strs)
    h = np.array([sum(s.encode()) for s in strs], dtype=np.int64).reshape(n)
    surv = (h % 3) != 0
    bits = np.zeros((n + 31) // 32, np.uint32)
    for i in np.flatnonzero(surv):
        bits[i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    out = {"ratio_max": h * 0.5, "resid_max": h * 0.25, "scale_at": h * 2.0, "n_finite": (h % 97).astype(np.int32),
           "n_votes": (h % 13).astype(np.int32), "ref_rs": np.repeat(h.astype(np.float64), 6).reshape(n, 3, 2),
           "confirm": np.stack([h % 5, h % 7], axis=1).astype(np.int32), "survivor_bits": bits.view(np.int32)}
    return BatchVerdict(strs, (h % 2).astype(np.uint8), out)


def _serve_worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pde_engine_b200.validator import GpuBatchValidator
    GpuBatchValidator._prefilter_local = _fake_local
    gv = object.__new__(GpuBatchValidator)          # the protocol only: no device behind it
    gv.group, gv.device = "auto", torch.device("cpu")
    strs = [f"rho**{k} + z*{k % 17}" for k in range(5003)]
    if rank == 0:
        bv = gv.prefilter(strs)                      # dealt out to both ranks
        small = gv.prefilter(strs[:100])             # below SHARD_MIN: stays on rank 0, no collective
        gv.shutdown()
        want = _fake_local(gv, strs)
        same = all(np.array_equal(getattr(bv, k), getattr(want, k)) for k in
                   ("ratio_max", "resid_max", "scale_at", "n_finite", "n_votes", "ref_rs", "confirm", "survivor", "flags"))
        q.put((same, bv.strs == strs, len(small.strs) == 100))
    else:
        served = gv.serve()
        assert served == 1
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_prefilter_protocol_gloo_world2():
    """GpuBatchValidator.prefilter under torch.distributed: rank 0 deals contiguous shards, rank 1 sits in serve(), the
    gathered verdict rows equal the single-process result, shutdown() releases the worker."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_serve_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok == (True, True, True)


def _enum_worker(rank, world, port, q):
    sys.path.insert(0, REPO)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pde_engine_b200 import core
    from pde_engine_b200.validator import GpuBatchValidator
    strs = [f"rho**{k} + z" for k in range(1, 41)]
    depth_begin = [0, 5, 40]
    n_total = 70001

    rng = np.random.default_rng(7)
    len_all = torch.from_numpy(rng.integers(1, 60, size=n_total).astype(np.uint8))
    first_all = torch.from_numpy((rng.random(n_total) < 0.8).astype(np.uint8))
    first_all[n_total // 2:] &= torch.from_numpy((rng.random(n_total - n_total // 2) < 0.5).astype(np.uint8))   # a cheaper second half

    def fake_all(self, exprs, db, depth, prune, L):
        assert list(db) == depth_begin and depth == 3 and prune is True and L == 96
        assert exprs.n == len(strs)                  # the worker compiled the broadcast operand strings itself
        return {"len": len_all, "pool": None, "off": None, "hash": None}, first_all

    windows = []

    def fake_local(self, exprs, db, depth, prune, L, first, count, cand=None, first_flags=None, session=None):
        assert L == 96 and first % 32 == 0 and cand is not None
        windows.append((first, count))
        k = np.arange(first, first + count, dtype=np.int64)
        surv = ((k * 2654435761) >> 7) & 1
        pad = np.zeros((count + 31) // 32 * 32, np.uint8)
        pad[:count] = surv
        return torch.from_numpy(np.packbits(pad, bitorder="little").view(np.int32).copy())

    GpuBatchValidator._enumerate_all = fake_all
    GpuBatchValidator._enum_filter_local = fake_local
    gv = object.__new__(GpuBatchValidator)          # the protocol only: no device behind it
    gv.group, gv.device = "auto", torch.device("cpu")
    gv.session = core.Session.for_problem("force_free")
    if rank == 0:
        got = gv.filter_enumerated(strs, depth_begin, 3, True, 96)
        gv.shutdown()
        k = np.arange(n_total, dtype=np.int64)
        # the cost-balanced cut lies before the middle (the second half is cheaper: rank 1 takes more candidates), 32-aligned
        (lo, cnt), = windows
        q.put(bool(np.array_equal(got, (((k * 2654435761) >> 7) & 1).astype(bool))) and lo == 0 and cnt % 32 == 0 and cnt < n_total // 2 - 1000)
    else:
        assert gv.serve() == 1
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_filter_enumerated_protocol_gloo_world2():
    """GpuBatchValidator.filter_enumerated under torch.distributed: the operand strings and depth boundaries are
    broadcast, each rank answers for its own window of the candidate index space, rank 0 gets the flags of all n."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_enum_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok is True


def test_cost_bounds_tile_the_index_space():
    """Windows of the string-free last-depth filter: 32-aligned, monotone, [0 .. n], balanced by cost."""
    from pde_engine_b200.validator import GpuBatchValidator as G
    rng = np.random.default_rng(3)
    for n in (0, 1, 10, 31, 32, 33, 1000, 258285):
        ln = torch.from_numpy(rng.integers(1, 60, size=n).astype(np.uint8))
        first = torch.from_numpy((rng.random(n) < 0.7).astype(np.uint8))
        for world in (1, 2, 3, 8):
            b = G._cost_bounds(ln, first, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == n
            assert all(x <= y for x, y in zip(b[:-1], b[1:]))
            assert all(x % 32 == 0 or x == n for x in b)
    # balanced: the windows' costs differ by less than 2 % on a large, skewed batch
    n = 258285
    ln = torch.from_numpy(rng.integers(1, 60, size=n).astype(np.uint8))
    first = torch.from_numpy((rng.random(n) < np.linspace(1.0, 0.3, n)).astype(np.uint8))
    b = G._cost_bounds(ln, first, 8)
    cost = np.where(first.numpy().astype(bool), ln.numpy().astype(np.int64) + 8, 1)
    per = [int(cost[x:y].sum()) for x, y in zip(b[:-1], b[1:])]
    assert max(per) < 1.02 * min(per), per
    assert b[1] - b[0] < b[-1] - b[-2]           # cheaper candidates at the end: the last window is longer

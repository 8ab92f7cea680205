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
    hash_all[n_total - 5] = hash_all[3]            # a cross-shard exact duplicate
    surv_all[3] = surv_all[n_total - 5] = True
    surv = surv_all[first:first + count]
    words = np.zeros((count + 31) // 32, dtype=np.uint32)
    for i in np.nonzero(surv)[0]:
        words[i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    g = gather_survivors(torch.from_numpy(words.view(np.int32)), torch.from_numpy(hash_all[first:first + count].copy()), count)
    if rank == 0:
        idx, hs = merge_survivors(g)
        want = [i for i in np.nonzero(surv_all)[0] if i != n_total - 5]   # duplicate keeps the lowest index
        q.put((idx == [int(i) for i in want], [int(h) for h in hs] == [int(hash_all[i]) for i in want]))
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

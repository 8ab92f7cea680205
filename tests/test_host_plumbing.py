"""Host-side plumbing that needs no GPU: shard sizes and blob cutting of the sharded prefilter, and the two patches
tools/refcopy.py applies to a copy of the reference (the INTEGRATION.md patch and the worker import repair)."""
import os
import random

import pytest

from conftest import REPO


def test_shard_bounds_cover_the_batch_and_finish_together():
    from pde_engine_b200.validator import GpuBatchValidator as G
    for n in (0, 1, 31, 32, 4096, 143461, 10**6 + 7):
        for world in (1, 2, 3, 4, 8):
            b = G._shard_bounds(n, world)
            assert len(b) == world + 1 and b[0] == 0 and b[-1] == n
            assert all(x <= y for x, y in zip(b, b[1:]))
            assert all(x % 32 == 0 for x in b[:-1])                       # survivor words never straddle two shards
    # rank 0 packs everybody's strings and starts last: its shard is the smallest, rank 1's the largest
    sizes = [y - x for x, y in zip(G._shard_bounds(143461, 8), G._shard_bounds(143461, 8)[1:])]
    assert sizes[0] == min(sizes) and sizes[1] == max(sizes) and sizes[1:] == sorted(sizes[1:], reverse=True)
    rho = G.PACK_RATIO                                                    # finish times equal within the 32-alignment
    fin = [rho * sum(sizes[1:r + 1]) + sizes[r] for r in range(1, 8)] + [rho * sum(sizes[1:]) + sizes[0]]
    assert max(fin) - min(fin) < 200


def test_cut_blob_matches_string_slices():
    from pde_engine_b200.core import pack_strings
    from pde_engine_b200.validator import _cut_blob
    rng = random.Random(7)
    for _ in range(100):
        n = rng.randint(1, 600)
        strs = ["".join(rng.choice("rhoz+-*/() 123") for _ in range(rng.randint(0, 40))) for _ in range(n)]
        blob, cnt = pack_strings(strs)
        assert cnt == n
        bounds = sorted(set([0, n] + [rng.randint(0, n) // 32 * 32 for _ in range(4)]))
        got = []
        for b0, b1, lo, hi in _cut_blob(blob, n, bounds):
            part = blob[b0:b1].decode().split("\0")[:-1]
            assert len(part) == hi - lo
            got += part
        assert got == strs


def test_reference_patches_are_exactly_what_integration_md_says(tmp_path):
    """The drop-in test and the reference CPU baseline run from a copy of the unmodified reference with two patches;
    each must hit exactly its anchor (GM:1243, GM:1694) and change nothing else."""
    import sys
    sys.path.insert(0, REPO)
    from tools import refcopy
    if not os.path.exists(os.path.join(refcopy.BASELINE_REF, refcopy.GM)):
        pytest.skip("baseline/_ref is made by tools/refcopy.py in the build container")
    assert refcopy.patched_copy(str(tmp_path / "plain")) == ""
    d1 = refcopy.patched_copy(str(tmp_path / "gpu"), install_gpu=True, P=1024)
    plus = [l[1:].strip() for l in d1.splitlines() if l.startswith("+") and not l.startswith("+++")]
    minus = [l for l in d1.splitlines() if l.startswith("-") and not l.startswith("---")]
    assert plus == ["from pde_engine_b200.engine import install", "install(discovery, P=1024)"] and not minus
    d2 = refcopy.patched_copy(str(tmp_path / "pool"), repair_workers=True)
    plus = [l[1:].strip() for l in d2.splitlines() if l.startswith("+") and not l.startswith("+++")]
    minus = [l[1:].strip() for l in d2.splitlines() if l.startswith("-") and not l.startswith("---")]
    assert plus == ["from problems import load_problem"] and minus == ["from physics_agent.problems import load_problem"]
    # code only: no run databases, caches or images travel with the copy
    for root, _, files in os.walk(refcopy.BASELINE_REF):
        assert not [f for f in files if f.endswith((".db", ".db-wal", ".db-shm", ".png"))], root

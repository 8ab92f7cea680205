#!/usr/bin/env python3
"""Generate enumeration golden fixtures by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  The reference
writes caches next to its sources, so it is imported from a scratch copy
(code only, no committed ``*.db`` caches -> fresh caches).

What is recorded, per problem and depth:
  * the ordered candidate strings the reference hands to its normaliser
    (captured by wrapping ``normalize_batch``; LBF:200-202),
  * the ``on_batch`` calls (ordered new uniques per 2000-candidate chunk;
    LBF:204-212),
  * candidate -> normalised string pairs.

Depth-4 normalisation costs ~30 min single threaded (SymPy ``expand``), so the
reference's own SQLite cache is pre-warmed in parallel by calling the
reference's own ``LeanNormalizer.normalize`` (LB:67-92) on the candidate
strings produced by ``oracle.enumerate`` -- if the oracle's candidate list
differed from the reference's, the run below would simply miss the cache and
the recorded lists (which come from the reference) would expose it.

Usage:  python tests/golden/make_golden_enum.py force_free 4
        python tests/golden/make_golden_enum.py kerr_magnetosphere 3
"""
from __future__ import annotations

import gzip
import hashlib
import json
import multiprocessing as mp
import os
import shutil
import sqlite3
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF_SRC = "/root/reference"
WORK = os.environ.get("PDE_REF_WORK", "/tmp/pde_ref_work")


def make_work_copy() -> str:
    if not os.path.isdir(WORK):
        def ignore(d, names):
            return [n for n in names if n.endswith((".db", ".db-shm", ".db-wal", ".png")) or n == ".git"]
        shutil.copytree(REF_SRC, WORK, ignore=ignore)
    return WORK


_norm = None


def _init_worker(work):
    global _norm
    sys.path.insert(0, work)
    os.chdir(work)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        from lean_normalizer.lean_bridge import LeanNormalizer as Base
        _norm = Base()


def _normalize_one(s):
    return s, _norm.normalize(s)


def main():
    problem = sys.argv[1]
    max_depth = int(sys.argv[2])
    full_list_max_depth = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    work = make_work_copy()
    sys.path.insert(0, work)
    sys.path.insert(0, REPO)
    os.chdir(work)
    from lean_normalizer.lean_bridge_fixed import LeanNormalizer, FastExpressionGenerator
    from problems import load_problem
    from oracle import enumerate as oenum

    spec = load_problem(problem)
    prim_strs = [str(p) for p in spec.primitives]
    cache_db = os.path.join(work, f"golden_norm_cache_{spec.slug}.db")
    if os.path.exists(cache_db):
        os.remove(cache_db)
    normalizer = LeanNormalizer(cache_db=cache_db)

    # ---- pre-warm the reference's cache depth by depth (parallel) ----
    pool = mp.Pool(min(8, os.cpu_count() or 1), initializer=_init_worker, initargs=(work,))
    E = {1: list(prim_strs)}
    seen = set()
    t0 = time.time()
    for depth in range(2, max_depth + 1):
        cands = oenum.candidates_for_depth(E, depth)
        distinct = list(dict.fromkeys(cands))
        pairs = pool.map(_normalize_one, distinct, chunksize=64)
        for s, n in pairs:
            normalizer.cache_conn.execute(
                "INSERT OR REPLACE INTO normalized_cache (expr_hash, expr_str, normalized) VALUES (?, ?, ?)",
                (hashlib.sha256(s.encode()).hexdigest(), s, n))
        normalizer.cache_conn.commit()
        m = dict(pairs)
        uniq = []
        for s in cands:
            sig = oenum.signature(m[s])
            if sig not in seen:
                seen.add(sig)
                uniq.append(m[s])
        E[depth] = uniq
        print(f"[prewarm] depth {depth}: {len(cands)} cands, {len(distinct)} distinct, {len(uniq)} uniq, {time.time()-t0:.1f}s", flush=True)
    pool.close()

    # ---- the reference run (unmodified generator + its own cache) ----
    cand_log = {}
    pair_log = {}
    orig = normalizer.normalize_batch

    def wrapped(batch):
        res = orig(batch)
        for (s, d), r in zip(batch, res):
            cand_log.setdefault(d, []).append(s)
            pair_log.setdefault(d, {})[s] = r["normalized"]
        return res

    normalizer.normalize_batch = wrapped
    gen = FastExpressionGenerator(normalizer)
    batches = {}

    def on_batch(depth, exprs):
        batches.setdefault(depth, []).append(list(exprs))

    t1 = time.time()
    gen.stream_generate(primitives=spec.primitives, unary_ops=spec.unary_ops,
                        binary_ops=spec.all_binary_ops, max_depth=max_depth,
                        batch_size=2000, on_batch=on_batch, prune=True)
    print(f"[reference] stream_generate took {time.time()-t1:.1f}s", flush=True)

    out = {"problem": spec.slug, "primitives": prim_strs, "batch_size": 2000,
           "sympy": __import__("sympy").__version__, "depths": {}}
    for depth in range(1, max_depth + 1):
        uniq = [s for b in batches.get(depth, []) for s in b]
        rec = {"uniques": uniq, "batch_sizes": [len(b) for b in batches.get(depth, [])]}
        if depth >= 2:
            cands = cand_log[depth]
            rec["n_candidates"] = len(cands)
            rec["candidates_sha256"] = hashlib.sha256("\n".join(cands).encode()).hexdigest()
            if depth <= full_list_max_depth:
                rec["candidates"] = cands
                rec["normalized"] = [pair_log[depth][s] for s in cands]
        out["depths"][str(depth)] = rec
        print(depth, len(uniq), rec.get("n_candidates"))
    path = os.path.join(REPO, "tests", "golden", f"enum_{spec.slug}_d{max_depth}.json.gz")
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()

"""PoolNormalizer -- the CPU canonicaliser on all host cores.

The canonicaliser stays on the CPU by design (BASELINE.json north_star) and the
generator accepts ANY object with the reference's `normalize_batch` protocol
(lean_normalizer/lean_bridge_fixed.py:42-68) -- in production the reference's own
`LeanNormalizer`.  That object is single threaded and commits to SQLite once per
cache miss (LBF:60-64): 6.8 ms per candidate, 29 minutes for the 258 285 depth-4
candidates, and it sits between every two GPU depths.  This class
is the same contract on a process pool:

* `normalize(expr_str) -> str`: what "Lean normalisation" is at the surveyed
  commit (lean_bridge.py:67-112): `sympify` WITHOUT locals (the six generator
  names neg/inv/square/... stay opaque functions), `expand`, `collect` over
  (rho, z) when both occur, the five substitution rules, `str`; any exception
  returns the input string unchanged.
* `normalize_batch([(expr_str, index), ...]) -> [{'normalized', 'index',
  'signature'}]` in input order, `signature = sha256(normalized)[:16]`.

Misses are normalised in chunks by worker processes (`spawn`: safe next to a CUDA
context); results are memoised in memory and, if `cache_db` is given, written to
the reference's `normalized_cache` table in ONE transaction per batch.
"""
from __future__ import annotations

import hashlib
import multiprocessing as mp
import os
import sqlite3
from typing import Any, Dict, List, Optional, Sequence, Tuple


def canonical_string(expr_str: str) -> str:
    """One expression through the reference's canonical form (LB:67-112)."""
    import sympy as sp
    try:
        e = sp.expand(sp.sympify(expr_str))
        r_plain, z_plain = sp.Symbol("rho"), sp.Symbol("z")
        if e.has(r_plain) and e.has(z_plain):
            e = sp.collect(e, [r_plain, z_plain])
        # the paper's rewrite rules, stated on a positive rho (LB:95-110)
        rp, zz = sp.Symbol("rho", positive=True), sp.Symbol("z")
        for old, new in ((sp.exp(sp.log(rp)), rp), (sp.log(sp.exp(zz)), zz), (sp.sqrt(rp ** 2), rp),
                         (rp / rp, 1), (zz - zz, 0)):
            e = e.subs(old, new)
        return str(e)
    except Exception:
        return expr_str


def _chunk(strs: Sequence[str]) -> List[str]:
    return [canonical_string(s) for s in strs]


class PoolNormalizer:
    def __init__(self, n_workers: Optional[int] = None, cache_db: Optional[str] = None, chunk: int = 64,
                 start_method: str = "spawn"):
        self.n_workers = max(1, n_workers or (os.cpu_count() or 1))
        self.chunk = chunk
        self._memo: Dict[str, str] = {}
        self._pool = None
        self._ctx = mp.get_context(start_method)
        self.stats = {"requests": 0, "misses": 0}
        self._db = None
        if cache_db:
            self._db = sqlite3.connect(cache_db)
            self._db.execute("CREATE TABLE IF NOT EXISTS normalized_cache (expr_hash TEXT PRIMARY KEY, expr_str TEXT, "
                             "normalized TEXT, timestamp DATETIME DEFAULT CURRENT_TIMESTAMP)")        # LBF:28-38
            for s, n in self._db.execute("SELECT expr_str, normalized FROM normalized_cache"):
                self._memo[s] = n

    # ---- lifecycle ----
    def _ensure_pool(self):
        if self._pool is None and self.n_workers > 1:
            self._pool = self._ctx.Pool(self.n_workers)
        return self._pool

    def close(self) -> None:
        if self._pool is not None:
            self._pool.terminate()
            self._pool.join()
            self._pool = None
        if self._db is not None:
            self._db.commit()
            self._db.close()
            self._db = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- the reference's protocol ----
    def normalize(self, expr_str: str) -> str:
        hit = self._memo.get(expr_str)
        if hit is None:
            hit = canonical_string(expr_str)
            self._memo[expr_str] = hit
        return hit

    def normalize_many(self, strs: Sequence[str]) -> List[str]:
        self.stats["requests"] += len(strs)
        missing = list(dict.fromkeys(s for s in strs if s not in self._memo))
        if missing:
            self.stats["misses"] += len(missing)
            pool = self._ensure_pool() if len(missing) >= 2 * self.chunk else None
            if pool is None:
                done = _chunk(missing)
            else:
                parts = [missing[i:i + self.chunk] for i in range(0, len(missing), self.chunk)]
                done = [n for part in pool.map(_chunk, parts) for n in part]
            self._memo.update(zip(missing, done))
            if self._db is not None:                      # one transaction per batch (the reference: one per miss)
                self._db.executemany("INSERT OR REPLACE INTO normalized_cache (expr_hash, expr_str, normalized) VALUES (?, ?, ?)",
                                     [(hashlib.sha256(s.encode()).hexdigest(), s, n) for s, n in zip(missing, done)])
                self._db.commit()
        return [self._memo[s] for s in strs]

    def normalize_batch(self, expressions: List[Tuple[str, int]]) -> List[Dict[str, Any]]:
        norms = self.normalize_many([s for s, _ in expressions])
        return [{"normalized": n, "index": idx, "signature": hashlib.sha256(n.encode()).hexdigest()[:16]}
                for n, (_, idx) in zip(norms, expressions)]

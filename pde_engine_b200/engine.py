"""Glue between the GPU hot path and the reference's engine.

Two ways in:

* ``install(discovery)`` -- the drop-in: takes the reference's
  ``GeneralFoliationDiscovery`` instance (GM:62-132) and replaces its two
  injection points, ``fast_generator`` (GM:107) and ``validator`` (GM:73-75),
  with ``GpuExpressionGenerator`` / ``GpuBatchValidator``.  Everything else of
  the reference (CLI, run database, writer/monitor processes, reports) keeps
  running unchanged; the generator filters each ``on_batch`` chunk on the device
  right before ``emit_to_db`` (GM:1251) sees it, so the per-row
  ``validator.validate`` calls are cache hits that only reach the CPU validator
  for survivors.  CUDA must be initialised inside the generator process
  (GM:1243 rebuilds the discovery object there), so call ``install`` in that
  process.

* ``run_discovery(...)`` -- a self-contained driver with the semantics of
  ``_parallel_generator_worker.emit_to_db`` (GM:1251-1411, inline validation)
  and the reference's run-DB schema (GM:655-683), used by the tests and the
  depth-4 wall-time measurement where the reference tree is not available.
"""
from __future__ import annotations

import hashlib
import json
import sqlite3
import time
from typing import Any, Callable, Dict, List, Optional

from .generator import GpuExpressionGenerator
from .validator import GpuBatchValidator


def install(discovery: Any, P: int = 4096, **gpu_kwargs) -> Any:
    """Swap the two hot-path objects of a reference ``GeneralFoliationDiscovery``."""
    slug = discovery.problem.slug
    gv = GpuBatchValidator(discovery.validator, slug, P=P, sympify_locals=dict(discovery._sympify_locals), **gpu_kwargs)
    discovery.validator = gv
    discovery.problem.validator = gv
    discovery.fast_generator = GpuExpressionGenerator(discovery.normalizer, slug, pre_batch_hook=gv.prefetch)
    return discovery


SCHEMA = """
CREATE TABLE IF NOT EXISTS {table} (
    id INTEGER PRIMARY KEY AUTOINCREMENT,
    expression TEXT NOT NULL,
    normalized TEXT NOT NULL UNIQUE,
    signature INTEGER,
    depth INTEGER NOT NULL,
    validation_status TEXT DEFAULT 'pending',
    is_valid BOOLEAN,
    validation_reason TEXT,
    validator_method TEXT,
    validator_math TEXT,
    is_paper_solution BOOLEAN DEFAULT 0,
    paper_solution_name TEXT,
    created_at TIMESTAMP DEFAULT CURRENT_TIMESTAMP,
    validated_at TIMESTAMP,
    validator_evidence TEXT
)"""   # GM:655-683


def run_discovery(spec: Any, normalizer: Any, max_depth: int, *, db_path: Optional[str] = None,
                  run_id: str = "gpu_run", db_normalize: Optional[Callable[[str], str]] = None,
                  is_degenerate: Optional[Callable[[Any], bool]] = None, batch_size: int = 2000,
                  match_known: bool = True, confirm_pool: Any = None,
                  share_confirmations: Any = None) -> Dict[str, Any]:
    """Enumerate to ``max_depth`` and validate every unique (GM:1251-1411 semantics).

    spec          ProblemSpec whose ``validator`` is a GpuBatchValidator
    normalizer    the CPU canonicaliser (``normalize_batch``)
    db_normalize  the DB-normalisation ``str(simplify(expand(.)))`` of GM:1267-1278 (CPU,
                  out of scope; returns None for rows the reference drops as degenerate);
                  None = use the expression string (no DB dedup)
    is_degenerate ``_has_degenerate_denominator`` of GM:134-199 (CPU, out of scope)
    confirm_pool  a ``confirm.ConfirmationPool``: the survivors of each chunk are confirmed by its CPU
                  worker processes (most plausible first, wall cap per candidate) instead of inline;
                  rows that hit the cap get ``is_valid = None`` / status ``'timeout'``
    share_confirmations  opt-in, with ``confirm_pool``: a ``fingerprint.GpuFingerprinter``.  Survivors that
                  denote the same FUNCTION (equal device fingerprints) are confirmed once --
                  by their shortest string -- and the verdict is copied to the others, with the
                  representative named in ``validator_evidence``.  Every row is still stored; the reference
                  confirms each string on its own, so this is a deviation wherever SymPy decides two forms
                  of one function differently
    Returns {'rows': [...], 'stats': {...}}; rows carry the run-DB columns.
    """
    import sympy as sp
    gv: GpuBatchValidator = spec.validator
    locs = spec.sympify_locals()
    gen = GpuExpressionGenerator(normalizer, getattr(spec, "coordinate_system", spec.slug), pre_batch_hook=gv.prefetch)
    rows: List[dict] = []
    seen_norm = set()
    t_start = time.time()
    conn = cur = None
    table = f"expressions_{run_id.replace('-', '_')}"
    if db_path:
        conn = sqlite3.connect(db_path)
        cur = conn.cursor()
        cur.execute(SCHEMA.format(table=table))
    desc = gv.describe()
    coord = list(spec.symbols.values())
    known = {}
    if match_known:
        for s, name in spec.known_solutions.items():
            try:
                known[sp.sympify(s, locals=locs)] = name
            except Exception:
                pass

    confirmed_functions: Dict[int, tuple] = {}     # fingerprint key -> (representative string, is_valid, reason, paper)
    stats_extra = {"confirmations_shared": 0}

    def insert(row: dict) -> None:
        rows.append(row)
        if cur is not None:
            cur.execute(f"INSERT INTO {table} (expression, normalized, signature, depth, validation_status, is_valid,"
                        " validation_reason, validator_method, validator_math, is_paper_solution, paper_solution_name,"
                        " validator_evidence, validated_at) VALUES (?,?,?,?,?,?,?,?,?,?,?,?,CURRENT_TIMESTAMP)",
                        (row["expression"], row["normalized"], row["signature"], row["depth"], row["validation_status"],
                         row["is_valid"], row["validation_reason"], row["validator_method"], row["validator_math"],
                         row["is_paper_solution"], row["paper_solution_name"], row["validator_evidence"]))

    def emit(depth: int, expr_list: List[str]) -> None:
        pending: List[dict] = []          # rows of this chunk, in order; survivors wait for the pool
        for expr_str in expr_list:
            try:
                u = sp.sympify(expr_str, locals=locs)                      # GM:1257
            except Exception:
                u = None
            if u is not None and is_degenerate is not None and is_degenerate(u):
                continue                                                   # GM:1261
            normalized = db_normalize(expr_str) if db_normalize else expr_str     # GM:1267-1278
            if normalized is None:                                         # degenerate without locals, GM:1270-1276
                continue
            if normalized in seen_norm:                                    # UNIQUE(normalized), GM:1407
                continue
            seen_norm.add(normalized)
            sig = int(hashlib.sha256(normalized.encode()).hexdigest()[:8], 16)   # GM:1281
            row = dict(id=0, expression=expr_str, normalized=normalized, signature=sig, depth=depth,
                       validator_method=desc.get("method_name"), validator_math=desc.get("math_definition"),
                       is_paper_solution=0, paper_solution_name=None)
            wait = False
            try:
                if u is None:
                    u = sp.sympify(expr_str, locals=locs)
                if not any(u.has(v) for v in coord):
                    is_valid, reason, evidence = False, "constant-only (skipped)", gv.last_evidence()   # GM:1293-1294
                elif confirm_pool is not None:
                    survivor, evidence, reason = gv.gpu_verdict(u)
                    is_valid = False
                    wait = survivor
                else:
                    is_valid, reason = gv.validate(u, check_regularity=False, fast_point_only=False,
                                                   lean_first=True, defer_heavy_checks=True, enforce_anchor=False)
                    evidence = gv.last_evidence()
            except Exception as e:                                         # GM:1336-1339
                is_valid, reason, evidence = None, f"Validator Error: {e}", {}
            paper = None
            if not wait and is_valid and known:                            # GM:1785-1798 (the worker path's matching)
                for k_expr, name in known.items():
                    try:
                        if sp.simplify(u - k_expr) == 0:
                            paper = name
                            break
                    except Exception:
                        pass
            row.update(is_valid=is_valid, validation_reason=reason, evidence=evidence, wait=wait, paper=paper)
            pending.append(row)
        if confirm_pool is not None:
            # most plausible first: smallest residual ratio, unevaluated candidates last
            todo = [k for k, r in enumerate(pending) if r["wait"]]
            todo.sort(key=lambda k: (pending[k]["evidence"].get("n_finite", 0) <= 0, pending[k]["evidence"].get("ratio_max", 0.0)))
            shared_from: Dict[int, str] = {}
            verdicts: Dict[int, tuple] = {}
            reps: List[tuple] = []
            ask = todo
            if share_confirmations is not None and todo:
                keys = share_confirmations.fingerprint([pending[k]["expression"] for k in todo]).key.tolist()
                groups: Dict[int, List[int]] = {}
                for k, key in zip(todo, keys):
                    if key:
                        groups.setdefault(key, []).append(k)
                ask_set = {k for k, key in zip(todo, keys) if not key}                  # unknown: confirmed on their own
                for key, members in groups.items():
                    if key in confirmed_functions:                                      # settled earlier in this run
                        for k in members:
                            shared_from[k] = confirmed_functions[key][0]
                            verdicts[k] = confirmed_functions[key][1:]
                        continue
                    rep = min(members, key=lambda k: len(pending[k]["expression"]))     # shortest form: cheapest for SymPy
                    reps.append((key, rep, members))
                    ask_set.add(rep)
                ask = [k for k in todo if k in ask_set]                                 # keeps the most-plausible-first order
            verdicts.update(zip(ask, confirm_pool.confirm([pending[k]["expression"] for k in ask])))
            for key, rep, members in reps:
                if verdicts[rep][0] is not None:                                        # a timeout is not a verdict
                    confirmed_functions[key] = (pending[rep]["expression"],) + tuple(verdicts[rep])
                for k in members:
                    if k != rep:
                        shared_from[k] = pending[rep]["expression"]
                        verdicts[k] = verdicts[rep]
            for k in todo:
                ok, reason, paper = verdicts[k]
                ev = {"gpu": pending[k]["evidence"], "confirmed_by": "confirm.ConfirmationPool"}
                if k in shared_from:
                    ev["same_function_as"] = shared_from[k]
                    stats_extra["confirmations_shared"] += 1
                else:
                    gv.stats["cpu_confirmed"] += 1
                pending[k].update(is_valid=ok, validation_reason=reason, paper=paper, evidence=ev)
        for r in pending:
            ok, reason = r["is_valid"], r["validation_reason"]
            status = "completed" if ok is not None else ("timeout" if str(reason).startswith("Timeout") else "error")
            row = {k: r[k] for k in ("expression", "normalized", "signature", "depth", "validator_method", "validator_math")}
            row.update(id=len(rows) + 1, validation_status=status, is_valid=ok, validation_reason=reason,
                       is_paper_solution=int(r["paper"] is not None), paper_solution_name=r["paper"],
                       validator_evidence=json.dumps(r["evidence"], default=str))
            insert(row)
        if conn is not None:
            conn.commit()

    gen.stream_generate(primitives=spec.primitives, unary_ops=spec.unary_ops, binary_ops=spec.all_binary_ops,
                        max_depth=max_depth, batch_size=batch_size, on_batch=emit, prune=True)
    if conn is not None:
        conn.close()
    stats = dict(gv.stats)
    stats.update(rows=len(rows), valid=sum(1 for r in rows if r["is_valid"]), wall_s=time.time() - t_start,
                 enumeration=gen.stats, **stats_extra)
    return {"rows": rows, "stats": stats, "table": table}

"""Reports and heavy checks from the run database with the device in front of SymPy.

Two DB-only passes of the reference's engine, kept column- and output-compatible:

* ``novel_equivalence_classes`` / ``generate_report_from_db`` -- the "novel solutions, deduplicated by mathematical
  equivalence" block of ``_generate_report_from_db`` (general_method_paper_reproduction.py:1826-2020, bucketing
  GM:1918-2008).  The reference runs, PER valid row, up to 7 ``simplify(e - known)`` calls (GM:1937-1945) and a
  canonicalisation pipeline ``together -> cancel -> powsimp -> powdenest -> simplify -> rewrite -> together(cancel)``
  (GM:1921-1935), then groups by the srepr of the result.  Here one ``pde_fingerprint`` pass evaluates every row (and
  every known solution) on the device (64 points, both signs of the second coordinate, generic parameter values);
  rows whose values differ by more than 1e-7 relative somewhere are DIFFERENT functions, so

    - a row can only equal a known solution whose fingerprint it shares: ``simplify(e - known)`` runs for those pairs only;
    - a row can only share a class with rows of its fingerprint bucket: the canonicalisation pipeline runs only in
      buckets with more than one member (a singleton bucket is a class of its own, whatever its canonical form);
    - rows the device cannot evaluate (key 0) take the reference's full path.

  The classes, their sizes, representatives (``_rep_cost``, GM:1953-1970) and print order (GM:2010) are the
  reference's: SymPy still decides every merge, the device only proves non-equivalence.

* ``heavy_validate_from_db`` -- GM:2024-2136: re-validates rows with ``defer_heavy_checks=False`` (finiteness,
  regularity, small-spin anchor; KV:318-341) and writes ``heavy_is_valid / heavy_reason / heavy_validated_at``.  The
  heavy checks sit BEHIND the exact-zero test of the residual (KV:265-316), so with scope 'all' the reference
  spends ~0.15 s of SymPy per row to learn again that the residual is not zero; here the batched residual filter
  answers that for the whole table in one launch and only its survivors reach the validator's heavy path.
"""
from __future__ import annotations

import sqlite3
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple


# ---- the reference's helpers, GM:1921-1970 (semantics must match: they decide merges and representatives) ----
def canonical_key(e) -> str:
    import sympy as sp
    try:
        en = sp.together(e)
        en = sp.cancel(en)
        en = sp.powsimp(en, force=True)
        en = sp.powdenest(en, force=True)
        en = sp.simplify(en)
        en = en.rewrite(sp.Pow)
        en = sp.together(sp.cancel(en))
        return sp.srepr(en)
    except Exception:
        try:
            return sp.srepr(sp.simplify(e))
        except Exception:
            return str(e)


def rep_cost(e) -> tuple:
    import sympy as sp

    def depth(x) -> int:
        try:
            return 1 + max((depth(a) for a in x.args), default=0)
        except Exception:
            return 1
    try:
        c_ops = int(sp.count_ops(e, visual=False))
    except Exception:
        c_ops = 10 ** 6
    try:
        d = depth(e)
    except Exception:
        d = 999999
    try:
        s_len = len(sp.srepr(e))
    except Exception:
        s_len = len(str(e))
    try:
        pen = 10 * int(e.has(sp.zoo)) + 20 * int(e.has(sp.nan)) + 5 * int(any(isinstance(a, sp.Float) for a in e.atoms(sp.Float)))
    except Exception:
        pen = 0
    return (c_ops, d, s_len, pen)


def _value_components(fingerprinter: Any, strs: List[str], rtol: float = 1e-7) -> List[int]:
    """Connected components of "values agree within rtol wherever both are finite on the fingerprint grid" (1-based ids, 0 = no finite value)."""
    import numpy as np
    keep = fingerprinter.keep_values
    fingerprinter.keep_values = True
    try:
        fp = fingerprinter.fingerprint(strs)
    finally:
        fingerprinter.keep_values = keep
    V = np.asarray(fp.values, dtype=np.float64)
    n = V.shape[0]
    fin = np.isfinite(V)
    Vz = np.where(fin, V, 0.0)
    parent = list(range(n))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    # pairwise on the points where BOTH rows are finite (two forms of one function may overflow at different points;
    # keeping such pairs together only costs a SymPy call).  n is the number of VALID rows of a run: cheap.
    with np.errstate(all="ignore"):
        for i in range(n - 1):
            if not fin[i].any():
                continue
            common = fin[i + 1:] & fin[i]
            scale = np.maximum(np.maximum(np.abs(Vz[i]), np.abs(Vz[i + 1:])), 1e-3)
            agree = (np.abs(Vz[i + 1:] - Vz[i]) <= rtol * scale) | ~common
            close = agree.all(axis=1) & (common.sum(axis=1) >= 1)
            for j in np.flatnonzero(close):
                a, b = find(i), find(i + 1 + int(j))
                if a != b:
                    parent[b] = a
    return [0 if not fin[i].any() else find(i) + 1 for i in range(n)]


def novel_equivalence_classes(rows: Sequence[Tuple[int, str]], sympify_locals: dict, known_solutions: Dict[str, str],
                              fingerprinter: Any = None) -> Tuple[List[dict], dict]:
    """rows: (id, expression) of the valid non-paper rows in id order.  Returns (classes, stats); classes are
    ``{'rep_id', 'rep_str', 'size'}`` in the reference's print order.  fingerprinter=None: the reference's own
    path for every row (what the device path is tested against when no golden report exists)."""
    import sympy as sp
    rows = [(int(i), s) for i, s in rows]
    stats = {"rows": len(rows), "known_checks": 0, "known_checks_reference": 0, "canonical_keys": 0,
             "canonical_keys_reference": 0, "device_unknown": 0}
    known = []
    for ks in known_solutions:
        try:
            known.append((ks, sp.sympify(ks, locals=sympify_locals)))
        except Exception:
            pass
    # Device screen.  comp[i] = id of the set of rows (and known solutions) that MAY denote row i's function: the same
    # finiteness pattern and values within 1e-7 relative at every point of the fingerprint grid (connected components;
    # a tolerance, not the exact 64-bit key, so that two forms of one function can never be separated by a value that
    # sits on a rounding boundary).  0 = the device could not evaluate the row: it takes the reference's full path.
    keys = [0] * len(rows)
    known_keys: List[int] = [0] * len(known)
    if fingerprinter is not None and rows:
        comp = _value_components(fingerprinter, [s for _, s in rows] + [ks for ks, _ in known])
        keys, known_keys = comp[:len(rows)], comp[len(rows):]
        stats["device_unknown"] = sum(1 for k in keys if k == 0)
    bucket_size: Dict[int, int] = {}
    for k in keys:
        if k:
            bucket_size[k] = bucket_size.get(k, 0) + 1

    buckets: Dict[str, dict] = {}
    for (expr_id, expr_str), k in zip(rows, keys):
        try:
            expr = sp.sympify(expr_str, locals=sympify_locals)
        except Exception:                                          # GM:1976-1984
            key = f"RAW::{expr_str}"
            entry = buckets.get(key)
            if entry is None:
                buckets[key] = {"rep_id": expr_id, "rep_str": expr_str, "rep_expr": None, "size": 1}
            else:
                entry["size"] += 1
            continue
        # equivalent to a known solution?  (GM:1986-1991)  Only a known solution with the same fingerprint can be.
        stats["known_checks_reference"] += len(known)
        is_known = False
        for (ks, kexpr), kk in zip(known, known_keys):
            if fingerprinter is not None and k and kk and k != kk:
                continue
            stats["known_checks"] += 1
            try:
                if sp.simplify(expr - kexpr) == 0:
                    is_known = True
                    break
            except Exception:
                pass
        if is_known:
            continue
        stats["canonical_keys_reference"] += 1
        if fingerprinter is not None and k and bucket_size.get(k, 0) == 1:
            key = f"FP::{k}"                                       # alone in its component: a class of its own
        else:
            stats["canonical_keys"] += 1
            try:
                key = canonical_key(expr)
            except Exception:
                key = str(expr)
        entry = buckets.get(key)
        if entry is None:
            buckets[key] = {"rep_id": expr_id, "rep_str": expr_str, "rep_expr": expr, "size": 1}
        else:
            entry["size"] += 1
            try:                                                    # GM:2003-2008: prefer a simpler representative
                if rep_cost(expr) < rep_cost(entry["rep_expr"] if entry["rep_expr"] is not None else expr):
                    entry["rep_id"], entry["rep_str"], entry["rep_expr"] = expr_id, expr_str, expr
            except Exception:
                pass
    items = sorted(buckets.values(), key=lambda e: (-e["size"], e["rep_str"]))           # GM:2010
    return [{"rep_id": e["rep_id"], "rep_str": e["rep_str"], "size": e["size"]} for e in items], stats


def generate_report_from_db(db_path: str, table: str, spec: Any, fingerprinter: Any = None, out: Callable[[str], None] = print) -> dict:
    """The report of ``_generate_report_from_db`` (GM:1826-2020) from a run database: same figures, same lines."""
    con = sqlite3.connect(db_path)
    cur = con.cursor()
    try:
        total, valid, paper_distinct = cur.execute(
            f"SELECT COUNT(*), SUM(CASE WHEN is_valid = 1 THEN 1 ELSE 0 END), "
            f"COUNT(DISTINCT CASE WHEN is_paper_solution = 1 THEN signature END) FROM {table}").fetchone()
        paper = cur.execute(f"SELECT paper_solution_name, MIN(expression), MIN(id) FROM {table} WHERE is_paper_solution = 1 "
                            f"GROUP BY signature, paper_solution_name ORDER BY paper_solution_name").fetchall()
        depth_counts = cur.execute(f"SELECT depth, COUNT(*) FROM {table} GROUP BY depth ORDER BY depth").fetchall()
        novel_rows = cur.execute(f"SELECT id, expression FROM {table} WHERE is_valid = 1 AND "
                                 f"(is_paper_solution IS NULL OR is_paper_solution = 0)").fetchall()
    finally:
        con.close()
    locs = spec.sympify_locals() if hasattr(spec, "sympify_locals") else dict(spec)
    classes, stats = novel_equivalence_classes(novel_rows, locs, dict(getattr(spec, "known_solutions", {}) or {}), fingerprinter)
    out(f"Total expressions generated: {total}")
    out(f"Valid foliations found: {valid or 0}")
    out(f"Known solutions found: {paper_distinct or 0} (distinct canonical forms)")
    out("\nExpression counts by depth:")
    for d, c in depth_counts:
        out(f"  Depth {d}: {c}")
    if paper:
        out("\nKnown solutions found (deduplicated by signature):")
        for name, expr, ex_id in paper:
            out(f"  ✓ {name} (id={ex_id}): {expr}")
    out("\nNovel solutions (Lean-valid, not matching known set; deduplicated by mathematical equivalence):")
    out(f"Novel valid rows (non-paper): {len(novel_rows)}")
    out(f"Novel equivalence classes: {len(classes)}")
    for c in classes:
        out(f"  • id={c['rep_id']} size={c['size']} expr={c['rep_str']}")
    if not classes:
        out("  (none)")
    return {"total": total, "valid": valid or 0, "paper_distinct": paper_distinct or 0, "depth_counts": depth_counts,
            "paper_solutions": paper, "novel_rows": len(novel_rows), "classes": classes, "stats": stats}


HEAVY_KW = dict(fast_point_only=False, lean_first=True, defer_heavy_checks=False)


def heavy_validate_from_db(db_path: str, table: str, spec: Any, scope: str = "valid", check_regularity: bool = True,
                           enforce_anchor: bool = True, anchor_target: str = "either", out: Callable[[str], None] = print) -> dict:
    """GM:2024-2136 with the batched residual filter in front.  ``spec.validator`` is a GpuBatchValidator wrapping the
    problem's CPU validator; rows the device rejects get the reference's fast-point-check verdict (KV:265-271) without
    SymPy, the others go through ``validate(..., defer_heavy_checks=False, enforce_anchor=...)`` exactly as GM:2087-2099."""
    import sympy as sp
    assert scope in ("valid", "all")
    con = sqlite3.connect(db_path)
    cur = con.cursor()
    cols = {r[1] for r in cur.execute(f"PRAGMA table_info({table})")}
    for col, typ in (("heavy_is_valid", "BOOLEAN"), ("heavy_reason", "TEXT"), ("heavy_validated_at", "TIMESTAMP")):   # GM:2036-2044
        if col not in cols:
            cur.execute(f"ALTER TABLE {table} ADD COLUMN {col} {typ}")
    con.commit()
    where = "WHERE is_valid = 1" if scope == "valid" else ""
    rows = cur.execute(f"SELECT id, expression FROM {table} {where} ORDER BY id").fetchall()
    gv = spec.validator
    cpu = getattr(gv, "cpu_validator", None) or gv
    for name, val in (("monopole_target", anchor_target), ("require_monopole_extension", bool(enforce_anchor))):      # GM:2071-2078
        if hasattr(cpu, name):
            setattr(cpu, name, val)
    locs = spec.sympify_locals()
    stats = {"rows": len(rows), "device_rejected": 0, "cpu_heavy": 0, "ok": 0, "fail": 0}
    strs = [s for _, s in rows]
    bv = gv.prefilter(strs) if (hasattr(gv, "prefilter") and strs) else None
    updates = []
    for i, (expr_id, expr_str) in enumerate(rows):
        try:
            if bv is not None and not bv.survivor[i]:
                stats["device_rejected"] += 1
                r0 = None if bv.ref_rs is None else float(bv.ref_rs[i][0][0])
                is_valid, reason = False, gv._reject_reason(r0, bv.evidence(i))
            else:
                stats["cpu_heavy"] += 1
                u = sp.sympify(expr_str, locals=locs)
                try:
                    is_valid, reason = cpu.validate(u, check_regularity=check_regularity, enforce_anchor=enforce_anchor, **HEAVY_KW)
                except TypeError:                                   # GM:2096-2099
                    is_valid, reason = cpu.validate(u, check_regularity=check_regularity, fast_point_only=False)
            stats["ok" if is_valid else "fail"] += 1
            updates.append((None if is_valid is None else bool(is_valid), reason, expr_id))
        except Exception as e:                                      # GM:2109-2111
            stats["fail"] += 1
            updates.append((None, f"Heavy validator error: {e}", expr_id))
    cur.executemany(f"UPDATE {table} SET heavy_is_valid = ?, heavy_reason = ?, heavy_validated_at = CURRENT_TIMESTAMP WHERE id = ?", updates)
    con.commit()
    con.close()
    out(f"Heavy summary: ok={stats['ok']} / {len(rows)} (failed {stats['fail']})")
    return stats


__all__ = ["novel_equivalence_classes", "generate_report_from_db", "heavy_validate_from_db", "canonical_key", "rep_cost"]

"""SURVEY 8f rank 4 on the device: the report's equivalence bucketing behind pde_fingerprint (GM:1918-2008) against the
reference's own report, and heavy validation from the run database (GM:2024-2136) behind the batched residual filter
against the validator alone."""
import json
import os
import sqlite3
import time

import pytest

from conftest import GOLDEN, load_golden
from oracle import symbolic as osym

pytestmark = pytest.mark.gpu


def _make_db(path, table, rows):
    from pde_engine_b200.engine import SCHEMA
    con = sqlite3.connect(path)
    con.execute(SCHEMA.format(table=table))
    for r in rows:
        con.execute(f"INSERT INTO {table} (id, expression, normalized, signature, depth, validation_status, is_valid, validation_reason)"
                    " VALUES (?,?,?,?,?,?,?,?)", (r["id"], r["expression"], r["normalized"], r["signature"], r["depth"],
                                                  r.get("status") or r.get("validation_status"), r["is_valid"],
                                                  r.get("reason") or r.get("validation_reason")))
    con.commit()
    con.close()


def test_report_classes_equal_the_reference_report(cuda_device, tmp_path):
    from pde_engine_b200.fingerprint import GpuFingerprinter
    from pde_engine_b200.problems import load_problem
    from pde_engine_b200.report import generate_report_from_db
    gold = json.load(open(os.path.join(GOLDEN, "report_force_free_d2.json")))
    fx = load_golden("ref_fixtures.json")["ff_run_db"]
    db = str(tmp_path / "run.db")
    _make_db(db, "expressions_t", fx)
    spec = load_problem("force_free", make_gpu=False)
    lines = []
    t0 = time.time()
    rep = generate_report_from_db(db, "expressions_t", spec, GpuFingerprinter("force_free"), out=lines.append)
    wall = time.time() - t0
    assert rep["classes"] == gold["classes"] and rep["novel_rows"] == gold["novel_rows"] == 62
    st = rep["stats"]
    assert st["device_unknown"] == 0
    assert st["known_checks"] <= 12 and st["canonical_keys"] <= 16       # the reference: 7 simplify + 1 pipeline per row
    print(f"report: {len(rep['classes'])} classes from {rep['novel_rows']} rows in {wall:.1f} s "
          f"(reference: {gold['wall_s']} s); SymPy calls {st['known_checks']} + {st['canonical_keys']} "
          f"instead of {st['known_checks_reference']} + {st['canonical_keys_reference']}")


def test_heavy_validation_from_db_kerr(cuda_device, tmp_path):
    """scope 'all' on the 306 rows of the reference's own Kerr depth-2 run: same heavy_is_valid column as the
    validator alone (nothing solves the Kerr surrogate, SURVEY 8f: every row fails at the residual), with SymPy
    called only for the rows the device cannot decide."""
    from pde_engine_b200.problems import load_problem
    from pde_engine_b200.report import heavy_validate_from_db
    gold = json.load(open(os.path.join(GOLDEN, "run_kerr_magnetosphere_d2.json")))
    rows = gold["rows"][:120]
    db = str(tmp_path / "kerr.db")
    _make_db(db, "expressions_k", rows)
    spec = load_problem("kerr_magnetosphere", cpu_validator=osym.KerrSymbolicValidator(), P=1024)
    st = heavy_validate_from_db(db, "expressions_k", spec, scope="all", out=lambda s: None)
    con = sqlite3.connect(db)
    got = con.execute("SELECT id, expression, heavy_is_valid, heavy_reason, heavy_validated_at FROM expressions_k ORDER BY id").fetchall()
    assert len(got) == len(rows) and all(g[4] is not None for g in got)
    assert st["device_rejected"] > 0.8 * len(rows) and st["cpu_heavy"] < 0.2 * len(rows) and st["ok"] == 0
    # the validator alone (the reference's heavy path, GM:2087-2099) on the same rows
    cpu = osym.KerrSymbolicValidator()
    import sympy as sp
    locs = spec.sympify_locals()
    for (i, s, ok, reason, _), r in zip(got, rows):
        want, _ = cpu.validate(sp.sympify(s, locals=locs), check_regularity=True, fast_point_only=False, lean_first=True,
                               defer_heavy_checks=False, enforce_anchor=True)
        assert bool(ok) == bool(want) == bool(r["is_valid"]), (s, reason)
    # scope 'valid' selects nothing here (GM:2049): a no-op that still creates the columns
    st2 = heavy_validate_from_db(db, "expressions_k", spec, scope="valid", out=lambda s: None)
    assert st2["rows"] == 0

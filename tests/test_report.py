"""Report equivalence bucketing (GM:1918-2008) -- the restated reference path and the device-screened path against
the reference's OWN report on its committed run database (tests/golden/report_force_free_d2.json, made by
tests/golden/make_golden_report.py)."""
import json
import os
import sqlite3

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import jets as J
from oracle import parser as op


def _rows_and_spec():
    from pde_engine_b200.problems import load_problem
    fx = load_golden("ref_fixtures.json")["ff_run_db"]
    spec = load_problem("force_free", make_gpu=False)
    rows = [(r["id"], r["expression"]) for r in fx if r["is_valid"] == 1]
    return fx, spec, rows


class OracleFingerprinter:
    """CPU stand-in with GpuFingerprinter's interface (values on the fingerprint grid by the oracle's interpreter):
    lets the CPU suite exercise the screening logic of report.novel_equivalence_classes."""
    keep_values = True

    def fingerprint(self, strs):
        from pde_engine_b200.fingerprint import fingerprint_grid
        pts = np.ascontiguousarray(fingerprint_grid("force_free", 64).T)
        sess = op.Session.for_problem("force_free")
        V = np.full((len(strs), 64), np.nan)
        for i, s in enumerate(strs):
            c = op.compile_expr(s, sess)
            if c.flags:
                continue
            with np.errstate(all="ignore"):
                V[i] = J.evaluate(c.whole(), pts, 0, sess.const_vals, sess.pow_vals)[0]

        class R:
            values = V
        return R


def test_reference_path_reproduces_the_reference_report():
    from pde_engine_b200.report import novel_equivalence_classes
    gold = json.load(open(os.path.join(GOLDEN, "report_force_free_d2.json")))
    _, spec, rows = _rows_and_spec()
    assert len(rows) == gold["novel_rows"] == 62
    classes, stats = novel_equivalence_classes(rows, spec.sympify_locals(), spec.known_solutions, None)
    assert classes == gold["classes"]
    assert stats["canonical_keys"] == stats["canonical_keys_reference"] and stats["known_checks"] > 300


def test_screened_path_same_report_fewer_sympy_calls():
    from pde_engine_b200.report import novel_equivalence_classes
    gold = json.load(open(os.path.join(GOLDEN, "report_force_free_d2.json")))
    _, spec, rows = _rows_and_spec()
    classes, stats = novel_equivalence_classes(rows, spec.sympify_locals(), spec.known_solutions, OracleFingerprinter())
    assert classes == gold["classes"] and len(classes) == gold["n_classes"] == 54
    # only rows that share their values with a known solution / another row reach SymPy
    assert stats["known_checks"] <= 12 and stats["known_checks_reference"] >= 7 * 54
    assert stats["canonical_keys"] <= 16 and stats["canonical_keys_reference"] >= 54


def test_generate_report_from_db_prints_the_reference_lines(tmp_path):
    from pde_engine_b200.engine import SCHEMA
    from pde_engine_b200.report import generate_report_from_db
    gold = json.load(open(os.path.join(GOLDEN, "report_force_free_d2.json")))
    fx, spec, _ = _rows_and_spec()
    db = str(tmp_path / "run.db")
    con = sqlite3.connect(db)
    con.execute(SCHEMA.format(table="expressions_t"))
    for r in fx:
        con.execute("INSERT INTO expressions_t (id, expression, normalized, signature, depth, validation_status, is_valid, validation_reason)"
                    " VALUES (?,?,?,?,?,?,?,?)", (r["id"], r["expression"], r["normalized"], r["signature"], r["depth"], r["status"], r["is_valid"], r["reason"]))
    con.commit()
    con.close()
    lines = []
    rep = generate_report_from_db(db, "expressions_t", spec, OracleFingerprinter(), out=lines.append)
    assert (rep["total"], rep["valid"], rep["paper_distinct"], rep["novel_rows"]) == (gold["total"], gold["valid"], gold["paper_distinct"], gold["novel_rows"])
    assert rep["classes"] == gold["classes"]
    assert "Novel equivalence classes: 54" in lines and "  • id=7 size=2 expr=inv(rho)" in lines

#!/usr/bin/env python3
"""Extract the reference's COMMITTED fixtures for this path into small JSON files
(build container only; read-only access to /root/reference).

  * problems/force_free/outputs/parallel_runs_paper_repro_20250908_052653_2a9752f9.db
      112 rows (id, depth, expression, normalized, signature, is_valid, reason)
      -- reproduced bit-exactly by the current reference code;
      validity columns trustworthy for ids 1-85.
  * lean_normalizer/physics_expressions.db : candidate string -> normalised string
  * problems/force_free/outputs/validator_cache.db : str(u) -> verdict
  * problems/kerr_magnetosphere/outputs/parallel_runs_paper_repro_20250908_044710_566b0ea9.db
"""
import json
import os
import sqlite3

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def ro(path):
    return sqlite3.connect(f"file:{path}?mode=ro&immutable=1", uri=True)


def run_db(path):
    c = ro(path)
    tb = [r[0] for r in c.execute("select name from sqlite_master where type='table'") if r[0].startswith("expressions_")][0]
    rows = c.execute(f"select id, depth, expression, normalized, signature, is_valid, validation_reason, validation_status from {tb} order by id").fetchall()
    return [dict(id=r[0], depth=r[1], expression=r[2], normalized=r[3], signature=r[4], is_valid=r[5], reason=r[6], status=r[7]) for r in rows]


def main():
    out = {}
    out["ff_run_db"] = run_db(os.path.join(REF, "problems/force_free/outputs/parallel_runs_paper_repro_20250908_052653_2a9752f9.db"))
    out["kerr_run_db"] = run_db(os.path.join(REF, "problems/kerr_magnetosphere/outputs/parallel_runs_paper_repro_20250908_044710_566b0ea9.db"))
    c = ro(os.path.join(REF, "lean_normalizer/physics_expressions.db"))
    out["normalizer_cache"] = [list(r) for r in c.execute("select expr_str, normalized from normalized_cache order by rowid")]
    c = ro(os.path.join(REF, "problems/force_free/outputs/validator_cache.db"))
    out["ff_validator_cache"] = [list(r) for r in c.execute("select expr_str, is_valid, reason from validation_cache order by rowid")]
    rep = json.load(open(os.path.join(REF, "problems/force_free/outputs/reproduction_20250815_162643.json")))
    out["ff_sequential_report_valid"] = rep["valid_solutions"]   # the 6 known solutions that validate (+ 4 primitives)
    path = os.path.join(OUT, "ref_fixtures.json")
    json.dump(out, open(path, "w"), indent=0)
    print("wrote", path, os.path.getsize(path), {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Run the UNMODIFIED reference CLI in the build container and freeze its run database
(build container only: needs /root/reference; the copy it runs from is baseline/_ref, tools/refcopy.py).

    python general_method_paper_reproduction.py --problem kerr_magnetosphere --max-depth 2 --validators 0

(BASELINE configs: the as-shipped path, fresh caches, SURVEY 6.2: 306 rows, 0 valid, ~50 s.)  The rows
(id, depth, expression, normalized, signature, status, is_valid, reason, paper-solution columns) go to
tests/golden/run_<problem>_d<depth>.json; tests/test_gpu_dropin.py diffs the run database the SAME CLI writes
with the GPU path installed against them.

Force-free depth 2 is pinned by the reference's own committed run database (ref_fixtures.json: ff_run_db); a fresh
run of it needs > 30 min of SymPy here (single rows take 6+ min) and is optional: --problem force_free.
"""
import argparse
import json
import os
import sqlite3
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from tools.refcopy import GM, ensure_baseline_ref, patched_copy   # noqa: E402


def export_run_db(out_dir):
    dbs = sorted(f for f in os.listdir(out_dir) if f.startswith("parallel_runs_") and f.endswith(".db"))
    assert len(dbs) == 1, dbs
    c = sqlite3.connect(os.path.join(out_dir, dbs[0]))
    tb = [r[0] for r in c.execute("select name from sqlite_master where type='table'") if r[0].startswith("expressions_")][0]
    cols = ["id", "depth", "expression", "normalized", "signature", "validation_status", "is_valid", "validation_reason",
            "validator_method", "is_paper_solution", "paper_solution_name"]
    rows = c.execute(f"select {', '.join(cols)} from {tb} order by id").fetchall()
    return [dict(zip(cols, r)) for r in rows]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--problem", default="kerr_magnetosphere")
    ap.add_argument("--max-depth", type=int, default=2)
    ap.add_argument("--timeout", type=float, default=3600)
    a = ap.parse_args()
    assert ensure_baseline_ref()
    with tempfile.TemporaryDirectory() as tmp:
        dst = os.path.join(tmp, "ref")
        diff = patched_copy(dst)
        assert diff == "", "the golden run uses the reference as shipped"
        t0 = time.time()
        subprocess.run([sys.executable, GM, "--problem", a.problem, "--max-depth", str(a.max_depth), "--validators", "0"],
                       cwd=dst, check=True, timeout=a.timeout, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT,
                       start_new_session=True)
        wall = time.time() - t0
        rows = export_run_db(os.path.join(dst, "problems", a.problem, "outputs"))
    path = os.path.join(HERE, f"run_{a.problem}_d{a.max_depth}.json")
    json.dump(dict(command=f"python {GM} --problem {a.problem} --max-depth {a.max_depth} --validators 0",
                   wall_s=round(wall, 1), cores=1, n_rows=len(rows), rows=rows), open(path, "w"), indent=0)
    print("wrote", path, len(rows), "rows", f"{wall:.1f} s")


if __name__ == "__main__":
    main()

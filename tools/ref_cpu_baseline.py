#!/usr/bin/env python3
"""The reference's OWN CPU validator pool, timed on this box (north_star: "reported next to the reference's CPU
validator run with --validators equal to the host core count on the same box in the same run").

    python general_method_paper_reproduction.py --problem P --max-depth D --validators C        (C = os.cpu_count())

run from a scratch copy of baseline/_ref (the unmodified reference code, tools/refcopy.py) with FRESH caches and the
one-line import repair of the validator worker (GM:1694, `from physics_agent.problems import load_problem` names a
package that does not exist; without it every worker dies at GM:1701) -- the unified diff of the repair is
part of the result.  A full depth-3 run needs hours (single rows take minutes of SymPy), so the run is BOUNDED: after
`wall_s` seconds the whole process group is killed and the run database is read for what the pool finished:

    rows_per_s       = rows with validation_status 'completed' / wall
    evals_per_s      = rows_per_s x points per row (force-free validates at 1 point, FFV:296-297; Kerr at 3, KV:163-192)

This is a reported baseline (bench.py `cpu_baseline_reference`), never a product path: nothing here is imported by
pde_engine_b200.
"""
from __future__ import annotations

import os
import signal
import sqlite3
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from tools.refcopy import BASELINE_REF, GM, patched_copy   # noqa: E402

POINTS_PER_ROW = {"force_free": 1, "kerr_magnetosphere": 3}


def available() -> bool:
    return os.path.exists(os.path.join(BASELINE_REF, GM))


def _read_run_db(out_dir: str) -> dict:
    dbs = sorted(f for f in os.listdir(out_dir) if f.startswith("parallel_runs_") and f.endswith(".db"))
    if not dbs:
        return dict(rows=0, completed=0, valid=0, by_depth={})
    con = sqlite3.connect(f"file:{os.path.join(out_dir, dbs[-1])}?mode=ro", uri=True, timeout=30)
    tb = [r[0] for r in con.execute("select name from sqlite_master where type='table'") if r[0].startswith("expressions_")][0]
    rows = con.execute(f"select depth, validation_status, is_valid from {tb}").fetchall()
    by_depth = {}
    for d, st, _ in rows:
        e = by_depth.setdefault(int(d), dict(rows=0, completed=0))
        e["rows"] += 1
        e["completed"] += st == "completed"
    return dict(rows=len(rows), completed=sum(st == "completed" for _, st, _ in rows),
                valid=sum(bool(v) for _, st, v in rows if st == "completed"), by_depth=by_depth)


def run_reference_validators(problem: str = "force_free", max_depth: int = 3, validators: int | None = None,
                             wall_s: float = 60.0) -> dict:
    """Bounded run of the reference CLI with its validator pool; see the module docstring."""
    validators = validators or (os.cpu_count() or 1)
    if not available():
        raise FileNotFoundError(f"{BASELINE_REF} is missing (made by tools/refcopy.py in the build container)")
    with tempfile.TemporaryDirectory() as tmp:
        dst = os.path.join(tmp, "ref")
        diff = patched_copy(dst, repair_workers=True)
        out_dir = os.path.join(dst, "problems", problem, "outputs")
        log = open(os.path.join(tmp, "run.log"), "w")
        t0 = time.time()
        proc = subprocess.Popen([sys.executable, GM, "--problem", problem, "--max-depth", str(max_depth), "--validators", str(validators)],
                                cwd=dst, stdout=log, stderr=subprocess.STDOUT, start_new_session=True)
        finished = False
        try:
            proc.wait(timeout=wall_s)
            finished = True
        except subprocess.TimeoutExpired:
            pass
        wall = time.time() - t0
        stats = _read_run_db(out_dir)            # what the pool finished within the window
        if not finished:
            try:
                os.killpg(proc.pid, signal.SIGKILL)      # exactly the process group this call started
            except ProcessLookupError:
                pass
            proc.wait()
        log.close()
        text = open(os.path.join(tmp, "run.log")).read()
        started = text.count("Validator process started")
        # the reference's own monitor line: "Status (running): generated G, validated V/G (val r/s, gen r/s)" (GM:826-900).
        # The run database is in WAL mode and is being written when the window closes: a second connection sometimes
        # sees none of the UPDATEs yet, so the reference's own counter is taken when it is ahead of the database read.
        import re
        mon = [int(m.group(1)) for m in re.finditer(r"validated (\d+)/\d+", text)]
        monitor_validated = max(mon) if mon else 0
        db_completed = stats["completed"]
        if monitor_validated > stats["completed"]:
            stats["completed"] = monitor_validated
        log_tail = text[-600:] if stats["completed"] == 0 else ""
    pts = POINTS_PER_ROW[problem]
    rps = stats["completed"] / wall if wall > 0 else 0.0
    return dict(kind="reference", command=f"python {GM} --problem {problem} --max-depth {max_depth} --validators {validators}",
                validators=validators, validator_processes_started=started, cores=os.cpu_count(), wall_s=round(wall, 2),
                run_finished=finished, rows_inserted=stats["rows"], rows_validated=stats["completed"], rows_valid=stats["valid"],
                rows_validated_db=db_completed, rows_validated_monitor=monitor_validated,
                by_depth=stats["by_depth"], rows_per_s=rps, points_per_row=pts, value=rps * pts, unit="evals/s",
                patch=diff, log_tail=log_tail, note="bounded window: the process group is killed after wall_s; rows the pool completed are read from the run database")


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--problem", default="force_free")
    ap.add_argument("--max-depth", type=int, default=3)
    ap.add_argument("--validators", type=int, default=0)
    ap.add_argument("--wall", type=float, default=60.0)
    a = ap.parse_args()
    print(json.dumps(run_reference_validators(a.problem, a.max_depth, a.validators or None, a.wall), indent=1))

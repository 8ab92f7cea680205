"""The real thing: the reference's OWN CLI (`general_method_paper_reproduction.py --problem ... --max-depth 2
--validators 0`), unmodified except for the two added lines of INTEGRATION.md 2 after GM:1243, run in a subprocess
on the GPU box from a scratch copy of baseline/_ref (the reference's code, git-ignored, shipped with the snapshot;
tools/refcopy.py).  The run database it writes -- the reference's writer, schema and report code, untouched -- is
diffed against

  * force-free: the reference's committed 112-row run database (tests/golden/ref_fixtures.json: ff_run_db),
  * Kerr: the rows of the SAME command run here with the reference as shipped (tests/golden/run_kerr_magnetosphere_d2.json,
    made by tests/golden/make_golden_runs.py).

`/root/reference` is never read at run time.
"""
import json
import os
import sqlite3
import subprocess
import sys
import time

import pytest

from conftest import GOLDEN, REPO, load_golden

pytestmark = pytest.mark.gpu

sys.path.insert(0, REPO)
from tools import refcopy   # noqa: E402


def _run_cli(problem, max_depth, dst, timeout=1500):
    diff = refcopy.patched_copy(dst, install_gpu=True)
    added = [l for l in diff.splitlines() if l.startswith("+") and not l.startswith("+++")]
    assert [a.strip("+ ").split("(")[0] for a in added] == ["from pde_engine_b200.engine import install", "install"], diff
    env = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get("PYTHONPATH", ""))
    t0 = time.time()
    p = subprocess.run([sys.executable, refcopy.GM, "--problem", problem, "--max-depth", str(max_depth), "--validators", "0"],
                       cwd=dst, env=env, timeout=timeout, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       start_new_session=True)
    wall = time.time() - t0
    assert p.returncode == 0, p.stdout[-4000:]
    out_dir = os.path.join(dst, "problems", problem, "outputs")
    dbs = [f for f in os.listdir(out_dir) if f.startswith("parallel_runs_") and f.endswith(".db")]
    assert len(dbs) == 1
    con = sqlite3.connect(os.path.join(out_dir, dbs[0]))
    tb = [r[0] for r in con.execute("select name from sqlite_master where type='table'") if r[0].startswith("expressions_")][0]
    cols = ["id", "depth", "expression", "normalized", "signature", "validation_status", "is_valid", "validation_reason",
            "validator_method", "is_paper_solution", "paper_solution_name", "validator_evidence"]
    rows = [dict(zip(cols, r)) for r in con.execute(f"select {', '.join(cols)} from {tb} order by id")]
    return rows, wall, p.stdout


def _need_ref():
    if not os.path.exists(os.path.join(refcopy.BASELINE_REF, refcopy.GM)):
        pytest.skip("baseline/_ref (the reference's code) is not in this snapshot: run tools/refcopy.py in the build container")


def test_reference_cli_force_free_depth2(cuda_device, tmp_path):
    """BASELINE configs[0] through the reference's own engine with the GPU generator + filter installed."""
    _need_ref()
    rows, wall, log = _run_cli("force_free", 2, str(tmp_path / "ref"))
    fx = load_golden("ref_fixtures.json")["ff_run_db"]
    assert len(rows) == len(fx) == 112, (len(rows), log[-2000:])
    n_gpu_rejected = 0
    for got, want in zip(rows, fx):
        for k in ("id", "depth", "expression", "normalized", "signature"):
            assert got[k] == want[k], (k, got, want)
        assert got["validation_status"] == "completed", got
        # committed verdicts: trustworthy for ids 1-85 and for the valid cache hits among 86-112
        if want["id"] <= 85 or want["is_valid"] == 1:
            assert bool(got["is_valid"]) == bool(want["is_valid"]), (got, want["reason"])
            if want["is_valid"] or want["reason"] in ("constant-only (skipped)", "Zero gradient (constant expression)"):
                assert got["validation_reason"] == want["reason"], (got, want["reason"])
        ev = json.loads(got["validator_evidence"] or "{}")
        if str(ev.get("gpu_filter", "")).startswith("pde_engine_b200") and not got["is_valid"]:      # decided on the device
            n_gpu_rejected += 1
            assert got["validation_reason"].startswith("Invalid ("), got                 # FFV:395 wording
            assert "Error" not in got["validation_reason"] and "Could not" not in got["validation_reason"]   # GM:218
    assert n_gpu_rejected >= 20, n_gpu_rejected          # the slow non-solutions never reached SymPy
    assert sum(bool(r["is_valid"]) for r in rows) >= 62
    # (is_paper_solution stays 0 on this path: the reference's inline validation never fills it, only its
    #  worker pool does, GM:1785-1798 -- the committed run database has 0 there as well)
    assert not any(r["is_paper_solution"] for r in rows)
    print(f"reference CLI + GPU path, force_free depth 2: {len(rows)} rows in {wall:.1f} s ({n_gpu_rejected} rejected on the device)")


def test_reference_cli_kerr_depth2(cuda_device, tmp_path):
    """`--problem kerr_magnetosphere --max-depth 2`: same 306 rows and verdicts as the reference as shipped."""
    _need_ref()
    gold = json.load(open(os.path.join(GOLDEN, "run_kerr_magnetosphere_d2.json")))
    rows, wall, log = _run_cli("kerr_magnetosphere", 2, str(tmp_path / "ref"))
    assert len(rows) == gold["n_rows"] == 306, (len(rows), log[-2000:])
    for got, want in zip(rows, gold["rows"]):
        for k in ("id", "depth", "expression", "normalized", "signature", "validation_status", "is_paper_solution"):
            assert got[k] == want[k], (k, got, want)
        assert bool(got["is_valid"]) == bool(want["is_valid"]), (got, want)
        if want["validation_reason"] == "constant-only (skipped)":
            assert got["validation_reason"] == want["validation_reason"]
        else:
            assert got["validation_reason"].startswith("PDE residual != 0"), got          # KV:269 wording
    print(f"reference CLI + GPU path, kerr depth 2: {len(rows)} rows in {wall:.1f} s (reference as shipped: {gold['wall_s']} s on 1 core)")

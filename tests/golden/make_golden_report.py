#!/usr/bin/env python3
"""The reference's OWN report on its committed force-free run database (build container only).

    python general_method_paper_reproduction.py --problem force_free --print-run-id paper_repro_20250908_052653_2a9752f9

run from a scratch copy of the unmodified reference (tools/refcopy.py) next to a copy of the committed database
problems/force_free/outputs/parallel_runs_paper_repro_20250908_052653_2a9752f9.db (the 112 rows of
tests/golden/ref_fixtures.json: ff_run_db).  The "novel solutions, deduplicated by mathematical equivalence" block
(_generate_report_from_db, GM:1826-2020) is parsed into tests/golden/report_force_free_d2.json:
novel row count, class count and the printed classes (rep id, size, representative) in print order.
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
from tools.refcopy import GM, patched_copy   # noqa: E402

RUN_ID = "paper_repro_20250908_052653_2a9752f9"
DB = f"/root/reference/problems/force_free/outputs/parallel_runs_{RUN_ID}.db"


def main():
    with tempfile.TemporaryDirectory() as tmp:
        dst = os.path.join(tmp, "ref")
        assert patched_copy(dst) == ""
        shutil.copy(DB, os.path.join(dst, "problems", "force_free", "outputs", os.path.basename(DB)))
        t0 = time.time()
        p = subprocess.run([sys.executable, GM, "--problem", "force_free", "--print-run-id", RUN_ID], cwd=dst,
                           capture_output=True, text=True, timeout=3600, check=True)
        wall = time.time() - t0
    out = p.stdout
    novel = int(re.search(r"Novel valid rows \(non-paper\): (\d+)", out).group(1))
    ncls = int(re.search(r"Novel equivalence classes: (\d+)", out).group(1))
    classes = [dict(rep_id=int(m.group(1)), size=int(m.group(2)), rep_str=m.group(3))
               for m in re.finditer(r"^  • id=(\d+) size=(\d+) expr=(.*)$", out, re.M)]
    assert len(classes) == ncls
    head = {k: int(re.search(pat, out).group(1)) for k, pat in
            (("total", r"Total expressions generated: (\d+)"), ("valid", r"Valid foliations found: (\d+)"),
             ("paper_distinct", r"Known solutions found: (\d+)"))}
    path = os.path.join(HERE, "report_force_free_d2.json")
    json.dump(dict(command=f"python {GM} --problem force_free --print-run-id {RUN_ID}", wall_s=round(wall, 1), **head,
                   novel_rows=novel, n_classes=ncls, classes=classes), open(path, "w"), indent=0)
    print("wrote", path, head, novel, ncls, f"{wall:.1f} s")


if __name__ == "__main__":
    main()

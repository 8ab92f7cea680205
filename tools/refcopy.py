#!/usr/bin/env python3
"""Working copies of the UNMODIFIED reference for the drop-in test and the CPU baseline.

`/root/reference` exists only in the build container.  `ensure_baseline_ref()` copies its CODE (no committed
run databases / caches / images: stale caches change run time and verdict text) into the git-ignored
`baseline/_ref/`, which travels to the GPU box with the snapshot.  Nothing of it is tracked in this repo.

`patched_copy(dst, ...)` makes a scratch copy of `baseline/_ref` and applies, to the COPY only:

  install_gpu     the 2 added lines of INTEGRATION.md 2 after GM:1243 (the generator process rebuilds its
                  discovery object there; CUDA is initialised inside that process)
  repair_workers  the import repair of the validator worker (GM:1694: `from physics_agent.problems import`
                  names a package that does not exist, so every `--validators N` worker dies with a NameError
                  at GM:1701) -- needed only for the reference's own CPU baseline
and returns the unified diff of what it changed, so every result can name the patch it ran with.
"""
from __future__ import annotations

import difflib
import os
import shutil

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASELINE_REF = os.path.join(REPO, "baseline", "_ref")
REFERENCE = "/root/reference"
GM = "general_method_paper_reproduction.py"


def ensure_baseline_ref(force: bool = False) -> bool:
    """Create baseline/_ref from /root/reference (code only).  True if it exists afterwards."""
    if os.path.exists(os.path.join(BASELINE_REF, GM)) and not force:
        return True
    if not os.path.exists(os.path.join(REFERENCE, GM)):
        return False
    if os.path.exists(BASELINE_REF):
        shutil.rmtree(BASELINE_REF)

    def ignore(d, names):
        return [n for n in names if n.endswith((".db", ".db-shm", ".db-wal", ".png", ".pyc", ".json", ".txt")) or n == "__pycache__"]

    os.makedirs(BASELINE_REF)
    for name in (GM, "expression_operations.py", "LICENSE"):
        shutil.copy(os.path.join(REFERENCE, name), os.path.join(BASELINE_REF, name))
    for d in ("lean_normalizer", "problems"):
        shutil.copytree(os.path.join(REFERENCE, d), os.path.join(BASELINE_REF, d), ignore=ignore)
    for slug in ("force_free", "kerr_magnetosphere"):
        os.makedirs(os.path.join(BASELINE_REF, "problems", slug, "outputs"), exist_ok=True)      # the engine writes its run DBs here
    for root, dirs, files in os.walk(BASELINE_REF):
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755 if n in dirs or n == GM else 0o644)
    return True


INSTALL_ANCHOR = "discovery = GeneralFoliationDiscovery(use_lean_normalizer=True, problem_name=problem_name)"
WORKER_BAD_IMPORT = "from physics_agent.problems import load_problem"


def patched_copy(dst: str, install_gpu: bool = False, repair_workers: bool = False, P: int = 4096) -> str:
    """Scratch copy of baseline/_ref at `dst` with the requested patches; returns their unified diff."""
    if not os.path.exists(os.path.join(BASELINE_REF, GM)):
        raise FileNotFoundError(f"{BASELINE_REF} is missing: run tools/refcopy.py in the build container")
    if os.path.exists(dst):
        shutil.rmtree(dst)
    shutil.copytree(BASELINE_REF, dst)
    path = os.path.join(dst, GM)
    old = open(path).read().splitlines(keepends=True)
    new = []
    n_install = n_repair = 0
    for line in old:
        if repair_workers and WORKER_BAD_IMPORT in line:
            line = line.replace(WORKER_BAD_IMPORT, "from problems import load_problem")
            n_repair += 1
        new.append(line)
        if install_gpu and line.strip() == INSTALL_ANCHOR and "_parallel_generator_worker" in "".join(old[max(0, len(new) - 30):len(new)]):
            pad = line[:len(line) - len(line.lstrip())]
            new.append(f"{pad}from pde_engine_b200.engine import install\n")
            new.append(f"{pad}install(discovery, P={P})\n")
            n_install += 1
    if install_gpu and n_install != 1:
        raise RuntimeError(f"GM:1243 anchor matched {n_install} times")
    if repair_workers and n_repair != 1:
        raise RuntimeError(f"GM:1694 anchor matched {n_repair} times")
    open(path, "w").writelines(new)
    return "".join(difflib.unified_diff(old, new, f"a/{GM}", f"b/{GM}", n=1))


if __name__ == "__main__":
    ok = ensure_baseline_ref(force=True)
    print("baseline/_ref", "ready" if ok else "NOT available (no /root/reference here)")

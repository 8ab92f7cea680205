#!/usr/bin/env python3
"""Reference verdicts (is_valid, reason) for force-free / Kerr candidates.

Build-container only.  Calls the UNMODIFIED reference validator exactly as the
engine does (general_method_paper_reproduction.py:1302-1316: sympify with the
discovery locals, then validate(u, check_regularity=False, fast_point_only=False,
lean_first=True, defer_heavy_checks=True, enforce_anchor=False) with the
TypeError fallback to the 2-kwarg form), one candidate per task, fresh caches,
hard per-candidate wall cap (the symbolic zero test takes 0.01 s - >20 min,
SURVEY 0.5): candidates that hit the cap are recorded as {"timeout": true}.

Usage: python tests/golden/make_golden_verdicts.py force_free 3 15 120 [workers] [wall budget in seconds]
       (problem, depth, take every k-th unique, cap seconds)

With a wall budget the run stops scheduling when it is spent and writes what is finished; the fixture always holds
EVERY record made so far with this cap (union over all strides used, in enumeration order).

Resumable: every finished record is appended to $PDE_REF_WORK/verdicts_<problem>_d<depth>.jsonl and
records already present there (or in the committed fixture, when it was made with the same cap) are
not recomputed -- the full depth-3 force-free set is ~30 CPU-hours.
"""
from __future__ import annotations

import gzip
import json
import multiprocessing as mp
import os
import signal
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
WORK = os.environ.get("PDE_REF_WORK", "/tmp/pde_ref_work")
_st = {}


def _init(problem, cap):
    import io, contextlib
    sys.path.insert(0, WORK)
    os.chdir(WORK)
    import sympy as sp
    with contextlib.redirect_stdout(io.StringIO()):
        from problems import load_problem
        from expression_operations import UNARY_OPS
        spec = load_problem(problem)
    # fresh, per-process validator cache (the reference caches verdicts by str(u))
    if hasattr(spec.validator, "_check_cache"):
        spec.validator._check_cache = lambda h: None
        spec.validator._save_to_cache = lambda *a, **k: None
    locs = {}
    locs.update(spec.symbols)
    locs.update(spec.constants)
    locs.update(UNARY_OPS)
    _st.update(sp=sp, spec=spec, locs=locs, cap=cap)


def _alarm(*_):
    raise TimeoutError()


def _one(s):
    sp, spec = _st["sp"], _st["spec"]
    signal.signal(signal.SIGALRM, _alarm)
    signal.alarm(_st["cap"])
    t0 = time.time()
    try:
        u = sp.sympify(s, locals=_st["locs"])
        syms = list(spec.symbols.values())
        if not any(u.has(v) for v in syms):
            return {"s": s, "is_valid": False, "reason": "constant-only (skipped)", "t": 0.0}   # GM:1293-1294
        try:
            ok, reason = spec.validator.validate(u, check_regularity=False, fast_point_only=False,
                                                 lean_first=True, defer_heavy_checks=True, enforce_anchor=False)
        except TypeError:
            ok, reason = spec.validator.validate(u, check_regularity=False, fast_point_only=False)
        return {"s": s, "is_valid": bool(ok), "reason": reason[:160], "t": round(time.time() - t0, 2)}
    except TimeoutError:
        return {"s": s, "timeout": True, "t": round(time.time() - t0, 2)}
    except Exception as e:  # noqa
        return {"s": s, "error": repr(e)[:160]}
    finally:
        signal.alarm(0)


def main():
    problem, depth, step, cap = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    workers = int(sys.argv[5]) if len(sys.argv) > 5 else min(8, os.cpu_count() or 1)
    budget = float(sys.argv[6]) if len(sys.argv) > 6 else float("inf")
    enum_file = {"force_free": "enum_force_free_d4", "kerr_magnetosphere": "enum_kerr_magnetosphere_d3"}[problem]
    g = json.load(gzip.open(os.path.join(REPO, "tests", "golden", enum_file + ".json.gz"), "rt"))
    uniques = g["depths"][str(depth)]["uniques"]
    exprs = uniques[::step]
    print(len(exprs), "expressions", flush=True)
    path = os.path.join(REPO, "tests", "golden", f"verdicts_{problem}_d{depth}.json")
    log_path = os.path.join(WORK, f"verdicts_{problem}_d{depth}.jsonl")
    done = {}
    if os.path.exists(path):
        old = json.load(open(path))
        if old.get("cap_seconds") == cap:
            done.update({r["s"]: r for r in old["records"]})
    if os.path.exists(log_path):
        for line in open(log_path):
            r = json.loads(line)
            done[r["s"]] = r
    todo = [s for s in exprs if s not in done]
    print(len(done), "already done,", len(todo), "to do,", workers, "workers", flush=True)
    t0 = time.time()
    if budget <= 0:
        todo = []                 # wall budget 0: only fold the records of the work log into the fixture
    with mp.Pool(workers, initializer=_init, initargs=(problem, cap), maxtasksperchild=20) as pool, open(log_path, "a") as log:
        k = 0
        for r in pool.imap_unordered(_one, todo, chunksize=1):
            done[r["s"]] = r
            log.write(json.dumps(r) + "\n")
            log.flush()
            k += 1
            if k % 50 == 0:
                print(k, round(time.time() - t0), flush=True)
            if time.time() - t0 > budget:
                print("wall budget spent after", k, "records", flush=True)
                pool.terminate()
                break
    # the reference's validator catches the alarm inside validate() (FFV:434-437: `except Exception as e: "Error: {e}"`),
    # so a candidate that hit the wall cap comes back as (False, "Error: ") after `cap` seconds: record it as a timeout
    for r in done.values():
        if r.get("reason", "").strip() == "Error:" and r.get("t", 0) >= cap - 1:
            r.pop("is_valid", None)
            r.pop("reason", None)
            r["timeout"] = True
    recs = [done[s] for s in uniques if s in done]
    strides = sorted(set([step] + [int(x) for x in str(old.get("step", "")).split("+") if x.isdigit()])) if os.path.exists(path) else [step]
    out = {"problem": problem, "depth": depth, "step": "+".join(str(x) for x in strides), "cap_seconds": cap,
           "n_uniques": len(uniques), "records": recs}
    json.dump(out, open(path, "w"), indent=0)
    nv = sum(1 for r in recs if r.get("is_valid"))
    print("wrote", path, "valid", nv, "invalid", sum(1 for r in recs if r.get("is_valid") is False),
          "timeout", sum(1 for r in recs if r.get("timeout")), "error", sum(1 for r in recs if "error" in r))


if __name__ == "__main__":
    main()

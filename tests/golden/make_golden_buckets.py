#!/usr/bin/env python3
"""Exact function buckets of the depth <= D force-free / Kerr uniques (ground truth for the device fingerprints).

Build-container only: takes the problem's symbols, constants and operator table from the UNMODIFIED reference
(/root/reference: problems.load_problem, expression_operations.UNARY_OPS) so that every string is parsed exactly as
``emit_to_db`` parses the argument of ``validate`` (general_method_paper_reproduction.py:1257), evaluates it with
40-digit arithmetic at 6 generic points (oracle/fingerprint.py) and records, per unique string, the index of the
first string with the same values.

Usage: python tests/golden/make_golden_buckets.py force_free 3
"""
from __future__ import annotations

import gzip
import hashlib
import json
import multiprocessing as mp
import os
import sys
import tempfile

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("PDE_REFERENCE", "/root/reference")
_st = {}


def _init(problem, pts):
    import contextlib, io
    sys.path.insert(0, REPO)
    sys.path.insert(0, REF)
    work = tempfile.mkdtemp(prefix="pde_buckets_")          # the reference writes its caches relative to the cwd
    for slug in ("force_free", "kerr_magnetosphere"):
        os.makedirs(os.path.join(work, "problems", slug, "outputs"), exist_ok=True)
    os.chdir(work)
    import numpy as np
    with contextlib.redirect_stdout(io.StringIO()):
        from problems import load_problem
        from expression_operations import UNARY_OPS
        spec = load_problem(problem)
    from oracle import fingerprint as ofp
    locs = {}
    locs.update(spec.symbols)
    locs.update(spec.constants)
    locs.update(UNARY_OPS)
    extra = {}
    if problem != "force_free":
        import sympy as sp
        # generic parameter values (pde_engine_b200/fingerprint.py:GENERIC_CONSTS)
        extra = {spec.constants["M"]: sp.Rational(1.1378240173), spec.constants["a"]: sp.Rational(0.2718653942)}
    syms = list(spec.symbols.values())
    _st.update(ofp=ofp, locs=locs, points=ofp.exact_points(syms, np.asarray(pts), 6, extra))


def _one(s):
    try:
        return _st["ofp"].exact_signature(s, _st["locs"], _st["points"])
    except Exception:
        return None


def main():
    problem, depth = sys.argv[1], int(sys.argv[2])
    sys.path.insert(0, REPO)
    from oracle import fingerprint as ofp
    from oracle.residuals import splitmix64_stream
    enum_depth = {"force_free": 4, "kerr_magnetosphere": 3}[problem]
    with gzip.open(os.path.join(REPO, "tests", "golden", f"enum_{problem}_d{enum_depth}.json.gz"), "rt") as f:
        enum = json.load(f)
    strs = []
    for d in range(1, depth + 1):
        strs += enum["depths"][str(d)]["uniques"]
    # 6 generic points, both signs of the second coordinate (any generic points give the same partition)
    g = splitmix64_stream(0xB0C4E75)
    u = [(next(g) >> 11) / float(1 << 53) for _ in range(18)]
    if problem == "force_free":
        pts = [[0.25 + 1.75 * u[3 * k] for k in range(6)],
               [(0.25 + 1.75 * u[3 * k + 1]) * (1 if k % 2 else -1) for k in range(6)]]
    else:
        pts = [[2.2 + 3.8 * u[3 * k] for k in range(6)], [(-0.9 + 1.8 * u[3 * k + 1]) for k in range(6)]]
    with mp.Pool(os.cpu_count(), initializer=_init, initargs=(problem, pts)) as pool:
        sigs = pool.map(_one, strs, chunksize=16)
    buckets = ofp.partition(sigs)
    out = {"problem": problem, "depth": depth, "n": len(strs), "points": pts,
           "strings_sha256": hashlib.sha256("\n".join(strs).encode()).hexdigest(),
           "n_functions": len({b for b in buckets if b >= 0}), "n_unknown": sum(1 for b in buckets if b < 0),
           "bucket": buckets}
    path = os.path.join(REPO, "tests", "golden", f"function_buckets_{problem}_d{depth}.json.gz")
    with gzip.open(path, "wt") as f:
        json.dump(out, f)
    print(path, "rows", len(strs), "functions", out["n_functions"], "unknown", out["n_unknown"])


if __name__ == "__main__":
    main()

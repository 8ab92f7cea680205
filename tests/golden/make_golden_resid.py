#!/usr/bin/env python3
"""Per-point jet / residual golden vectors from the UNMODIFIED reference.

Build-container only (needs /root/reference, imported from a scratch copy).

force-free: the reference's ``PreciseFoliationValidator.validate`` is called
as the engine calls it (GM:1302-1316); the module-level ``det`` it uses
(FFV:347) is wrapped so the 2x2 matrix [[LT_A, LT_B],[L2T_A, L2T_B]] the
reference built (FFV:305-344) is captured, then the call is aborted (the
symbolic zero test is not needed for numeric vectors).  Values are
``complex(expr.subs(point).evalf(50))`` -- the reference's own recipe
(FFV:388, LBF:287).

Kerr: ``KerrMagnetosphereValidator._lhs(u)`` (KV:77-91) with M=1, a=1/10,
``float(N(val, 40))`` (KV:176-181).

u-jets: repeated ``sympy.diff`` + subs + evalf(50).

Points: the first NPTS points of ``oracle.residuals.collocation_grid`` (the
reference's rational test points first, then float64 grid points taken as
exact rationals).
"""
from __future__ import annotations

import gzip
import json
import multiprocessing as mp
import os
import signal
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
WORK = os.environ.get("PDE_REF_WORK", "/tmp/pde_ref_work")
NPTS = 8
TIMEOUT = 90

_state = {}


class _Captured(Exception):
    pass


def _init(problem):
    import io, contextlib
    sys.path.insert(0, WORK)
    sys.path.insert(0, REPO)
    os.chdir(WORK)
    import sympy as sp
    with contextlib.redirect_stdout(io.StringIO()):
        from problems import load_problem
        from expression_operations import UNARY_OPS
        spec = load_problem(problem)
    locs = {}
    locs.update(spec.symbols)
    locs.update(spec.constants)
    locs.update(UNARY_OPS)          # GM:85-93
    from oracle.residuals import collocation_grid
    pts = collocation_grid("force_free" if problem == "force_free" else "kerr", NPTS)
    v0, v1 = list(spec.symbols.values())
    rat_pts = []
    ref_rat = ([(sp.Rational(4, 5), sp.Rational(6, 7)), (sp.Rational(3, 4), sp.Rational(5, 6)), (sp.Rational(7, 8), sp.Rational(1, 2))]
               if problem == "force_free" else
               [(sp.Rational(5, 2), sp.Rational(3, 5)), (sp.Rational(7, 3), sp.Rational(1, 3)), (sp.Integer(5), sp.Rational(-2, 5))])
    for k in range(NPTS):
        if k < 3:
            rat_pts.append(ref_rat[k])
        else:
            rat_pts.append((sp.Rational(float(pts[k, 0])), sp.Rational(float(pts[k, 1]))))
    _state.update(sp=sp, spec=spec, locs=locs, v0=v0, v1=v1, rat_pts=rat_pts, problem=problem)
    if problem == "force_free":
        import problems.force_free.validator as ffv
        def fake_det(M):
            _state["M"] = M
            raise _Captured()
        ffv.det = fake_det


def _num(e, subs):
    sp = _state["sp"]
    try:
        v = complex(e.subs(subs).evalf(50))
    except Exception:
        return None
    if v != v or abs(v.imag) > 1e-30 * max(1.0, abs(v.real)) or abs(v) == float("inf"):
        return None
    return v.real


def _alarm(*_):
    raise TimeoutError()


def _one(s):
    sp = _state["sp"]
    spec = _state["spec"]
    v0, v1 = _state["v0"], _state["v1"]
    signal.signal(signal.SIGALRM, _alarm)
    signal.alarm(TIMEOUT)
    try:
        u = sp.sympify(s, locals=_state["locs"])
        order = 4 if _state["problem"] == "force_free" else 2
        consts = {}
        if _state["problem"] != "force_free":
            val = spec.validator
            consts = {val.M: val.M_value, val.a: val.a_value}
        ders = []
        for n in range(order + 1):
            for j in range(n + 1):
                e = u
                for _ in range(n - j):
                    e = sp.diff(e, v0)
                for _ in range(j):
                    e = sp.diff(e, v1)
                ders.append(e)
        if _state["problem"] == "force_free":
            _state.pop("M", None)
            # the call the engine makes (GM:1302-1316 falls back to the 2-kwarg form)
            spec.validator._check_cache = lambda h: None
            spec.validator._save_to_cache = lambda *a, **k: None
            spec.validator.validate(u, check_regularity=False, fast_point_only=False)
            M = _state.get("M")
            entries = None if M is None else [M[0, 0], M[0, 1], M[1, 0], M[1, 1]]
        else:
            entries = [spec.validator._lhs(u)]
        rec = {"s": s, "jets": [], "parts": [], "R": []}
        for (a, b) in _state["rat_pts"]:
            subs = {v0: a, v1: b, **consts}
            rec["jets"].append([_num(e, subs) for e in ders])
            if entries is None:
                rec["parts"].append(None)
                rec["R"].append(None)
                continue
            pv = [_num(e, subs) for e in entries]
            rec["parts"].append(pv)
            if _state["problem"] == "force_free":
                if any(p is None for p in pv):
                    rec["R"].append(None)
                else:
                    rec["R"].append(_num(entries[0] * entries[3] - entries[1] * entries[2], subs))
            else:
                rec["R"].append(pv[0])
        return rec
    except TimeoutError:
        return {"s": s, "timeout": True}
    except Exception as e:  # noqa
        return {"s": s, "error": repr(e)[:200]}
    finally:
        signal.alarm(0)


def main():
    problem = sys.argv[1]
    sys.path.insert(0, REPO)
    enum_file = {"force_free": "enum_force_free_d4", "kerr_magnetosphere": "enum_kerr_magnetosphere_d3"}[problem]
    g = json.load(gzip.open(os.path.join(REPO, "tests", "golden", enum_file + ".json.gz"), "rt"))
    exprs = list(g["depths"]["1"]["uniques"]) + list(g["depths"]["2"]["uniques"])
    step = {"force_free": 23, "kerr_magnetosphere": 101}[problem]
    exprs += g["depths"]["3"]["uniques"][::step]
    if problem == "force_free":
        exprs += g["depths"]["4"]["uniques"][::1201]
        exprs += ["rho**2", "rho**2*z", "1 - z/sqrt(rho**2 + z**2)", "rho**2/(rho**2 + z**2)**(3/2)",
                  "sqrt(rho**2 + z**2) - z", "sqrt(z**2 + (rho - 1)**2) - sqrt(z**2 + (rho + 1)**2)",
                  "rho**2*exp(-2*z)", "rho*z", "rho**3 + z**2", "rho**2/z", "z/(1 - rho)", "rho**2 + rho*z",
                  "rho/(rho**2 + z**2)", "rho + exp(rho/z)", "rho*exp(rho/z)"]
    else:
        exprs += ["1 - x", "r*x", "exp(-r)*(1 - x)", "sqrt(r)/(1 - x)"]
    exprs = list(dict.fromkeys(exprs))
    print(len(exprs), "expressions", flush=True)
    with mp.Pool(min(8, os.cpu_count() or 1), initializer=_init, initargs=(problem,)) as pool:
        recs = pool.map(_one, exprs, chunksize=2)
    from oracle.residuals import collocation_grid
    pts = collocation_grid("force_free" if problem == "force_free" else "kerr", NPTS)
    out = {"problem": problem, "points": pts.tolist(), "records": recs,
           "consts": {} if problem == "force_free" else {"M": 1.0, "a": 0.1}}
    n_ok = sum(1 for r in recs if "jets" in r)
    print("ok", n_ok, "timeout", sum(1 for r in recs if r.get("timeout")), "error", sum(1 for r in recs if "error" in r))
    path = os.path.join(REPO, "tests", "golden", f"resid_{problem}.json.gz")
    with gzip.open(path, "wt", compresslevel=9) as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()

"""CPU confirmation stage for the survivors of the GPU filter.

Replaces the reference's validator worker pool (`_parallel_validator_worker`,
general_method_paper_reproduction.py:1672-1824), which is broken at the surveyed
commit (it imports a package that does not exist, GM:1694, and references an
un-imported class, GM:1701): every `--validators N > 0` run leaves its rows
`pending`.  The contract kept from that worker:

* one validator object PER PROCESS, rebuilt inside the worker from a picklable
  factory (validators hold sqlite handles and SymPy caches and are not pickled,
  GM:1243, GM:1695);
* the validator is called like the engine calls it: 5-kwarg form first, 2-kwarg
  form on `TypeError` (GM:1301-1316), exceptions become `(None, "Validator Error: ...")`
  (GM:1336-1339);
* `is_paper_solution` matching for valid rows (GM:1785-1798).

What is new: symbolic validation takes 0.01-450 s per candidate and
31 % of a depth-3 sample exceeds two minutes, so every task has a WALL CAP; a
worker that exceeds it is killed and replaced (a SymPy `expand` cannot be
interrupted from inside) and the row is reported as `(None, "Timeout (> cap s) ...")`
-- undecided, never "invalid".  Tasks are fed in the order given: the GPU batch
validator hands its survivors sorted by residual ratio, most plausible first.

No CUDA in here: workers are plain CPU processes (`spawn`, so a parent that
holds a CUDA context is safe).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

Verdict = Tuple[Optional[bool], str]


def _call_validator(validator: Any, u: Any, kwargs: Dict[str, Any]) -> Verdict:
    """GM:1301-1316: full kwarg set first, the two protocol kwargs on TypeError."""
    try:
        try:
            return validator.validate(u, check_regularity=False, fast_point_only=False, **kwargs)
        except TypeError:
            return validator.validate(u, check_regularity=False, fast_point_only=False)
    except Exception as e:  # GM:1336-1339
        return None, f"Validator Error: {e}"


def _worker_main(conn, factory: Callable[[], Tuple[Any, dict, dict]], kwargs: Dict[str, Any]) -> None:
    """One process = one validator (GM:1695-1701).  Protocol: recv (task_id, expr_str) | None; send
    (task_id, is_valid, reason, paper_name, seconds)."""
    import sympy as sp
    validator, locals_, known = factory()
    known_exprs = []
    for s, name in (known or {}).items():
        try:
            known_exprs.append((sp.sympify(s, locals=locals_), name))
        except Exception:
            pass
    conn.send(("ready", os.getpid()))
    while True:
        msg = conn.recv()
        if msg is None:                      # shutdown sentinel (GM:884, 1726)
            return
        tid, expr = msg
        t0 = time.perf_counter()
        paper = None
        try:
            u = sp.sympify(expr, locals=locals_)                     # GM:1257
            ok, reason = _call_validator(validator, u, kwargs)
            if ok:
                for k_expr, name in known_exprs:                     # GM:1785-1798
                    try:
                        if sp.simplify(u - k_expr) == 0:
                            paper = name
                            break
                    except Exception:
                        pass
        except Exception as e:
            ok, reason = None, f"Validator Error: {e}"
        conn.send((tid, ok, reason, paper, time.perf_counter() - t0))


class _Worker:
    def __init__(self, ctx, factory, kwargs):
        self.parent, child = ctx.Pipe()
        self.proc = ctx.Process(target=_worker_main, args=(child, factory, kwargs), daemon=True)
        self.proc.start()
        child.close()
        self.task: Optional[int] = None
        self.deadline = 0.0
        self.ready = False

    def kill(self):
        try:
            self.proc.kill()
            self.proc.join(5)
        finally:
            self.parent.close()


class ConfirmationPool:
    """`n_workers` CPU processes, each owning one validator; `time_cap_s` wall seconds per candidate."""

    def __init__(self, factory: Callable[[], Tuple[Any, dict, dict]], n_workers: Optional[int] = None,
                 time_cap_s: float = 60.0, validate_kwargs: Optional[Dict[str, Any]] = None, start_method: str = "spawn"):
        self.factory = factory
        self.n_workers = max(1, n_workers or (os.cpu_count() or 1))
        self.time_cap_s = float(time_cap_s)
        self.kwargs = dict(validate_kwargs if validate_kwargs is not None else
                           {"lean_first": True, "defer_heavy_checks": True, "enforce_anchor": False})   # GM:1304-1309
        self.ctx = mp.get_context(start_method)
        self.workers: List[_Worker] = []
        self.stats = {"confirmed": 0, "valid": 0, "timeouts": 0, "errors": 0, "cpu_seconds": 0.0, "respawned": 0}

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _spawn(self) -> _Worker:
        w = _Worker(self.ctx, self.factory, self.kwargs)
        self.workers.append(w)
        return w

    def close(self) -> None:
        for w in self.workers:
            try:
                if w.proc.is_alive() and w.task is None and w.ready:
                    w.parent.send(None)
                    w.proc.join(2)
            except Exception:
                pass
            if w.proc.is_alive():
                w.kill()
        self.workers = []

    def confirm(self, expr_strs: Sequence[str], on_result: Optional[Callable[[int, Verdict, Optional[str]], None]] = None
                ) -> List[Tuple[Optional[bool], str, Optional[str]]]:
        """Validate every string (fed in the given order); returns, per input, (is_valid | None, reason,
        paper_solution_name | None).  `on_result(index, (is_valid, reason), paper)` is called as results arrive."""
        n = len(expr_strs)
        out: List[Optional[Tuple[Optional[bool], str, Optional[str]]]] = [None] * n
        nxt = done = 0
        startup_failures = 0
        while len(self.workers) < min(self.n_workers, max(n, 1)):
            self._spawn()

        def finish(i: int, ok, reason: str, paper, secs: float) -> None:
            nonlocal done
            out[i] = (ok, reason, paper)
            done += 1
            self.stats["confirmed"] += 1
            self.stats["cpu_seconds"] += secs
            if ok:
                self.stats["valid"] += 1
            if ok is None:
                self.stats["timeouts" if reason.startswith("Timeout") else "errors"] += 1
            if on_result is not None:
                on_result(i, (ok, reason), paper)

        while done < n:
            progressed = False
            now = time.perf_counter()
            for k, w in enumerate(list(self.workers)):
                # results / readiness
                try:
                    while w.parent.poll(0):
                        msg = w.parent.recv()
                        progressed = True
                        if msg[0] == "ready":
                            w.ready = True
                            continue
                        tid, ok, reason, paper, secs = msg
                        w.task = None
                        finish(tid, ok, reason, paper, secs)
                except (EOFError, OSError):
                    # the worker died (e.g. out of memory inside SymPy): its task is undecided
                    if not w.ready:
                        startup_failures += 1
                        if startup_failures > 2 * self.n_workers:
                            self.close()
                            raise RuntimeError("ConfirmationPool: worker processes die before becoming ready "
                                               "(is the factory importable from a spawned process?)")
                    if w.task is not None:
                        finish(w.task, None, "Validator Error: worker process died", None, now - (w.deadline - self.time_cap_s))
                    w.kill()
                    self.workers[k] = _Worker(self.ctx, self.factory, self.kwargs)
                    self.stats["respawned"] += 1
                    progressed = True
                    continue
                # wall cap: a SymPy call cannot be interrupted from inside -> kill and replace the process
                if w.task is not None and now > w.deadline:
                    finish(w.task, None, f"Timeout (> {self.time_cap_s:g} s of symbolic validation; undecided)", None, self.time_cap_s)
                    w.kill()
                    self.workers[k] = _Worker(self.ctx, self.factory, self.kwargs)
                    self.stats["respawned"] += 1
                    progressed = True
                    continue
                # feed
                if w.ready and w.task is None and nxt < n:
                    w.parent.send((nxt, expr_strs[nxt]))
                    w.task = nxt
                    w.deadline = time.perf_counter() + self.time_cap_s
                    nxt += 1
                    progressed = True
            if not progressed:
                time.sleep(0.002)
        return out  # type: ignore[return-value]


# ---- picklable validator factory (top-level class: `spawn` pickles it by reference) ----
# (the tests use factories over the oracle's restatement of the validators: tests/pool_factories.py)

class ReferenceFactory:
    """Factory for the reference's own validator: rebuilds the problem inside the worker from its slug,
    exactly like GM:1695 does.  `reference_root` is put on sys.path first."""

    def __init__(self, slug: str, reference_root: str):
        self.slug, self.root = slug, reference_root

    def __call__(self):
        import sys
        if self.root not in sys.path:
            sys.path.insert(0, self.root)
        from problems import load_problem as ref_load_problem          # PI:355-361
        problem = ref_load_problem(self.slug)
        locals_ = {**problem.symbols, **getattr(problem, "constants", {}), **problem.unary_ops}           # GM:85-93
        return problem.validator, locals_, dict(problem.known_solutions)

"""Picklable validator factories for the confirmation-pool tests (top-level functions: the `spawn`
start method pickles them by reference).  They build the ORACLE restatement of the reference's
validators -- test infrastructure, which the product package itself never imports."""
import time


def oracle_force_free_factory():
    from oracle import symbolic as osym
    from pde_engine_b200.problems import load_problem
    spec = load_problem("force_free", make_gpu=False)
    return osym.ForceFreeSymbolicValidator(), spec.sympify_locals(), dict(spec.known_solutions)


class _SlowValidator:
    """validate() that sleeps: the cap must kill and replace the worker, the other tasks go on."""

    def validate(self, u, check_regularity=True, fast_point_only=False, **kw):
        s = str(u)
        if "exp" in s:
            time.sleep(30)
        return ("rho" in s), f"checked {s}"


def slow_factory():
    import sympy as sp
    rho = sp.Symbol("rho", real=True, positive=True)
    z = sp.Symbol("z", real=True)
    return _SlowValidator(), {"rho": rho, "z": z}, {}

"""CPU confirmation pool (replaces the broken worker pool GM:1672-1824):
per-process validators, wall cap per candidate, results in input order, paper-solution matching."""
import time

from pde_engine_b200.confirm import ConfirmationPool
from pool_factories import oracle_force_free_factory, slow_factory


def test_pool_confirms_known_solutions_and_names_them():
    strs = ["rho**2*z", "rho*z", "-z + sqrt(rho**2 + z**2)", "square(rho)", "rho + z", "1"]
    with ConfirmationPool(oracle_force_free_factory, n_workers=2, time_cap_s=120) as pool:
        got = pool.confirm(strs)
    # rho + z is a valid (unnamed) foliation: B = 2 is constant, both Lie derivatives of B vanish
    assert [g[0] for g in got] == [True, False, True, True, True, False]
    assert got[0][1].startswith("Valid foliation") and got[1][1].startswith("Invalid")
    assert got[5][1].startswith("Zero gradient")                       # FFV:309-312
    assert [g[2] for g in got] == ["X-point", None, "Parabolic", "Vertical field", None, None]   # GM:1785-1798
    assert pool.stats["confirmed"] == 6 and pool.stats["valid"] == 4 and pool.stats["timeouts"] == 0


def test_wall_cap_kills_and_replaces_the_worker():
    strs = ["rho", "exp(z)", "z", "rho*z", "exp(rho)", "rho + 1"]
    seen = []
    t0 = time.time()
    with ConfirmationPool(slow_factory, n_workers=2, time_cap_s=1.0) as pool:
        got = pool.confirm(strs, on_result=lambda i, v, p: seen.append(i))
    assert time.time() - t0 < 25                                       # two 30 s sleeps were cut at 1 s
    assert got[1][0] is None and got[1][1].startswith("Timeout") and got[4][0] is None
    assert "Error" not in got[1][1] and "Could not" not in got[1][1]   # GM:218 would read those as errors
    assert [g[0] for i, g in enumerate(got) if i not in (1, 4)] == [True, False, True, True]
    assert sorted(seen) == list(range(6))
    assert pool.stats["timeouts"] == 2 and pool.stats["respawned"] == 2


def test_validator_exceptions_become_error_rows():
    with ConfirmationPool(oracle_force_free_factory, n_workers=1, time_cap_s=60) as pool:
        got = pool.confirm(["rho +* z", "rho**2"])
    assert got[0][0] is None and got[0][1].startswith("Validator Error")    # GM:1336-1339
    assert got[1][0] is True

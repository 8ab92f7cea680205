"""CPU oracle for the pde-engine hot path (TEST INFRASTRUCTURE, NOT PRODUCT).

This package restates, in plain Python / numpy, the algorithms of the reference
(PimDeWitte/pde-engine) on the hot path named by BASELINE.json:

* ``oracle.enumerate``  -- candidate construction + prune predicates
                           (lean_normalizer/lean_bridge_fixed.py:113-215)
* ``oracle.normalizer`` -- the "Lean" canonicaliser, which is SymPy
                           (lean_normalizer/lean_bridge.py:67-92,
                           lean_bridge_fixed.py:42-68)
* ``oracle.parser``     -- string -> term-structured postfix bytecode
                           (what sympify does at general_method_paper_reproduction.py:1257)
* ``oracle.jets``       -- float64 Taylor-mode jets, order N, two variables
                           (generalises problems/force_free/validator.py:70-180)
* ``oracle.residuals``  -- the two PDE residual operators
                           (problems/force_free/validator.py:305-347,
                           problems/kerr_magnetosphere/validator.py:77-91)
* ``oracle.symbolic``   -- restatement of the reference's symbolic validators
                           used as the CPU confirmation stage in tests

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``pde_engine_b200``) never does: it fails loudly when its CUDA library is
missing.

Parity pinning: the oracle is checked against (a) the reference's committed
fixtures (``tests/golden/ref_fixture_*.json``, extracted from the reference's
run DB / caches) and (b) outputs of the unmodified reference run in the build
container (``tests/golden/make_golden.py``), see ``tests/test_oracle_*.py``.
"""

"""Oracle: the reference's "Lean" canonicaliser, which is SymPy (TEST INFRASTRUCTURE).

Restates ``LeanNormalizer.normalize`` / ``_canonical_form`` / ``_apply_rules``
(lean_normalizer/lean_bridge.py:67-112) and ``normalize_batch``
(lean_normalizer/lean_bridge_fixed.py:42-68, minus the SQLite cache, which
only memoises).  The arithmetic lives in SymPy (third party; the reference
pins no version, the build container has sympy 1.14.0 / mpmath 1.3.0 and the
committed fixtures reproduce bit-exactly with it).

The canonicaliser stays on the CPU by design (BASELINE.json north_star); the
product takes *any* object with ``normalize_batch`` -- in production the
reference's own ``LeanNormalizer``, in the tests this restatement.
"""
from __future__ import annotations

import hashlib
from typing import Any, Dict, List, Tuple

import sympy as sp


def normalize(expr_str: str) -> str:
    """LB:67-79 (sympify WITHOUT locals: neg/inv/square/... stay opaque)."""
    try:
        expr = sp.sympify(expr_str)
        expr = sp.expand(expr)  # LB:84
        if expr.has(sp.Symbol("rho")) and expr.has(sp.Symbol("z")):  # LB:87-88
            expr = sp.collect(expr, [sp.Symbol("rho"), sp.Symbol("z")])
        # LB:95-112 -- with a *positive* rho these patterns auto-evaluate to
        # their replacements, so the substitutions are identities; kept for
        # faithfulness.
        rho = sp.Symbol("rho", positive=True)
        z = sp.Symbol("z")
        for pattern, replacement in [
            (sp.exp(sp.log(rho)), rho),
            (sp.log(sp.exp(z)), z),
            (sp.sqrt(rho ** 2), rho),
            (rho / rho, 1),
            (z - z, 0),
        ]:
            expr = expr.subs(pattern, replacement)
        return str(expr)
    except Exception:
        return expr_str


class OracleNormalizer:
    """Duck-type of the reference's ``LeanNormalizer`` (LBF:19-68)."""

    def __init__(self) -> None:
        self._memo: Dict[str, str] = {}

    def normalize(self, expr_str: str) -> str:
        hit = self._memo.get(expr_str)
        if hit is None:
            hit = normalize(expr_str)
            self._memo[expr_str] = hit
        return hit

    def normalize_batch(self, expressions: List[Tuple[str, int]]) -> List[Dict[str, Any]]:
        out = []
        for expr_str, idx in expressions:
            norm = self.normalize(expr_str)
            sig = hashlib.sha256(norm.encode()).hexdigest()[:16]
            out.append({"normalized": norm, "index": idx, "signature": sig})
        return out

    def preload(self, mapping: Dict[str, str]) -> None:
        """Warm the memo (e.g. from a golden fixture) -- memoisation only."""
        self._memo.update(mapping)

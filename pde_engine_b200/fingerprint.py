"""Function fingerprints on the device.

The stage-1 normaliser and the DB-normalisation of ``emit_to_db`` (GM:1267-1278) both parse WITHOUT the
problem's locals, so the custom operators stay opaque to them: ``neg(neg(rho))``, ``inv(inv(rho))`` and
``rho`` are three unique rows, each validated on its own -- and for a survivor that means
0.2-400 s of SymPy per row for a verdict that is a property of the FUNCTION.

``GpuFingerprinter`` evaluates every candidate (parsed WITH the locals, like ``validate``'s argument,
GM:1257) at a few collocation points with the stage-2 interpreter and returns a 64-bit key of the rounded
values (``pde_fingerprint``): rows with the same key denote the same function up to 2^-mantissa_bits at
every point.  A key of 0 (no finite value: not compilable, non-finite or complex everywhere -- the device
analogue of ``_has_degenerate_denominator``, GM:134-199) means "unknown" and is left to the CPU.

The bucketing is numeric, so it only ever drives opt-in behaviour: ``run_discovery(share_confirmations=...)``
confirms one representative per bucket, ``FunctionDedup`` is a run-wide first-occurrence filter for
callers that want one row per function.  The default path is unchanged.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import core
from .grids import RANGES, canonical_slug, splitmix64

# Generic values for the symbolic parameters: with the reference's check values (M = 1, a = 1/10, PI:283)
# `M*r` and `r` would share a fingerprint.
GENERIC_CONSTS = {
    "force_free": {},
    "kerr_magnetosphere": {"M": 1.1378240173, "a": 0.2718653942},
}
VARS = {"force_free": ("rho", "z"), "kerr_magnetosphere": ("r", "x")}
FP_GRID_SEED = 0x5EEDF1A6


def fingerprint_grid(problem: str, P: int, seed: int = FP_GRID_SEED) -> np.ndarray:
    """[2, P] points for fingerprints: uniform in the validation ranges, none of the reference's rational test
    points (coincidences), and BOTH signs of the second coordinate -- z is real, not positive (PI:70-71), so
    ``sqrt(z**2) = Abs(z)`` must not share a key with ``z``."""
    lo0, w0, lo1, w1 = RANGES[canonical_slug(problem)]
    u = (splitmix64(seed, 3 * P) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    pts = np.empty((2, P))
    pts[0] = lo0 + w0 * u[0::3]
    pts[1] = lo1 + w1 * u[1::3]
    if lo1 > 0:
        pts[1] = np.where(u[2::3] < 0.5, -pts[1], pts[1])
    return pts


class Fingerprints:
    """Result of one batch: ``key`` uint64 [n] (0 = unknown), ``n_finite`` int32 [n], ``values`` f64 [n, P]."""

    def __init__(self, strs: List[str], key: np.ndarray, n_finite: np.ndarray, values: Optional[np.ndarray]):
        self.strs, self.key, self.n_finite, self.values = strs, key, n_finite, values

    def buckets(self) -> np.ndarray:
        """bucket[i] = index of the first row of the batch with the same key; -1 for unknown rows."""
        out = np.full(len(self.key), -1, np.int64)
        known = np.nonzero(self.key != 0)[0]
        if known.size:
            _, first, inv = np.unique(self.key[known], return_index=True, return_inverse=True)
            out[known] = known[first[inv]]
        return out


    def groups(self) -> List[dict]:
        """Equivalence classes of the batch, largest first -- the grouping the reference's report builds with a SymPy
        canonicalisation pipeline per row (GM:1918-2008): ``{'rep': index of the shortest member, 'members': [...]}``;
        unknown rows are singletons."""
        by_key: Dict[int, List[int]] = {}
        out: List[dict] = []
        for i, k in enumerate(self.key.tolist()):
            if k == 0:
                out.append({"rep": i, "members": [i]})
            else:
                by_key.setdefault(k, []).append(i)
        for members in by_key.values():
            out.append({"rep": min(members, key=lambda i: (len(self.strs[i]), i)), "members": members})
        out.sort(key=lambda g: (-len(g["members"]), self.strs[g["rep"]]))      # GM:2010: descending size, then representative
        return out


class GpuFingerprinter:
    def __init__(self, problem: str = "force_free", P: int = 64, mantissa_bits: int = 26, L: int = 128,
                 spill_slots: int = 4, device=None, keep_values: bool = False):
        import torch
        self.problem = canonical_slug(problem)
        self.session = core.Session(VARS[self.problem], GENERIC_CONSTS[self.problem])
        self.P, self.mantissa_bits, self.L, self.spill_slots = P, mantissa_bits, L, spill_slots
        self.keep_values = keep_values
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.pts_host = fingerprint_grid(self.problem, P)
        self.pts = torch.from_numpy(self.pts_host).to(self.device)
        self.stats = {"fingerprinted": 0, "unknown": 0}

    def fingerprint(self, expr_strs: Sequence[str]) -> Fingerprints:
        import torch
        strs = list(expr_strs)
        n = len(strs)
        if n == 0:
            return Fingerprints(strs, np.zeros(0, np.uint64), np.zeros(0, np.int32), np.zeros((0, self.P)) if self.keep_values else None)
        exprs = self.session.compile(strs)
        code_h, len_h = exprs.programs(self.L)
        code_t = torch.from_numpy(code_h).to(self.device)
        len_t = torch.from_numpy(len_h).to(self.device)
        values, key, n_finite = core.fingerprint(self.session, code_t, len_t, self.pts, None,
                                                 mantissa_bits=self.mantissa_bits, spill_slots=self.spill_slots)
        key_h = key.cpu().numpy().view(np.uint64)
        nf_h = n_finite.cpu().numpy()
        self.stats["fingerprinted"] += n
        self.stats["unknown"] += int((key_h == 0).sum())
        return Fingerprints(strs, key_h, nf_h, values.cpu().numpy() if self.keep_values else None)


class FunctionDedup:
    """Run-wide first-occurrence filter on fingerprint keys: one row per function (stricter than the reference's
    UNIQUE(normalized), GM:1407, which compares strings that keep the custom operators opaque).

    ``filter(strs)`` returns a boolean mask: True = keep (first row with this key in the run, or unknown)."""

    def __init__(self, fingerprinter: GpuFingerprinter):
        self.fp = fingerprinter
        self.seen: Dict[int, str] = {}
        self.stats = {"rows": 0, "function_duplicates": 0, "unknown": 0}

    def filter(self, expr_strs: Sequence[str]) -> Tuple[np.ndarray, List[Optional[str]]]:
        f = self.fp.fingerprint(expr_strs)
        keep = np.ones(len(f.strs), bool)
        same_as: List[Optional[str]] = [None] * len(f.strs)
        for i, k in enumerate(f.key.tolist()):
            if k == 0:
                self.stats["unknown"] += 1
                continue
            first = self.seen.get(k)
            if first is None:
                self.seen[k] = f.strs[i]
            else:
                keep[i] = False
                same_as[i] = first
        self.stats["rows"] += len(f.strs)
        self.stats["function_duplicates"] += int((~keep).sum())
        return keep, same_as

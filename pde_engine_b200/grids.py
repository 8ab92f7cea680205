"""Collocation grids: the reference's own rational test points first
(FFV:296-297 + LBF:278-282 for force-free, KV:168-172 for Kerr), then uniform
points drawn with splitmix64."""
from __future__ import annotations

import numpy as np

GRID_SEED = 0x5EED9017
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def _mix64(x: np.ndarray) -> np.ndarray:
    x = x.copy()
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def splitmix64(seed: int, n: int) -> np.ndarray:
    """First n outputs of the splitmix64 stream seeded with `seed`."""
    with np.errstate(over="ignore"):
        k = np.arange(1, n + 1, dtype=np.uint64)
        return _mix64(np.uint64(seed) + k * _GOLD)


REFERENCE_POINTS = {
    "force_free": [(4 / 5, 6 / 7), (3 / 4, 5 / 6), (7 / 8, 1 / 2)],
    "kerr_magnetosphere": [(5 / 2, 3 / 5), (7 / 3, 1 / 3), (5.0, -2 / 5)],
}
RANGES = {
    "force_free": (0.25, 1.75, 0.25, 1.75),          # rho, z in [0.25, 2]
    "kerr_magnetosphere": (2.2, 3.8, -0.9, 1.8),     # r in [2.2, 6], x in [-0.9, 0.9]
}


def canonical_slug(name: str) -> str:
    key = (name or "").strip().lower()
    if key in ("force_free", "forcefree", "foliation", "foliations"):
        return "force_free"
    if key in ("kerr", "kerr_magnetosphere", "kerr-magnetosphere"):
        return "kerr_magnetosphere"
    raise ValueError(f"Unknown problem '{name}'. Available: 'force_free', 'kerr_magnetosphere'")


def collocation_grid(problem: str, P: int, seed: int = GRID_SEED) -> np.ndarray:
    """[2, P] float64, SoA (row 0 = rho|r, row 1 = z|x)."""
    slug = canonical_slug(problem)
    ref = REFERENCE_POINTS[slug]
    lo0, w0, lo1, w1 = RANGES[slug]
    pts = np.empty((2, P))
    nref = min(len(ref), P)
    for k in range(nref):
        pts[0, k], pts[1, k] = ref[k]
    m = P - nref
    if m > 0:
        u = (splitmix64(seed, 2 * m) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        pts[0, nref:] = lo0 + w0 * u[0::2]
        pts[1, nref:] = lo1 + w1 * u[1::2]
    return pts

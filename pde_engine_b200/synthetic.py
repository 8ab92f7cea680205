"""Synthetic depth-5 workload of SURVEY 8d (BASELINE.json configs[4]):
10^7 bytecode trees x 4096 collocation points, force-free residual."""
from __future__ import annotations

from . import core

SEED_TREES = 0x5EED0005
PRIM_EXPRS = ("rho**2 + z**2", "rho/z")   # PRIM(0), PRIM(1): problems/__init__.py:76-77


def primitive_jets(session: core.Session, program: core.ResidualProgram, pts, table):
    """[n_prim, n_coef, P] jets of the PRIM(p) leaves, computed on the device by the
    interpreter itself from the primitives' own programs."""
    import torch
    es = session.compile(list(PRIM_EXPRS))
    code, ln = es.programs(16)
    jets, _, _ = core.eval_points(session, program, torch.from_numpy(code).to(pts.device),
                                  torch.from_numpy(ln).to(pts.device), pts, table, None,
                                  spill_slots=2, want_resid=False)
    return jets.contiguous()

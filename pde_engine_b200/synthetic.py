"""Synthetic depth-5 workload of SURVEY 8d (BASELINE.json configs[4]):
10^7 bytecode trees x 4096 collocation points, force-free residual."""
from __future__ import annotations

from . import core

SEED_TREES = 0x5EED0005
PRIM_EXPRS = ("rho**2 + z**2", "rho/z")   # PRIM(0), PRIM(1): problems/__init__.py:76-77


PRIM_ROW = 16   # coefficient rows per 32-point stripe block of the device table (n_coef padded)


def pack_primitive_table(jets, maj=None):
    """[n_prim, n_coef, P] jets (+ [n_prim, 3, P] float32 majorants (V, D, W)) -> the [n_prim, P/32, 16, 32] table
    pde_validate reads (include/pde_b200.h): per 32-point stripe the coefficients are rows of 32 consecutive
    lanes, so a warp fetches a PRIM leaf with fully coalesced loads at immediate offsets (g * 256 B) from one
    per-lane base address; row 15 carries the leaf's (D, W) as two float32 in one 8-byte element."""
    import torch
    n_prim, n_coef, P = jets.shape
    assert P % 32 == 0 and n_coef < PRIM_ROW
    tab = torch.zeros((n_prim, P // 32, PRIM_ROW, 32), dtype=torch.float64, device=jets.device)
    tab[:, :, :n_coef, :] = jets.reshape(n_prim, n_coef, P // 32, 32).permute(0, 2, 1, 3)
    if maj is not None:
        pair = maj[:, 1:3, :].permute(0, 2, 1).contiguous().view(torch.float64).reshape(n_prim, P // 32, 32)   # (D, W) -> one f64 slot
        tab[:, :, PRIM_ROW - 1, :] = pair
    return tab.contiguous()


def unpack_primitive_table(tab, n_coef: int):
    """inverse of pack_primitive_table -> [n_prim, n_coef, P]"""
    n_prim, nb = tab.shape[0], tab.shape[1]
    return tab[:, :, :n_coef, :].permute(0, 2, 1, 3).reshape(n_prim, n_coef, nb * 32).contiguous()


def primitive_jets(session: core.Session, program: core.ResidualProgram, pts, table, packed: bool = True):
    """Jets of the PRIM(p) leaves, computed on the device by the interpreter itself from the
    primitives' own programs: the packed [n_prim, P/32, 16, 32] device table, or (packed=False) the plain
    [n_prim, n_coef, P] jets."""
    import torch
    es = session.compile(list(PRIM_EXPRS))
    code, ln = es.programs(16)
    jets, _, _, _, maj = core.eval_points(session, program, torch.from_numpy(code).to(pts.device),
                                          torch.from_numpy(ln).to(pts.device), pts, table, None,
                                          spill_slots=2, want_resid=False, want_maj=True)
    return pack_primitive_table(jets, maj) if packed else jets.contiguous()

"""Algorithmic flop accounting of the jet interpreter.

One unit of work = one (candidate, collocation point) pair.  Per-op costs for
order-N, 2-variable jets with C = (N+1)(N+2)/2 coefficients and PI = C(N+4,4)
coefficient pairs in a truncated product (DFMA = 2 flop, DADD/DMUL/DDIV/sqrt/exp = 1):

    VAR/CONST 0 | PRIM(p) nnz(p) | add sub neg abs  C | mul 2 PI | square PI + C
    inv 2(PI-C)+C | exp 2(PI-C)+2C | exp_neg exp+C | sqrt pow(k) 3(PI-C)+C | div inv+mul

N=4: add 15, mul 140, square 85, inv 125, exp 140, exp_neg 155, pow 180, div 265.
N=2: add 6, mul 30, square 21, inv 24, exp 30, exp_neg 36, pow 33, div 54.
Residual programs (CSE'd straight-line form): force-free 103, Kerr 7.
No credit is taken for common sub-expressions or leaf sparsity: every candidate
is counted as an independent dense evaluation.
"""
from __future__ import annotations

import numpy as np

RESIDUAL_FLOPS = {"force_free": 206, "kerr_magnetosphere": 14}   # residual + its abs-propagated scale S ("add one more cost(residual)")
# non-zero jet coefficients of the synthetic primitives PRIM(0) = rho**2 + z**2, PRIM(1) = rho/z
PRIM_NNZ = (6, 9)


def op_cost_table(order: int, prim_nnz=PRIM_NNZ) -> np.ndarray:
    """cost[byte] for every opcode byte (include/pde_b200.h)."""
    C = (order + 1) * (order + 2) // 2
    PI = {4: 70, 2: 15}[order]
    add, mul, square = C, 2 * PI, PI + C
    inv = 2 * (PI - C) + C
    exp = 2 * (PI - C) + 2 * C
    powk = 3 * (PI - C) + C
    t = np.zeros(256, dtype=np.int64)
    for p in range(8):
        t[0x08 + p] = prim_nnz[p] if p < len(prim_nnz) else C
    t[0x10], t[0x11], t[0x12], t[0x13] = add, add, mul, inv + mul
    t[0x18], t[0x19], t[0x1A], t[0x1B] = add, add, powk, exp
    t[0x20], t[0x21], t[0x22], t[0x23], t[0x24], t[0x25] = add, inv, square, powk, powk, exp + C
    t[0x40:0x80] = powk
    return t


def batch_flops_per_point(code, problem: str = "force_free", order: int = 4) -> int:
    """Sum over a [n, L] uint8 CUDA/CPU tensor of programs of the algorithmic flops
    for ONE collocation point (multiply by P for the launch)."""
    import torch
    hist = torch.bincount(code.reshape(-1).to(torch.int64), minlength=256).cpu().numpy().astype(np.int64)
    ops = int((hist * op_cost_table(order)).sum())
    return ops + int(code.shape[0]) * RESIDUAL_FLOPS[problem]

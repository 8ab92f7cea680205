"""Oracle: synthetic depth-d trees + algorithmic flop table (TEST INFRASTRUCTURE).

SURVEY 8d: each tree is built like a reference depth-d candidate
(lean_bridge_fixed.py:139-195) -- with probability 8/13 a unary op (uniform over
the 8 of expression_operations.py:80-89) on a depth-(d-1) tree, else one of the
5 live binary ops on trees of depths (d1, d-d1), d1 uniform in 1..d-1; leaves
are the 5 force-free primitives (problems/__init__.py:73-79) uniformly; tree
semantics (no textual splice), no pruning.  Random numbers: splitmix64, one
stream per tree, state0 = mix64(seed ^ mix64(index + 1)).

Postfix order: left, right, op;  geom_sum(a, b) = a 1 b SUB DIV.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from . import bytecode as bc

SEED_TREES = 0x5EED0005
GOLD = 0x9E3779B97F4A7C15
LEAVES = (bc.OP_VAR0, bc.OP_VAR1, bc.OP_PRIM0, bc.OP_PRIM0 + 1, bc.OP_CONST0)
PRIM_EXPRS = ("rho**2 + z**2", "rho/z")     # PRIM(0), PRIM(1)


def tree(seed: int, index: int, depth: int) -> bytes:
    st = bc.mix64(seed ^ bc.mix64(index + 1))

    def nxt():
        nonlocal st
        st = (st + GOLD) & bc.MASK64
        return bc.mix64(st)

    out = bytearray()
    stack = [0x100 + depth]
    while stack:
        it = stack.pop()
        if it < 0x100:
            out.append(it)
            continue
        d = it - 0x100
        if d <= 1:
            out.append(LEAVES[nxt() % 5])
            continue
        r = nxt() % 13
        if r < 8:
            stack.append(bc.UNARY_OPCODES[r])
            stack.append(0x100 + d - 1)
        else:
            d1 = 1 + nxt() % (d - 1)
            bop = r - 8
            if bop == 4:
                stack += [bc.OP_DIV, bc.OP_SUB, 0x100 + d - d1, bc.OP_CONST0, 0x100 + d1]
            else:
                stack += [bc.OP_ADD + bop, 0x100 + d - d1, 0x100 + d1]
    return bytes(out)


def trees(seed: int, first: int, count: int, depth: int) -> List[bytes]:
    return [tree(seed, first + i, depth) for i in range(count)]


# ---------------------------------------------------------------------------
# Algorithmic flops per (candidate, point): SURVEY 8d table.
# C = coefficients, PI = coefficient pairs of a truncated product.
# ---------------------------------------------------------------------------

def flop_table(order: int, prim_nnz: Sequence[int] = (6, 9)) -> Dict[int, int]:
    C = (order + 1) * (order + 2) // 2
    PI = {4: 70, 2: 15}[order]
    add, mul, square = C, 2 * PI, PI + C
    inv = 2 * (PI - C) + C
    exp = 2 * (PI - C) + 2 * C
    powk = 3 * (PI - C) + C
    t: Dict[int, int] = {bc.OP_VAR0: 0, bc.OP_VAR1: 0}
    for k in range(bc.N_CONST):
        t[bc.OP_CONST0 + k] = 0
    for p in range(bc.N_PRIM):
        t[bc.OP_PRIM0 + p] = prim_nnz[p] if p < len(prim_nnz) else C
    t.update({bc.OP_ADD: add, bc.OP_SUB: add, bc.OP_MUL: mul, bc.OP_DIV: inv + mul,
              bc.OP_NEG: add, bc.OP_ABS: add, bc.OP_SQRT: powk, bc.OP_EXP: exp,
              bc.OP_FN_NEG: add, bc.OP_FN_INV: inv, bc.OP_FN_SQUARE: square,
              bc.OP_FN_POW32: powk, bc.OP_FN_POWN32: powk, bc.OP_FN_EXPNEG: exp + C})
    for k in range(bc.N_POW):
        t[bc.OP_POW0 + k] = powk
    return t


RESIDUAL_FLOPS = {"force_free": 206, "kerr_magnetosphere": 14}   # residual + its abs-propagated scale S ("add one more cost(residual)")   # SURVEY 8d (CSE'd straight-line form)


def program_flops(code: bytes, table: Dict[int, int]) -> int:
    return sum(table[b] for b in code)

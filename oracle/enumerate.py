"""Oracle: candidate enumeration (TEST INFRASTRUCTURE, NOT PRODUCT).

Restates ``FastExpressionGenerator.stream_generate`` of the reference,
lean_normalizer/lean_bridge_fixed.py:113-215:

* loop order: unary first (LBF:142-153), then binary over d1 = 1..d-1
  (LBF:155-195); both (d1, d2) and (d2, d1) are visited;
* operand swap for add / mul when ``a > b`` as Python strings (LBF:168-169);
* templates ``"(a + b)"``, ``"(a - b)"``, ``"(a * b)"``, ``"(a / (b))"``,
  ``"(a / (1 - b))"`` (LBF:170-195) -- operands are NOT parenthesised;
* prune predicates on the strings (LBF:134-136, 143-152, 162-195);
* the four "special" binary ops fall through every ``elif`` and emit nothing;
* dedup: first occurrence of ``sha256(normalised)[:16]`` wins, the seen-set is
  shared across depths >= 2 and does not contain the primitives (LBF:137,198-215);
* ``on_batch(depth, new_uniques)`` per ``batch_size`` chunk of candidates,
  only when the chunk produced something (LBF:211-212).

A candidate is described either by its string or by the triple
``(op, i, j)``: op index into UNARY/BINARY tables, i/j = *global* indices of
the operands in the concatenation E[1] ++ E[2] ++ ... (the representation the
device enumerator works on).
"""
from __future__ import annotations

import hashlib
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

# Order is the iteration order of the reference's dicts
# (expression_operations.py:80-106).
UNARY_NAMES = ("neg", "inv", "sqrt", "square", "pow_3_2", "pow_neg_3_2", "exp", "exp_neg")
BINARY_NAMES = ("add", "sub", "mul", "div", "geom_sum")
DEAD_BINARY_NAMES = ("sqrt_shift_neg", "sqrt_shift_pos", "exp_mul", "log_mul")

# op codes used in (op, i, j) triples: unary 0..7, binary 8..12
OP_UNARY0 = 0
OP_BINARY0 = 8


def has_vars(s: str) -> bool:
    """LBF:134-136 -- a *substring* test, not a symbol test."""
    return ("r" in s) or ("x" in s) or ("rho" in s) or ("z" in s)


def unary_string(op: str, a: str) -> str:
    return f"{op}({a})"  # LBF:153


def binary_string(op: str, a: str, b: str) -> Optional[str]:
    """String for ``op(a, b)`` *after* the add/mul swap; None for dead ops."""
    if op == "add":
        return f"({a} + {b})"
    if op == "sub":
        return f"({a} - {b})"
    if op == "mul":
        return f"({a} * {b})"
    if op == "div":
        return f"({a} / ({b}))"
    if op == "geom_sum":
        return f"({a} / (1 - {b}))"
    return None


def candidates_for_depth(
    E: Dict[int, List[str]],
    depth: int,
    unary_ops: Sequence[str] = UNARY_NAMES,
    binary_ops: Sequence[str] = BINARY_NAMES + DEAD_BINARY_NAMES,
    prune: bool = True,
    with_triples: bool = False,
):
    """Ordered candidate strings of ``depth`` from E[1..depth-1] (LBF:139-195).

    With ``with_triples`` also returns ``(op, i, j)`` per candidate where i, j
    index the concatenation E[1] ++ ... ++ E[depth-1] and, for add/mul, are
    listed *after* the swap (i is the left operand of the template).
    """
    base = {}
    off = 0
    for k in range(1, depth):
        base[k] = off
        off += len(E[k])
    out: List[str] = []
    triples: List[Tuple[int, int, int]] = []
    # unary, LBF:142-153
    for ia, expr in enumerate(E[depth - 1]):
        if prune and not has_vars(expr):
            continue
        for uo, op_name in enumerate(unary_ops):
            if prune:
                if op_name == "inv" and expr.startswith("inv("):
                    continue
                if op_name in ("sqrt", "square", "pow_3_2", "pow_neg_3_2") and expr == "1":
                    continue
            out.append(unary_string(op_name, expr))
            if with_triples:
                triples.append((OP_UNARY0 + UNARY_NAMES.index(op_name), base[depth - 1] + ia, -1))
    # binary, LBF:155-195
    for d1 in range(1, depth):
        d2 = depth - d1
        if d2 < 1 or d2 >= depth:
            continue
        for i1, expr1 in enumerate(E[d1]):
            for i2, expr2 in enumerate(E[d2]):
                if prune and (not has_vars(expr1)) and (not has_vars(expr2)):
                    continue
                for op_name in binary_ops:
                    a, b = expr1, expr2
                    ga, gb = base[d1] + i1, base[d2] + i2
                    if op_name in ("add", "mul") and a > b:
                        a, b = b, a
                        ga, gb = gb, ga
                    if op_name == "add":
                        pass
                    elif op_name == "sub":
                        if prune and a == b:
                            continue
                    elif op_name == "mul":
                        if prune and (a == "1" or b == "1"):
                            continue
                    elif op_name == "div":
                        if prune and (b == "1" or a == b):
                            continue
                    elif op_name == "geom_sum":
                        if prune and b == "1":
                            continue
                    else:
                        continue  # dead ops emit nothing
                    out.append(binary_string(op_name, a, b))
                    if with_triples:
                        triples.append((OP_BINARY0 + BINARY_NAMES.index(op_name), ga, gb))
    if with_triples:
        return out, triples
    return out


def count_candidates(E: Dict[int, List[str]], depth: int, prune: bool = True) -> int:
    """Number of candidates ``candidates_for_depth`` emits, by counting (LBF:139-195 rule by rule, no strings
    built): the size check for depths where the list itself is too long to build (depth 5: 1.2e7)."""
    prev = E[depth - 1]
    if not prune:
        n = len(UNARY_NAMES) * len(prev)
        for d1 in range(1, depth):
            n += len(BINARY_NAMES) * len(E[d1]) * len(E[depth - d1])
        return n
    n = sum(len(UNARY_NAMES) - (1 if e.startswith("inv(") else 0) for e in prev if has_vars(e))     # LBF:142-153
    for d1 in range(1, depth):
        A, B = E[d1], E[depth - d1]
        nv_a = sum(1 for a in A if not has_vars(a))
        nv_b = sum(1 for b in B if not has_vars(b))
        live = len(A) * len(B) - nv_a * nv_b                        # LBF:162: at least one operand has variables
        setb = set(B)
        same = sum(1 for a in A if a in setb and has_vars(a))       # pairs a == b that passed the line above
        one_a, one_b = ("1" in A), ("1" in setb)                    # '1' has no variables: its partner must have some
        with_one_b = (len(A) - nv_a) if one_b else 0                # live pairs (a, '1')
        with_one_a = (len(B) - nv_b) if one_a else 0                # live pairs ('1', b)
        n += live                                                   # add
        n += live - same                                            # sub: a != b
        n += live - with_one_a - with_one_b                         # mul: neither is '1' (a live pair has at most one)
        n += live - with_one_b - same                               # div: b != '1', a != b ('1' == '1' is not live)
        n += live - with_one_b                                      # geom_sum: b != '1'
    return n


def signature(normalized: str) -> str:
    """LBF:55,66"""
    return hashlib.sha256(normalized.encode()).hexdigest()[:16]


def stream_generate(
    primitive_strs: Sequence[str],
    normalize: Callable[[str], str],
    max_depth: int,
    batch_size: int = 1000,
    on_batch: Optional[Callable[[int, List[str]], None]] = None,
    prune: bool = True,
    unary_ops: Sequence[str] = UNARY_NAMES,
    binary_ops: Sequence[str] = BINARY_NAMES + DEAD_BINARY_NAMES,
    on_candidates: Optional[Callable[[int, List[str]], None]] = None,
) -> Dict[int, List[str]]:
    """Full restatement of LBF:113-215 with ``normalize`` injected."""
    E: Dict[int, List[str]] = {1: list(primitive_strs)}
    if on_batch:
        on_batch(1, list(primitive_strs))
    seen = set()
    for depth in range(2, max_depth + 1):
        cands = candidates_for_depth(E, depth, unary_ops, binary_ops, prune)
        if on_candidates:
            on_candidates(depth, cands)
        uniq: List[str] = []
        for i in range(0, len(cands), batch_size):
            chunk_out: List[str] = []
            for s in cands[i:i + batch_size]:
                norm = normalize(s)
                sig = signature(norm)
                if sig not in seen:
                    seen.add(sig)
                    uniq.append(norm)
                    chunk_out.append(norm)
            if on_batch and chunk_out:
                on_batch(depth, chunk_out)
        E[depth] = uniq
    return E

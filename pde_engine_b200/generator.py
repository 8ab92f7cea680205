"""GpuExpressionGenerator -- drop-in for the reference's FastExpressionGenerator
(lean_normalizer/lean_bridge_fixed.py:71-215).

Same constructor argument (a normaliser with ``normalize_batch``), same
``stream_generate`` / ``generate_expressions`` signatures, same ``on_batch``
protocol (GM:275-282, GM:1415-1423): ``on_batch(1, [str(p) ...])`` first, then,
per ``batch_size`` chunk of the reference's candidate list, the new unique
normalised strings in first-occurrence order.

What moved to the GPU: the candidate loops and prune predicates (LBF:139-195)
and the removal of exact duplicates *before* the CPU normaliser (the device
marks the first occurrence of every distinct spliced program; duplicates would
normalise to an already-seen signature, LBF:204-210, so skipping them changes
nothing but the CPU time).  The normaliser itself (SymPy) stays on the CPU by
design (BASELINE.json north_star).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

from . import core
from .grids import canonical_slug

UNARY_NAMES = ("neg", "inv", "sqrt", "square", "pow_3_2", "pow_neg_3_2", "exp", "exp_neg")   # EO:80-89
BINARY_NAMES = ("add", "sub", "mul", "div", "geom_sum")                                      # EO:91-97
DEAD_BINARY_NAMES = ("sqrt_shift_neg", "sqrt_shift_pos", "exp_mul", "log_mul")               # EO:99-104, no branch in LBF:170-195


def candidate_string(op: int, a: str, b: Optional[str]) -> str:
    """The reference's string templates (LBF:153, 170-195); operands already swapped."""
    if op < 8:
        return f"{UNARY_NAMES[op]}({a})"
    if op == 8:
        return f"({a} + {b})"
    if op == 9:
        return f"({a} - {b})"
    if op == 10:
        return f"({a} * {b})"
    if op == 11:
        return f"({a} / ({b}))"
    return f"({a} / (1 - {b}))"


class GpuExpressionGenerator:
    def __init__(self, normalizer: Any = None, problem: str = "force_free", L: int = 128,
                 session: Optional[core.Session] = None, pre_batch_hook: Optional[Callable[[int, List[str]], None]] = None,
                 last_depth_filter: Any = None):
        if normalizer is None:
            raise ValueError("GpuExpressionGenerator needs the CPU normaliser (an object with normalize_batch)")
        self.normalizer = normalizer
        self.problem = canonical_slug(problem)
        self.session = session or core.Session.for_problem(self.problem)
        self.L = L
        self.pre_batch_hook = pre_batch_hook
        # OPT-IN, not a drop-in: a GpuBatchValidator whose filter is applied to the RAW candidates of the last
        # depth while they are still resident on the device, so that only survivors reach the CPU normaliser
        # (candidates of the final depth are operands of nothing).  Rows the device rejects then never reach
        # on_batch / the run database; the default (None) keeps the reference's stream bit for bit.
        self.last_depth_filter = last_depth_filter
        self.seen_normalized = set()
        self.stats: Dict[int, Dict[str, int]] = {}
        self.last_enumeration: Dict[int, dict] = {}

    # -- LBF:82-111 ---------------------------------------------------------
    def generate_expressions(self, primitives: List, unary_ops: Dict, binary_ops: Dict, max_depth: int) -> Dict[int, List[str]]:
        print(f"\nGenerating expressions with Lean normalization up to depth {max_depth}")
        expressions_by_depth: Dict[int, List[str]] = {}

        def collect_batch(depth: int, expressions: List[str]):
            # NB the reference overwrites per batch (LBF:99-100); kept bit-compatible
            expressions_by_depth[depth] = expressions

        self.stream_generate(primitives=primitives, unary_ops=unary_ops, binary_ops=binary_ops,
                             max_depth=max_depth, on_batch=collect_batch)
        return expressions_by_depth

    @staticmethod
    def _check_vocabulary(unary_ops: Dict, binary_ops: Dict) -> None:
        un = tuple(unary_ops)
        live = tuple(n for n in binary_ops if n not in DEAD_BINARY_NAMES)
        if un != UNARY_NAMES or live != BINARY_NAMES:
            raise ValueError(
                "GpuExpressionGenerator implements the reference's op vocabulary "
                f"{UNARY_NAMES} / {BINARY_NAMES} in that order; got {un} / {live}")

    def enumerate_depth(self, E: Dict[int, List[str]], depth: int, prune: bool = True, keep_device: bool = False) -> dict:
        """Device enumeration of depth `depth` from E[1..depth-1]: triples, first-occurrence flags."""
        import torch
        all_strs: List[str] = []
        depth_begin = [0]
        for k in range(1, depth):
            all_strs.extend(E[k])
            depth_begin.append(len(all_strs))
        exprs = self.session.compile(all_strs)
        n = core.enumerate_count(exprs, depth_begin, depth, prune)
        # CSR rows: the programs packed in one byte pool (a third of the padded rows; include/pde_b200.h)
        dev = core.enumerate_candidates_csr(exprs, depth_begin, depth, prune, 0, n, self.L)
        first, n_unique = core.dedup_csr(dev["pool"], dev["off"], dev["len"], dev["hash"])
        out = {
            "n": n, "n_unique_programs": n_unique, "strings": all_strs, "depth_begin": depth_begin,
            "triple": dev["triple"].cpu().numpy(), "first": first.cpu().numpy().astype(bool),
            "n_uncompiled": int((dev["len"] == 0).sum().item()),
        }
        if keep_device:
            out["device"] = dev
            out["device_first"] = first
        torch.cuda.synchronize()
        return out

    # -- LBF:113-215 --------------------------------------------------------
    def stream_generate(self, primitives: List, unary_ops: Dict, binary_ops: Dict, max_depth: int,
                        batch_size: int = 1000, on_batch: Optional[Any] = None, prune: bool = True) -> None:
        self._check_vocabulary(unary_ops, binary_ops)
        primitive_strs = [str(p) for p in primitives]
        if on_batch:
            on_batch(1, list(primitive_strs))
        expressions_by_depth: Dict[int, List[str]] = {1: primitive_strs}
        seen_signatures = set()
        for depth in range(2, max_depth + 1):
            filt = self.last_depth_filter if depth == max_depth else None
            enum = self.enumerate_depth(expressions_by_depth, depth, prune, keep_device=filt is not None)
            n, triple, first, strs = enum["n"], enum["triple"], enum["first"], enum["strings"]
            n_filtered = 0
            if filt is not None and n > 0:
                # stage 1 -> stage 2 on the device: the spliced programs are validated where they were written; under
                # torch.distributed every rank takes a window of the index space (GpuBatchValidator.filter_enumerated)
                dev = enum.pop("device")
                surv = filt.filter_enumerated(strs, enum["depth_begin"], depth, prune, self.L, cand=dev,
                                              first_flags=enum.pop("device_first"), session=self.session)
                n_filtered = int((first & ~surv).sum())
                first = first & surv
                del dev
            print(f"Depth {depth}: {n} candidates to normalize")
            unique_expressions: List[str] = []
            n_normalized = 0
            for i in range(0, n, batch_size):
                idx = np.nonzero(first[i:i + batch_size])[0] + i
                batch = []
                for c in idx:
                    op, a, b = (int(v) for v in triple[c])
                    batch.append((candidate_string(op, strs[a], strs[b] if b >= 0 else None), depth))
                n_normalized += len(batch)
                results = self.normalizer.normalize_batch(batch) if batch else []
                out_chunk: List[str] = []
                for result in results:
                    sig = result.get("signature")
                    norm = result.get("normalized")
                    if sig not in seen_signatures:
                        seen_signatures.add(sig)
                        unique_expressions.append(norm)
                        out_chunk.append(norm)
                if on_batch and out_chunk:
                    if self.pre_batch_hook is not None:
                        self.pre_batch_hook(depth, out_chunk)
                    on_batch(depth, out_chunk)
            expressions_by_depth[depth] = unique_expressions
            self.stats[depth] = {"candidates": n, "normalized": n_normalized, "uniques": len(unique_expressions),
                                 "exact_duplicates_dropped_on_device": n - n_normalized - n_filtered,
                                 "rejected_on_device_before_normalisation": n_filtered}
            print(f"Depth {depth}: {len(unique_expressions)} unique expressions after normalization")
        self.expressions_by_depth = expressions_by_depth

"""Multi-GPU sharding of the hot path.

Candidates are independent units: the candidate index space [0, N) of a depth
is cut into `world` contiguous ranges that follow the reference's order, each
rank enumerates / validates its own range against a replicated grid and
residual program, and the ONLY exchange is a final gather of survivor bitmasks
and 64-bit structural hashes to rank 0 (torch.distributed: NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests).  No data-path collective.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [first, first+count) of rank `rank`; ranges tile [0, n) in order and
    start on multiples of 32 so survivor bitmask words never straddle ranks."""
    words = (n + 31) // 32
    base, rem = divmod(words, world)
    w0 = rank * base + min(rank, rem)
    w1 = w0 + base + (1 if rank < rem else 0)
    first = min(w0 * 32, n)
    last = min(w1 * 32, n)
    return first, last - first


def gather_survivors(bits, hashes, n_local: int, lens=None, group=None, dst: int = 0):
    """Gather (survivor bitmask words, structural hashes[, program lengths]) of every rank on `dst`.

    bits   int32 [(n_local+31)//32]  (pde_validate survivor_bits)
    hashes int64 [n_local]
    lens   uint8 [n_local] or None   (pde_enumerate len: 0 = not compiled on the device)
    Returns on dst: (list of per-rank bit tensors, list of per-rank hash tensors, list of n_local, list of per-rank
    length tensors or None); elsewhere None.  Shards may have different sizes: sizes are exchanged first, payloads
    are padded to the maximum."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [bits], [hashes], [n_local], (None if lens is None else [lens])
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = bits.device
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n_local], dtype=torch.int64, device=dev), group=group)
    ns = [int(s.item()) for s in sizes]
    nmax = max(ns)
    wmax = (nmax + 31) // 32

    def gather_padded(t, size):
        pad = torch.zeros(size, dtype=t.dtype, device=dev)
        pad[:t.numel()] = t
        got = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, got, dst=dst, group=group)
        return got

    gb = gather_padded(bits, wmax)
    gh = gather_padded(hashes, nmax)
    gl = gather_padded(lens, nmax) if lens is not None else None
    if rank != dst:
        return None
    return ([g[:(n + 31) // 32] for g, n in zip(gb, ns)], [g[:n] for g, n in zip(gh, ns)], ns,
            None if gl is None else [g[:n] for g, n in zip(gl, ns)])


def merge_survivors(gathered, fetch_programs: Optional[Callable[[Sequence[int]], Sequence[bytes]]] = None) -> Tuple[List[int], List[int]]:
    """Rank-0 merge: global survivor indices (in the reference's candidate order) and their hashes, with
    cross-shard EXACT duplicates removed (the lowest index stays, like LBF:204-210 keeps the first occurrence).

    Same contract as pde_dedup (include/pde_b200.h): a 64-bit hash match is only a HINT.  Two survivors are merged
    only when their programs are byte-identical -- `fetch_programs(indices)` returns the program bytes of the given
    global candidate indices (`enumerated_program_fetcher`: rank 0 re-enumerates them, stage 1 is a pure function of
    the index).  Candidates the device could not compile (len == 0: their hash is the constant hash of the empty row)
    are never merged, and without `fetch_programs` nothing is merged at all: a collision must never drop a candidate."""
    import numpy as np
    bits_l, hash_l, ns = gathered[0], gathered[1], gathered[2]
    len_l = gathered[3] if len(gathered) > 3 else None
    surv_idx, surv_hash, surv_len = [], [], []
    off = 0
    for r, (bits, hashes, n) in enumerate(zip(bits_l, hash_l, ns)):
        b = bits.cpu().numpy().view(np.uint32)
        k = np.arange(n)
        surv = ((b[k >> 5] >> (k & 31).astype(np.uint32)) & 1).astype(bool)
        i = np.flatnonzero(surv)
        surv_idx.append(i + off)
        surv_hash.append(hashes.cpu().numpy()[i])
        surv_len.append(len_l[r].cpu().numpy()[i].astype(np.int64) if len_l is not None else np.full(len(i), -1, np.int64))
        off += n
    idx = np.concatenate(surv_idx) if surv_idx else np.zeros(0, np.int64)
    hs = np.concatenate(surv_hash) if surv_hash else np.zeros(0, np.int64)
    ln = np.concatenate(surv_len) if surv_len else np.zeros(0, np.int64)
    keep = np.ones(len(idx), bool)
    if fetch_programs is not None and len(idx):
        # groups of survivors with equal (hash, len), len != 0: candidates for a merge, confirmed byte-wise
        order = np.lexsort((idx, ln, hs))
        same = (hs[order][1:] == hs[order][:-1]) & (ln[order][1:] == ln[order][:-1]) & (ln[order][1:] != 0)
        member = np.zeros(len(idx), bool)
        member[order[1:][same]] = True
        member[order[:-1][same]] = True
        todo = np.flatnonzero(member)
        if len(todo):
            progs = dict(zip((int(idx[t]) for t in todo), fetch_programs([int(idx[t]) for t in todo])))
            firsts: dict = {}
            for t in sorted(todo, key=lambda t: int(idx[t])):
                key = (int(hs[t]), int(ln[t]))
                mine = bytes(progs[int(idx[t])])
                seen = firsts.setdefault(key, [])
                if any(mine == other for other in seen):
                    keep[t] = False                    # byte-identical to an earlier survivor
                else:
                    seen.append(mine)                  # first of its program (or a true 64-bit collision: kept)
    return [int(i) for i in idx[keep]], [int(h) for h in hs[keep]]


def enumerated_program_fetcher(exprs, depth_begin, depth: int, prune: bool = True, L: int = 128, window: int = 1 << 22):
    """fetch_programs for `merge_survivors`: re-enumerate the requested candidates on THIS rank's GPU (stage 1 is a
    pure function of the candidate index: pde_enumerate produces any window) and return their program bytes."""
    from . import core

    def fetch(indices: Sequence[int]) -> List[bytes]:
        import numpy as np
        import torch
        want = np.asarray(list(indices), dtype=np.int64)
        out: dict = {}
        for lo in range(int(want.min()) // window * window, int(want.max()) + 1, window):
            sel = want[(want >= lo) & (want < lo + window)]
            if not len(sel):
                continue
            first, last = int(sel.min()), int(sel.max()) + 1
            c = core.enumerate_candidates(exprs, depth_begin, depth, prune, first, last - first, L)
            rows = c["code"][torch.from_numpy(sel - first).to(c["code"].device)].cpu().numpy()
            lens = c["len"][torch.from_numpy(sel - first).to(c["len"].device)].cpu().numpy()
            for g, row, n in zip(sel, rows, lens):
                out[int(g)] = bytes(row[:int(n)])
        return [out[int(g)] for g in want]

    return fetch

"""Multi-GPU sharding of the hot path.

Candidates are independent units: the candidate index space [0, N) of a depth
is cut into `world` contiguous ranges that follow the reference's order, each
rank enumerates / validates its own range against a replicated grid and
residual program, and the ONLY exchange is a final gather of survivor bitmasks
and 64-bit structural hashes to rank 0 (torch.distributed: NCCL over
NVLink/NVSwitch on GPUs, gloo in the CPU tests).  No data-path collective.
"""
from __future__ import annotations

from typing import List, Optional, Tuple


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


def gather_survivors(bits, hashes, n_local: int, group=None, dst: int = 0):
    """Gather (survivor bitmask words, structural hashes) of every rank on `dst`.

    bits   int32 [(n_local+31)//32]  (pde_validate survivor_bits)
    hashes int64 [n_local]
    Returns on dst: (list of per-rank bit tensors, list of per-rank hash tensors,
    list of n_local); elsewhere None.  Shards may have different sizes: sizes are
    exchanged first, payloads are padded to the maximum."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return [bits], [hashes], [n_local]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = bits.device
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([n_local], dtype=torch.int64, device=dev), group=group)
    ns = [int(s.item()) for s in sizes]
    nmax = max(ns)
    wmax = (nmax + 31) // 32
    pb = torch.zeros(wmax, dtype=bits.dtype, device=dev)
    pb[:bits.numel()] = bits
    ph = torch.zeros(nmax, dtype=hashes.dtype, device=dev)
    ph[:hashes.numel()] = hashes
    if rank == dst:
        gb = [torch.empty_like(pb) for _ in range(world)]
        gh = [torch.empty_like(ph) for _ in range(world)]
    else:
        gb = gh = None
    dist.gather(pb, gb, dst=dst, group=group)
    dist.gather(ph, gh, dst=dst, group=group)
    if rank != dst:
        return None
    return ([g[:(n + 31) // 32] for g, n in zip(gb, ns)], [g[:n] for g, n in zip(gh, ns)], ns)


def merge_survivors(gathered) -> Tuple[List[int], List[int]]:
    """Rank-0 merge: global survivor indices (in the reference's candidate order) and
    their hashes; cross-shard exact duplicates (same hash) keep the lowest index."""
    import numpy as np
    bits_l, hash_l, ns = gathered
    idx: List[int] = []
    hs: List[int] = []
    off = 0
    seen = set()
    for bits, hashes, n in zip(bits_l, hash_l, ns):
        b = bits.cpu().numpy().view(np.uint32)
        h = hashes.cpu().numpy()
        k = np.arange(n)
        surv = ((b[k >> 5] >> (k & 31).astype(np.uint32)) & 1).astype(bool)
        for i in np.nonzero(surv)[0]:
            hv = int(h[i])
            if hv in seen:
                continue
            seen.add(hv)
            idx.append(off + int(i))
            hs.append(hv)
        off += n
    return idx, hs

"""Sharding of a batch over ranks.

Strings are independent and the automaton tables are tiny, so a batch is cut into
contiguous string-index ranges -- balanced by BYTES, not by count -- one per rank;
each rank matches its range on its own GPU and writes its slice of the result
vector.  There is no data-path collective; the only cross-rank traffic is the final
gather of result bits / the sum of match counts, done here with torch.distributed
(NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_by_count(n_total: int, rank: int, world: int):
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def shard_by_bytes(offsets: np.ndarray, world: int):
    """offsets: uint64[n+1].  Returns world+1 string indices b[0]=0 <= ... <= b[world]=n such
    that rank r owns strings [b[r], b[r+1]) and every rank gets ~total/world bytes."""
    n = len(offsets) - 1
    total = int(offsets[-1]) - int(offsets[0])
    bounds = [0]
    for r in range(1, world):
        target = int(offsets[0]) + (total * r) // world
        k = int(np.searchsorted(offsets, target, side="left"))
        k = max(bounds[-1], min(n, k))
        bounds.append(k)
    bounds.append(n)
    return bounds


def local_view(chars: np.ndarray, offsets: np.ndarray, lo: int, hi: int):
    """The shard's own (chars, offsets) with offsets rebased to 0."""
    o = offsets[lo:hi + 1]
    return chars[int(o[0]):int(o[-1])], (o - o[0]).astype(np.uint64)


def gather_bits(local_bits, bounds, rank, world, dist, device="cpu"):
    """All ranks end up with the full result vector (uint8[n])."""
    import torch
    n = bounds[-1]
    sizes = [bounds[r + 1] - bounds[r] for r in range(world)]
    mx = max(sizes) if sizes else 0
    buf = torch.zeros(mx, dtype=torch.uint8, device=device)
    buf[:sizes[rank]] = torch.as_tensor(local_bits, dtype=torch.uint8, device=device)
    parts = [torch.zeros(mx, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = torch.empty(n, dtype=torch.uint8, device=device)
    for r in range(world):
        out[bounds[r]:bounds[r + 1]] = parts[r][:sizes[r]]
    return out


def total_matches(local_bits, dist, device="cpu"):
    """Optional final reduction: number of matching strings over all ranks (8-byte message)."""
    import torch
    t = torch.tensor([int(np.asarray(local_bits, dtype=np.uint64).sum())], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())

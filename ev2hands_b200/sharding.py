"""Host-side sharding of a batch of event windows over ranks (one process per GPU).

Windows are independent in eval mode (BatchNorm folded, every kernel is per window), so the
inference path needs no collective: each rank takes a contiguous slice of the batch *and of the
per-window FPS start indices*, which makes the result independent of the number of shards.
`torch.distributed` is used only to agree on timings (max over ranks) and, in training, for the
gradient all-reduce (the reference's `nn.DataParallel`, train.py:68, reduces gradients the same way).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world: int):
    """Contiguous, balanced split: the first n_items % world ranks get one extra item."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(tensors, rank: int, world: int):
    """Slice every tensor of a tuple along dim 0 with the same bounds."""
    n = tensors[0].shape[0]
    lo, hi = shard_bounds(n, rank, world)
    return tuple(t[lo:hi] for t in tensors)


def shard_windows(starts, counts, rank: int, world: int):
    """Shard a batch of event windows (row ranges of one raw event table, ``EventWindowBuilder``'s input) over ranks.

    -> ``(row_lo, row_hi, local_starts, local_counts, (lo, hi))``: the rank's contiguous share ``[lo, hi)`` of the
    windows, the slice ``events[row_lo:row_hi]`` of the table it has to hold (windows may overlap, so slices of
    neighbouring ranks may too) and the starts rebased to that slice.  Like the FPS start indices, the per-window
    point draws are made for the global batch in batch order and sharded with the same bounds, so the union of the
    ranks' windows equals the unsharded batch.  No collective."""
    import numpy as np
    starts = np.asarray(starts, dtype=np.int64)
    counts = np.asarray(counts, dtype=np.int64)
    if starts.shape != counts.shape or starts.ndim != 1:
        raise ValueError("starts and counts must be equally long 1-d sequences")
    lo, hi = shard_bounds(starts.shape[0], rank, world)
    if hi == lo:
        return 0, 0, starts[:0], counts[:0].astype(np.int32), (lo, hi)
    s, c = starts[lo:hi], counts[lo:hi]
    row_lo, row_hi = int(s.min()), int((s + c).max())
    return row_lo, row_hi, s - row_lo, c.astype(np.int32), (lo, hi)


def max_over_ranks(values, device=None) -> list:
    """Element-wise maximum of a list of floats over all ranks (identity without a process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def allreduce_mean_grads(params) -> int:
    """Average gradients over ranks with ONE flattened all-reduce (training config); returns the
    number of elements reduced."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sum(g.numel() for g in grads)
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return off

"""Multi-GPU: pairs are independent, so the batch is partitioned by pair index over the
ranks (one process per GPU) with no collective on the data path; only the final results
are gathered on rank 0."""
from __future__ import annotations

from typing import Any, List, Tuple


def shard_range(num_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of pair indices owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(num_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(local: List[Any]) -> List[Any]:
    """Concatenate per-rank result lists in rank order on rank 0 (others get [])."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local)
    world = dist.get_world_size()
    out = [None] * world if dist.get_rank() == 0 else None
    dist.gather_object(local, out, dst=0)
    if dist.get_rank() != 0:
        return []
    return [x for part in out for x in part]


def align_batch_sharded(pairs, k, device=None, **kw):
    """Each rank aligns its block on its own GPU; rank 0 returns the list of
    (status, aln0, aln1) for the whole batch."""
    import torch.distributed as dist
    from .batch import default_aligner
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if device is None:
        import os
        device = int(os.environ.get("LOCAL_RANK", rank))
    lo, hi = shard_range(len(pairs), rank, world)
    res = default_aligner(device).align_pairs(pairs[lo:hi], k, **kw)
    local = [(int(res.status[i]),) + res.alignment(i) for i in range(hi - lo)]
    return gather_results(local)

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


_HOST_GROUP = None


def _host_group():
    """A gloo group over all ranks for host-side gathers (the default group may be NCCL, which only moves device
    tensors; the results live in pinned host memory)."""
    global _HOST_GROUP
    import torch.distributed as dist
    if dist.get_backend() == "gloo":
        return None                                  # the default group already is one
    if _HOST_GROUP is None:
        _HOST_GROUP = dist.new_group(backend="gloo")
    return _HOST_GROUP


def gather_batch_result(res):
    """The shards' BatchResults (rank order = pair order) joined into one on rank 0, as arrays: per-pair columns
    are concatenated, the row / run buffers are appended and their offsets rebased.  Nothing per pair is pickled:
    every column travels as one tensor (sizes first, then the payload padded to the largest shard).  Other ranks
    get None.  Graph-mode flags and compact rows are not carried (they are per-shard diagnostics / wire forms)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from .batch import BatchResult
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return res
    world, rank, grp = dist.get_world_size(), dist.get_rank(), _host_group()
    runs = np.ascontiguousarray(res.runs, dtype=np.int32).reshape(-1)
    cols = [("status", np.int32, res.status), ("score", np.int64, res.score), ("dp_cells", np.int64, res.dp_cells),
            ("dp_segments", np.int64, res.dp_segments), ("n_anchors", np.int64, res.n_anchors),
            ("aln_len", np.int64, np.diff(res.aln_off)), ("run_len", np.int64, np.diff(res.run_off)),
            ("runs", np.int32, runs), ("aln0", np.uint8, res.aln0), ("aln1", np.uint8, res.aln1)]
    if res.sop is not None:
        cols.append(("sop", np.int32, np.ascontiguousarray(res.sop).reshape(-1)))
    sizes = torch.tensor([len(a) for _, _, a in cols], dtype=torch.int64)
    all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=grp)
    all_sizes = torch.stack(all_sizes).numpy()                           # [world, columns]
    parts = {}
    for c, (name, dt, a) in enumerate(cols):
        n_max = int(all_sizes[:, c].max())
        buf = np.zeros(max(n_max, 1), dt)
        buf[:len(a)] = np.asarray(a, dtype=dt)
        t = torch.from_numpy(buf)
        out = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, out, dst=0, group=grp)
        if rank == 0:
            parts[name] = np.concatenate([out[r].numpy()[:int(all_sizes[r, c])] for r in range(world)])
    if rank != 0:
        return None
    off = lambda lens: np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return BatchResult(status=parts["status"], score=parts["score"], dp_cells=parts["dp_cells"],
                       dp_segments=parts["dp_segments"], n_anchors=parts["n_anchors"], run_off=off(parts["run_len"]),
                       runs=parts["runs"].reshape(-1, 3), aln_off=off(parts["aln_len"]), aln0=parts["aln0"], aln1=parts["aln1"],
                       sop=parts["sop"].reshape(-1, 4) if "sop" in parts else None)


def align_batch_sharded(pairs, k, device=None, **kw):
    """Each rank aligns its block on its own GPU; rank 0 returns the list of
    (status, aln0, aln1) for the whole batch.  (Convenience form for small batches: it pickles a tuple per pair.
    For real batches gather the arrays: align_buffers_sharded / gather_batch_result.)"""
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


def align_buffers_sharded(aligner, seqs, offsets, prm):
    """dbga_align_batch on this rank's block of pairs (by index, shard_range) of the batch (seqs, offsets[2P+1]);
    rank 0 returns the whole batch's BatchResult (gather_batch_result), the others None."""
    import numpy as np
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    P = (len(offsets) - 1) // 2
    lo, hi = shard_range(P, rank, world)
    base = int(offsets[2 * lo])
    sub_offs = np.ascontiguousarray(offsets[2 * lo:2 * hi + 1] - base)
    res = aligner.align_buffers(np.ascontiguousarray(seqs[base:int(offsets[2 * hi])]), sub_offs, prm)
    return gather_batch_result(res)

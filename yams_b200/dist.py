"""Multi-GPU plumbing for the row-sharded scan: one process per GPU, rows sharded by range,
one all-gather of the per-shard partial top-k, merge on the device.  torch.distributed is only the transport
(NCCL on GPUs; gloo in the CPU tests)."""
from __future__ import annotations


def shard_rows(n_total: int, rank: int, world: int):
    """Contiguous row range of `rank`: (first_row, n_rows). Earlier ranks take the remainder rows."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allgather_partials(part_rowids, part_scores):
    """[Q,k] int64 / float32 partial top-k of this rank -> ([R,Q,k], [R,Q,k]) in rank order, the layout
    yams_b200_merge_partials_device consumes. Works for CUDA (NCCL) and CPU (gloo) tensors."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    all_r = torch.empty((world,) + tuple(part_rowids.shape), dtype=part_rowids.dtype, device=part_rowids.device)
    all_s = torch.empty((world,) + tuple(part_scores.shape), dtype=part_scores.dtype, device=part_scores.device)
    try:
        dist.all_gather_into_tensor(all_r, part_rowids.contiguous())
        dist.all_gather_into_tensor(all_s, part_scores.contiguous())
    except (RuntimeError, NotImplementedError):
        lr = [torch.empty_like(part_rowids) for _ in range(world)]
        ls = [torch.empty_like(part_scores) for _ in range(world)]
        dist.all_gather(lr, part_rowids.contiguous())
        dist.all_gather(ls, part_scores.contiguous())
        all_r, all_s = torch.stack(lr), torch.stack(ls)
    return all_r, all_s


def split_allowed(allowed, rowid_lo: int, rowid_hi: int):
    """Candidate sets on a row-sharded corpus: every rank keeps, of each query's ascending
    allowed-rowid list, the part that falls into its own rowid range [rowid_lo, rowid_hi).  The per-rank partial
    top-k are then gathered and merged exactly like the unfiltered scan."""
    import numpy as np
    out = []
    for a in allowed:
        a = np.asarray(a, dtype=np.int64)
        lo = int(np.searchsorted(a, rowid_lo, side="left"))
        hi = int(np.searchsorted(a, rowid_hi, side="left"))
        out.append(a[lo:hi])
    return out

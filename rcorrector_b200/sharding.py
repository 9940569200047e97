"""Multi-GPU plumbing: reads shard embarrassingly, the table is replicated.

* shard_bounds(): contiguous record-index ranges, both mates of a pair on the same rank.
* broadcast_table(): one torch.distributed broadcast of the raw (code, count) arrays from rank 0 (NCCL over NVLink on
  the GPU box, gloo in the CPU tests); every rank then inserts locally with rcx_table_build(src_on_device=1).
There is no collective on the data path.
"""
import numpy as np


def shard_bounds(n_records, world, rank, interleaved):
    """[lo, hi) of `rank`; in interleaved/paired mode the split points are even so mates 2i, 2i+1 stay together."""
    unit = 2 if interleaved else 1
    units = n_records // unit
    lo = units * rank // world * unit
    hi = units * (rank + 1) // world * unit
    if rank == world - 1:
        hi = n_records  # a dangling odd record (malformed interleaved input) stays with the last rank
    return lo, hi


def shard_batch(seq, qual, off, lo, hi):
    """Views of one shard with offsets rebased to 0."""
    o = off[lo:hi + 1].astype(np.int64)
    return seq[o[0]:o[-1]], qual[o[0]:o[-1]], (o - o[0]).astype(np.uint32)


def broadcast_table(codes, counts, device=None, src=0):
    """Broadcast (codes:int64 view of uint64, counts:int32) tensors from `src`; returns tensors on `device`."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    n = torch.tensor([0 if codes is None else len(codes)], dtype=torch.int64, device=device)
    dist.broadcast(n, src)
    n = int(n.item())
    if rank == src:
        t_codes = torch.as_tensor(np.ascontiguousarray(codes).view(np.int64)).to(device) if device else torch.from_numpy(np.ascontiguousarray(codes).view(np.int64).copy())
        t_counts = torch.as_tensor(np.ascontiguousarray(counts)).to(device) if device else torch.from_numpy(np.ascontiguousarray(counts).copy())
    else:
        t_codes = torch.empty(n, dtype=torch.int64, device=device)
        t_counts = torch.empty(n, dtype=torch.int32, device=device)
    dist.broadcast(t_codes, src)
    dist.broadcast(t_counts, src)
    return t_codes, t_counts

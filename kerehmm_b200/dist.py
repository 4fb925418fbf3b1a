"""Multi-GPU plumbing: one process per GPU (torchrun), sequences sharded across ranks, and ONE
collective per Baum-Welch iteration -- a sum all-reduce of the packed float64 count buffer
(include/khmm.h) over NCCL/NVLink (gloo on CPU in the tests).  Forward / backward / Viterbi need
no communication."""
from __future__ import annotations


def is_initialized() -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def world():
    import torch.distributed as dist
    if is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_sequences: int, rank: int = None, world_size: int = None):
    """Contiguous block of sequences owned by `rank`: the first (n mod W) ranks get one extra."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, extra = divmod(n_sequences, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_counts(buf, group=None):
    """In-place sum of the packed count buffer over all ranks (no-op for a single process).
    All-reduce results are bitwise identical on every rank, so the M-step that follows gives
    identical parameters everywhere without a broadcast."""
    if not is_initialized():
        return buf
    import torch.distributed as dist
    if dist.get_world_size(group) == 1:
        return buf
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf

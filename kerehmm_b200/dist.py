"""Multi-GPU plumbing: one process per GPU, sequences sharded across ranks, and ONE collective per
Baum-Welch iteration -- a sum all-reduce of the packed float64 count buffer (include/khmm.h).
Forward / backward / Viterbi need no communication.

The collective itself is libkhmm's: `khmm_comm_init_rank` / `khmm_allreduce_counts` (NCCL over
NVLink, bound by the library at run time), so a caller of the C ABI without PyTorch has the same
multi-GPU path.  Only the 128-byte rendezvous id travels over a side channel; here that is
`torch.distributed` (any backend) when it is initialised, or an explicit `id_bytes` argument.
On CPU tensors (the world_size-2 `gloo` tests) the sum falls back to `torch.distributed.all_reduce`
-- that is host-logic test plumbing, never the GPU product path.
"""
from __future__ import annotations

import ctypes

from . import _lib

_comm = None          # (handle, rank, world, device)
LOCAL = object()      # `group=LOCAL`: no collective, the statistics of this process's batch only


def is_initialized() -> bool:
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def world():
    if _comm is not None:
        return _comm[1], _comm[2]
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


def bind_to_device_cpus(device: int = None):
    """Pin this process to the CPU cores NVML reports as local to `device` (its NUMA node), so that page-locked host
    buffers allocated afterwards -- first touch -- sit next to the GPU's PCIe root.  With 4-8 ranks per box the pinned
    host -> device copy of an unbound process ran at half rate (2.4 -> 4.6 ms for 134 MB, round-1 scaling run): half
    the ranks' buffers were on the other socket.  Returns the core list, or None when NVML / the affinity call is
    unavailable (nothing is changed then).  Call it before allocating pinned memory."""
    import os
    try:
        import pynvml
        import torch
        if device is None:
            device = torch.cuda.current_device()
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(int(device))
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:          # no NVML, no permission, a container that hides the topology: leave the process as it is
        return None


def unique_id() -> bytes:
    """A fresh NCCL rendezvous id (rank 0 calls this and ships the 128 bytes to the other ranks)."""
    buf = ctypes.create_string_buffer(128)
    _lib.call("khmm_comm_unique_id", buf)
    return buf.raw


def init_comm(rank: int = None, world_size: int = None, device: int = None, id_bytes: bytes = None):
    """Create this process's libkhmm communicator (collective over all ranks).  With torch.distributed
    initialised, rank / world default to its values and rank 0's id is broadcast through it."""
    global _comm
    if _comm is not None:
        return _comm
    import torch
    if rank is None or world_size is None:
        if not is_initialized():
            raise RuntimeError("init_comm needs rank/world_size, or an initialised torch.distributed group")
        import torch.distributed as dist
        rank, world_size = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = torch.cuda.current_device()
    if id_bytes is None:
        if world_size == 1:
            id_bytes = unique_id()
        else:
            import torch.distributed as dist
            box = [unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            id_bytes = box[0]
    if len(id_bytes) != 128:
        raise ValueError("the rendezvous id is 128 bytes")
    handle = ctypes.c_void_p()
    _lib.call("khmm_comm_init_rank", ctypes.create_string_buffer(id_bytes, 128), int(rank), int(world_size), int(device),
              ctypes.byref(handle))
    _comm = (handle, int(rank), int(world_size), int(device))
    return _comm


def comm_info():
    """(rank, world, rank count NCCL itself reports) of the libkhmm communicator, or None."""
    if _comm is None:
        return None
    r, w, n = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.call("khmm_comm_info", _comm[0], ctypes.byref(r), ctypes.byref(w), ctypes.byref(n))
    return r.value, w.value, n.value


def destroy_comm():
    global _comm
    if _comm is not None:
        _lib.call("khmm_comm_destroy", _comm[0])
        _comm = None


def all_reduce_counts(buf, group=None):
    """In-place sum of the packed count buffer over all ranks (no-op for a single process).
    All-reduce results are bitwise identical on every rank, so the M-step that follows gives
    identical parameters everywhere without a broadcast.  CUDA buffers go through libkhmm's own
    communicator (created on first use from the default torch.distributed group); CPU buffers (tests) and
    explicit sub-groups go through torch.distributed."""
    if group is LOCAL:
        return buf
    if buf.is_cuda and group is None:
        if _comm is None:
            if not is_initialized():
                return buf
            import torch.distributed as dist
            if dist.get_world_size(group) == 1:
                return buf
            init_comm()
        if _comm[2] > 1:
            import torch
            _lib.call("khmm_allreduce_counts", _comm[0], buf.data_ptr(), buf.numel(),
                      torch.cuda.current_stream().cuda_stream)
        return buf
    if not is_initialized():
        return buf
    import torch.distributed as dist
    if dist.get_world_size(group) == 1:
        return buf
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf

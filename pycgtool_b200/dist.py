"""Frame sharding across GPUs.

Frames are independent in mapping and measurement (``pycgtool/mapping.py:446-463``
and the mdtraj kernels work frame by frame); only the per-bond statistics are
global.  So a trajectory is split into contiguous frame blocks, one per rank
(one process per GPU), each rank maps and measures its block with no
communication, and the packed float64 moments / histogram counts are summed
with one NCCL all-reduce over NVLink before the epilogue.  ``torch.distributed``
is the plumbing; with a single process every function here is a no-op.
"""

import torch
import torch.distributed as td


def is_distributed():
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def world():
    """(rank, world_size)."""
    if is_distributed():
        return td.get_rank(), td.get_world_size()
    return 0, 1


def frame_shard(n_frames, rank=None, world_size=None):
    """Contiguous block ``[begin, end)`` of frames owned by ``rank``.

    Blocks are in rank order, so concatenating the shards' ``values`` keeps the
    reference's frame-major layout; the first ``n_frames % world`` ranks get one
    extra frame.
    """
    if rank is None or world_size is None:
        rank, world_size = world()
    base, extra = divmod(n_frames, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_sum(tensor):
    """In-place sum over all ranks (NCCL for CUDA tensors, gloo for CPU tensors)."""
    if is_distributed():
        td.all_reduce(tensor, op=td.ReduceOp.SUM)
    return tensor


def broadcast(tensor, src=0):
    if is_distributed():
        td.broadcast(tensor, src=src)
    return tensor


def all_true(flag, device=None):
    """Logical AND of a Python bool over all ranks (e.g. "every box length is non-zero",
    ``mapping.py:406-408``, which the reference evaluates over the whole trajectory)."""
    if not is_distributed():
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    td.all_reduce(t, op=td.ReduceOp.MIN)
    return bool(t.item())


def merge_partials(shards):
    """Sum a list of per-shard partial tensors (what the all-reduce computes); used by tests
    and by single-process multi-shard runs."""
    total = shards[0].clone()
    for part in shards[1:]:
        total += part
    return total

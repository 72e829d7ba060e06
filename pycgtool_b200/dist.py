"""Frame sharding across GPUs.

Frames are independent in mapping and measurement (``pycgtool/mapping.py:446-463``
and the mdtraj kernels work frame by frame); only the per-bond statistics are
global.  So a trajectory is split into contiguous frame blocks, one per rank
(one process per GPU), each rank maps and measures its block with no
communication, and the packed float64 buffer ``[moments | histogram counts]`` is
summed with ONE NCCL all-reduce over NVLink before the epilogue.

The collective itself is issued by ``libpycgtool_b200.so`` (``cgb_allreduce_partials`` /
``cgb_comm_broadcast`` on a communicator the library creates with ``cgb_comm_init``) on the
stream the kernels run on -- no host synchronisation between measurement, reduction and
epilogue.  ``torch.distributed`` is only the rendezvous that hands the 128-byte NCCL id to the
other ranks, and the transport for host tensors (gloo, the CPU tests).  With a single process
every function here is a no-op.
"""

import ctypes

import numpy as np
import torch
import torch.distributed as td

_comm = None          # (handle, rank, world, device index) of the library's NCCL communicator


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


def communicator():
    """The library's NCCL communicator over all ranks (created on first use; collective)."""
    global _comm
    from . import _lib
    rank, size = world()
    dev = torch.cuda.current_device()
    if _comm is not None and _comm[1:] == (rank, size, dev):
        return _comm[0]
    lib = _lib.load()
    ident = np.zeros(_lib.COMM_ID_BYTES, dtype=np.uint8)
    if rank == 0:
        _lib.check(lib.cgb_comm_unique_id(ident.ctypes.data), "cgb_comm_unique_id")
    box = [ident.tobytes()]
    td.broadcast_object_list(box, src=0)        # rendezvous only: 128 bytes
    ident = np.frombuffer(box[0], dtype=np.uint8).copy()
    handle = ctypes.c_void_p()
    _lib.check(lib.cgb_comm_init(ident.ctypes.data, rank, size, ctypes.byref(handle)), "cgb_comm_init")
    _comm = (handle, rank, size, dev)
    return handle


def destroy_communicator():
    global _comm
    if _comm is not None:
        from . import _lib
        _lib.load().cgb_comm_destroy(_comm[0])
        _comm = None


def allreduce_packed(tensor):
    """In-place sum of a contiguous float64 buffer over all ranks.

    Device tensors: one ``ncclAllReduce`` issued by the library on the current stream.  Host tensors
    (the gloo tests): ``torch.distributed``."""
    if not is_distributed():
        return tensor
    if tensor.is_cuda:
        from . import _lib, device
        if tensor.dtype != torch.float64 or not tensor.is_contiguous():
            raise ValueError("allreduce_packed needs a contiguous float64 tensor")
        _lib.check(_lib.load().cgb_allreduce_partials(communicator(), tensor.data_ptr(), tensor.numel(),
                                                      device.stream_handle()), "cgb_allreduce_partials")
    else:
        td.all_reduce(tensor, op=td.ReduceOp.SUM)
    return tensor


def allreduce_sum(tensor):
    """In-place sum over all ranks through ``torch.distributed`` (any dtype; host-side bookkeeping)."""
    if is_distributed():
        td.all_reduce(tensor, op=td.ReduceOp.SUM)
    return tensor


def broadcast_device(tensor, src=0):
    """Broadcast a contiguous tensor from ``src`` (device tensors through the library, on the current stream)."""
    if not is_distributed():
        return tensor
    if tensor.is_cuda:
        from . import _lib, device
        if not tensor.is_contiguous():
            raise ValueError("broadcast_device needs a contiguous tensor")
        _lib.check(_lib.load().cgb_comm_broadcast(communicator(), tensor.data_ptr(),
                                                  tensor.numel() * tensor.element_size(), int(src),
                                                  device.stream_handle()), "cgb_comm_broadcast")
    else:
        td.broadcast(tensor, src=src)
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


def all_true_many(flags, device=None):
    """Element-wise logical AND of a list of Python bools over all ranks, in ONE collective."""
    flags = tuple(bool(f) for f in flags)
    if not is_distributed():
        return flags
    t = torch.tensor([1 if f else 0 for f in flags], dtype=torch.int32, device=device)
    td.all_reduce(t, op=td.ReduceOp.MIN)
    return tuple(bool(v) for v in t.tolist())


def merge_partials(shards):
    """Sum a list of per-shard partial tensors (what the all-reduce computes); used by tests
    and by single-process multi-shard runs."""
    total = shards[0].clone()
    for part in shards[1:]:
        total += part
    return total

"""Device plumbing: PyTorch as buffer and stream container, nothing more.

All arithmetic happens in ``libpycgtool_b200.so``; torch allocates HBM, moves
bytes and owns the CUDA stream the kernels are enqueued on.
"""

import numpy as np
import torch


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("pycgtool_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def default_device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def stream_handle():
    """Raw ``cudaStream_t`` of torch's current stream, as an int for ctypes."""
    return torch.cuda.current_stream().cuda_stream


def to_device(array, device=None, dtype=None):
    """numpy array (or tensor) -> contiguous tensor on the device."""
    require_cuda()
    device = device or default_device()
    if isinstance(array, torch.Tensor):
        t = array.to(device=device, non_blocking=True)
    else:
        array = np.ascontiguousarray(array, dtype=dtype)
        t = torch.from_numpy(array).to(device=device, non_blocking=True)
    return t.contiguous()


def empty(shape, dtype, device=None):
    require_cuda()
    return torch.empty(shape, dtype=dtype, device=device or default_device())


def zeros(shape, dtype, device=None):
    require_cuda()
    return torch.zeros(shape, dtype=dtype, device=device or default_device())


def dptr(tensor):
    """Device pointer of a contiguous tensor (None passes through as NULL)."""
    if tensor is None:
        return None
    if not tensor.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return tensor.data_ptr()

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


def bind_host_to_gpu(index=None):
    """Pin the calling process to the CPU cores NVML reports as local to the GPU (its NUMA node), so that
    pinned staging buffers allocated afterwards land in memory next to the GPU's PCIe root.  With one
    process per GPU this keeps eight concurrent host<->device streams off a single memory controller.
    Best effort: returns the CPU set, or None when NVML is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        index = torch.cuda.current_device() if index is None else index
        props = torch.cuda.get_device_properties(index)
        try:
            bus = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except (AttributeError, pynvml.NVMLError):
            handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return sorted(os.sched_getaffinity(0))
    except Exception:                      # noqa: BLE001 -- purely an optimisation
        return None

"""Host <-> device plumbing for the plug-in layer (PyTorch owns memory and streams)."""
from __future__ import annotations

import os
from typing import Optional, Union

import numpy as np
import torch

ArrayLike = Union[np.ndarray, torch.Tensor]

# bytes moved by the most recent to_device / to_host calls (bench bookkeeping)
h2d_bytes = 0
d2h_bytes = 0


# optional trace of the pipelined host path: a list that receives (label, CUDA event or None, host seconds
# since the call started) tuples; None = off
trace = None


def reset_counters() -> None:
    global h2d_bytes, d2h_bytes
    h2d_bytes = 0
    d2h_bytes = 0


def is_device(x) -> bool:
    return isinstance(x, torch.Tensor) and x.is_cuda


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("decode_tonal_langauge_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def bind_host_to_gpu(index: Optional[int] = None) -> Optional[list]:
    """Pin this process to the CPU cores NVML reports as local to the GPU (its NUMA node), so that
    the pinned staging buffers it allocates afterwards are first-touched next to the GPU's PCIe
    root and the copies of several ranks do not cross the socket interconnect.  Best effort:
    returns the core list, or None when NVML / sched_setaffinity is unavailable."""
    try:
        import pynvml
        if index is None:
            index = torch.cuda.current_device()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(visible.split(",")[index]) if visible and all(v.strip().isdigit() for v in visible.split(",")) else index
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        pass
    return None


def to_device(x: ArrayLike, dtype: Optional[torch.dtype] = torch.float32) -> torch.Tensor:
    """numpy / CPU tensor -> CUDA tensor.  float64 sources are narrowed on the DEVICE so the
    PCIe copy is the only host-side pass; pinned sources are copied asynchronously."""
    global h2d_bytes
    if is_device(x):
        return x if dtype is None or x.dtype == dtype else x.to(dtype)
    if isinstance(x, np.ndarray):
        if not x.flags.c_contiguous:
            x = np.ascontiguousarray(x)
        if not x.flags.writeable:
            x = x.copy()
        t = torch.from_numpy(x)
    elif isinstance(x, torch.Tensor):
        t = x.contiguous()
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    h2d_bytes += t.numel() * t.element_size()
    d = t.to(device(), non_blocking=t.is_pinned())
    if dtype is not None and d.dtype != dtype:
        d = d.to(dtype)
    return d


def to_host(t: torch.Tensor, dtype: Optional[np.dtype] = None) -> np.ndarray:
    """CUDA tensor -> numpy array (via a pinned buffer from torch's caching host allocator).
    The cast to the reference's dtype happens on the device, before the copy."""
    global d2h_bytes
    if dtype is not None:
        td = getattr(torch, np.dtype(dtype).name)
        if t.dtype != td:
            t = t.to(td)
    t = t.contiguous()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    d2h_bytes += host.numel() * host.element_size()
    return host.numpy()


def output_dtype(reference_dtype) -> np.dtype:
    """dtype handed back for numpy callers: the reference's own convention (SURVEY.md
    Appendix A7) unless ECOG_OUTPUT_DTYPE=float32 asks for the device's storage type."""
    forced = os.environ.get("ECOG_OUTPUT_DTYPE")
    return np.dtype(forced) if forced else np.dtype(reference_dtype)

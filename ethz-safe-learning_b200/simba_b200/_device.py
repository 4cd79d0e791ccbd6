"""Device-memory plumbing: torch is used only for allocation, streams and host<->device copies."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.SimbaError(-3, "no CUDA device: simba_b200 has no CPU fallback")


def ptr(t):
    """c_void_p of a (contiguous) torch tensor, numpy array or None."""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, np.ndarray):
        assert t.flags['C_CONTIGUOUS']
        return C.c_void_p(t.ctypes.data)
    assert t.is_contiguous()
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def to_device(x, dtype=torch.float32, device=None):
    """numpy / torch (any device) -> contiguous CUDA tensor; returns (tensor, kind_of_input)."""
    require_cuda()
    device = device or torch.device('cuda', torch.cuda.current_device())
    if isinstance(x, torch.Tensor):
        kind = 'torch_cuda' if x.is_cuda else 'torch_cpu'
        return x.to(device=device, dtype=dtype).contiguous(), kind
    arr = np.ascontiguousarray(np.asarray(x), dtype={torch.float32: np.float32,
                                                     torch.int32: np.int32}[dtype])
    return torch.from_numpy(arr).to(device), 'numpy'


def like_input(t, kind):
    if kind == 'numpy':
        return t.cpu().numpy()
    if kind == 'torch_cpu':
        return t.cpu()
    return t

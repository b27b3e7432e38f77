import ctypes

import torch

from .. import _lib
from .. import torch_ops as _T


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def require(t, name, dtype=torch.float32):
    """Same contract as the reference wrappers (contiguous CUDA tensors; e.g. knn.cpp:11-13,
    furthest_point_sample.py:24), raised as exceptions instead of C++ aborts."""
    if not torch.is_tensor(t):
        raise TypeError(f"{name} must be a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (pcreid_b200 has no CPU fallback)")
    assert t.is_contiguous(), f"{name} must be contiguous"
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")


class _NoBackward(torch.autograd.Function):
    """Forward-only scope: training is out of scope for this drop-in (SURVEY.md 8b 'Autograd')."""

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError("pcreid_b200 implements the inference path only")


lib = _lib.lib
check = _lib.check
OPS = _T.ops            # torch.ops.pcreid.* (torch_ops.py): the op wrappers call the C ABI through the dispatcher

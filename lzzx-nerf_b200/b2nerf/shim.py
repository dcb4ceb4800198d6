"""Tensor -> C-ABI marshalling shared by the four `_backend` shims.

Each shim function has the positional signature of the reference's pybind function of the same name
(raymarching/src/bindings.cpp:5-38, gridencoder/src/bindings.cpp:5-8, shencoder/src/bindings.cpp:5-8,
freqencoder/src/bindings.cpp:5-8): tensors in, nothing returned, outputs written in place.
"""
import torch

from ._lib import lib, B2NError

F32, F16 = 0, 1


def stream_ptr(t=None):
    return torch.cuda.current_stream(t.device if t is not None else None).cuda_stream


def dev_ptr(t, name, dtype=None, optional=False):
    """Validated device pointer of a tensor (the reference's CHECK_CUDA / CHECK_CONTIGUOUS / dtype checks)."""
    if t is None:
        if optional:
            return None
        raise B2NError(f"{name} must be a CUDA tensor, got None")
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise B2NError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise B2NError(f"{name} must be a contiguous tensor")
    if dtype is not None:
        ok = t.dtype in dtype if isinstance(dtype, (tuple, list)) else t.dtype == dtype
        if not ok:
            raise B2NError(f"{name} must be a {dtype} tensor, got {t.dtype}")
    return t.data_ptr() if t.numel() else _nonnull(t)


def _nonnull(t):
    # empty tensors have a null data_ptr; the library only dereferences when the count is non-zero
    return 1


def float_code(t, name):
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.float16:
        return F16
    raise B2NError(f"{name} must be a float32 or float16 tensor, got {t.dtype}")


def call(name, *args):
    lib().call(name, *args)

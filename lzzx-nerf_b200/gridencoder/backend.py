"""`_backend` of the drop-in `gridencoder` package: grid_encode_forward / grid_encode_backward with the positional
signatures of gridencoder/src/gridencoder.h:12-13, on libb2nerf.so.  Validation mirrors gridencoder.cu:425-441:
CUDA + contiguous for every tensor, int32 offsets, float32/float16 tables (float64 is not supported)."""
import types

import torch

from b2nerf.shim import call, dev_ptr, float_code, stream_ptr, B2NError

_flt = (torch.float32, torch.float16)


def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners):
    code = float_code(embeddings, "embeddings")
    if outputs.dtype != embeddings.dtype or (dy_dx is not None and dy_dx.dtype != embeddings.dtype):
        raise B2NError("outputs / dy_dx must have the dtype of embeddings")
    call("b2n_grid_encode_forward", dev_ptr(inputs, "inputs", torch.float32), dev_ptr(embeddings, "embeddings", _flt),
         dev_ptr(offsets, "offsets", torch.int32), dev_ptr(outputs, "outputs", _flt), B, D, C, L, float(S), H,
         dev_ptr(dy_dx, "dy_dx", _flt, optional=True), gridtype, int(bool(align_corners)), code, stream_ptr(inputs))


def grid_encode_forward_rows(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, gridtype, align_corners):
    """Not part of the reference's pybind surface: the forward written straight in GridEncoder.forward's [B, L*C] layout (no permute copy)."""
    code = float_code(embeddings, "embeddings")
    if outputs.dtype != embeddings.dtype:
        raise B2NError("outputs must have the dtype of embeddings")
    call("b2n_grid_encode_forward_rows", dev_ptr(inputs, "inputs", torch.float32), dev_ptr(embeddings, "embeddings", _flt),
         dev_ptr(offsets, "offsets", torch.int32), dev_ptr(outputs, "outputs", _flt), B, D, C, L, float(S), H, gridtype, int(bool(align_corners)), code,
         stream_ptr(inputs))


def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype, align_corners):
    code = float_code(grad, "grad")
    if grad_embeddings.dtype != grad.dtype:
        raise B2NError("grad_embeddings must have the dtype of grad")
    call("b2n_grid_encode_backward", dev_ptr(grad, "grad", _flt), dev_ptr(inputs, "inputs", torch.float32),
         dev_ptr(embeddings, "embeddings", _flt), dev_ptr(offsets, "offsets", torch.int32),
         dev_ptr(grad_embeddings, "grad_embeddings", _flt), B, D, C, L, float(S), H,
         dev_ptr(dy_dx, "dy_dx", _flt, optional=True), dev_ptr(grad_inputs, "grad_inputs", _flt, optional=True),
         gridtype, int(bool(align_corners)), code, stream_ptr(inputs))


def grid_level_scales(S, H, L, device="cuda"):
    """Per-level scales exp2f(l*S)*H - 1 with the kernels' own arithmetic (diagnostic; not part of the reference surface)."""
    out = torch.empty(L, dtype=torch.float32, device=device)
    call("b2n_grid_level_scales", float(S), H, L, out.data_ptr(), stream_ptr(out))
    return out


_backend = types.SimpleNamespace(grid_encode_forward=grid_encode_forward, grid_encode_backward=grid_encode_backward)
_rows_ok = lambda D, C, L: D <= 3 and C in (1, 2, 4, 8) and L * C <= 80
__all__ = ["_backend"]

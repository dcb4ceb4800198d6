"""`_backend` of the drop-in `shencoder` package (shencoder/src/shencoder.h:9-10) on libb2nerf.so."""
import types

import torch

from b2nerf.shim import call, dev_ptr, stream_ptr

f32 = torch.float32


def sh_encode_forward(inputs, outputs, B, D, C, dy_dx):
    call("b2n_sh_encode_forward", dev_ptr(inputs, "inputs", f32), dev_ptr(outputs, "outputs", f32), B, D, C,
         dev_ptr(dy_dx, "dy_dx", f32, optional=True), stream_ptr(inputs))


def sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs):
    call("b2n_sh_encode_backward", dev_ptr(grad, "grad", f32), dev_ptr(inputs, "inputs", f32), B, D, C,
         dev_ptr(dy_dx, "dy_dx", f32), dev_ptr(grad_inputs, "grad_inputs", f32), stream_ptr(inputs))


_backend = types.SimpleNamespace(sh_encode_forward=sh_encode_forward, sh_encode_backward=sh_encode_backward)
__all__ = ["_backend"]

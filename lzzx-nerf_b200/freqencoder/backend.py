"""`_backend` of the drop-in `freqencoder` package (freqencoder/src/freqencoder.h:7,10) on libb2nerf.so."""
import types

import torch

from b2nerf.shim import call, dev_ptr, stream_ptr

f32 = torch.float32


def freq_encode_forward(inputs, B, D, deg, C, outputs):
    call("b2n_freq_encode_forward", dev_ptr(inputs, "inputs", f32), B, D, deg, C, dev_ptr(outputs, "outputs", f32), stream_ptr(inputs))


def freq_encode_backward(grad, outputs, B, D, deg, C, grad_inputs):
    call("b2n_freq_encode_backward", dev_ptr(grad, "grad", f32), dev_ptr(outputs, "outputs", f32), B, D, deg, C,
         dev_ptr(grad_inputs, "grad_inputs", f32), stream_ptr(grad))


_backend = types.SimpleNamespace(freq_encode_forward=freq_encode_forward, freq_encode_backward=freq_encode_backward)
__all__ = ["_backend"]

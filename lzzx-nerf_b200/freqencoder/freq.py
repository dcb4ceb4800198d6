"""Drop-in `freqencoder`: sin/cos positional encoding [x, sin(2^0 x), cos(2^0 x), ...] (reference: freqencoder/freq.py)."""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from .backend import _backend


class _freq_encoder(Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, inputs, degree, output_dim):
        inputs = (inputs if inputs.is_cuda else inputs.cuda()).contiguous()
        B, D = inputs.shape
        outputs = inputs.new_empty(B, output_dim)
        _backend.freq_encode_forward(inputs, B, D, degree, output_dim, outputs)
        ctx.save_for_backward(inputs, outputs)
        ctx.geom = (B, D, degree, output_dim)
        return outputs

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, grad):
        inputs, outputs = ctx.saved_tensors
        B, D, degree, output_dim = ctx.geom
        grad_inputs = torch.zeros_like(inputs)
        _backend.freq_encode_backward(grad.contiguous(), outputs, B, D, degree, output_dim, grad_inputs)
        return grad_inputs, None, None


freq_encode = _freq_encoder.apply


class FreqEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree = input_dim, degree
        self.output_dim = input_dim * (1 + 2 * degree)

    def __repr__(self):
        return f"FreqEncoder: input_dim={self.input_dim} degree={self.degree} output_dim={self.output_dim}"

    def forward(self, inputs, **kwargs):
        lead = list(inputs.shape[:-1])
        flat = inputs.reshape(-1, self.input_dim)
        return freq_encode(flat, self.degree, self.output_dim).reshape(lead + [self.output_dim])

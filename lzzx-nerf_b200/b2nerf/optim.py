"""AdamW over one flat parameter buffer (csrc/optim.cu) — the optimizer line of the reference's training loop (train.py:274 AdamW(betas=(0, 0.99),
eps=1e-8) over network.py:315-357's two parameter groups; TrainerUtil.py:1040-1056 scaler.step / scaler.update).

Every parameter's storage is re-pointed into one contiguous fp32 buffer (tables first, then the networks), so the whole step — GradScaler
unscale, overflow skip, decoupled weight decay, moments, update — is ONE kernel over 683 509 floats instead of ~70 tensors through a multi-tensor
apply.  It plugs into torch.amp.GradScaler through the `_step_supports_amp_scaling` protocol: the scaler hands over its device-side scale and
found-inf flag, so the step never synchronises.
"""
import torch

from ._lib import lib


class FlatAdamW(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, group0, group1, lr0, lr1, weight_decay0=0.01, weight_decay1=0.0, betas=(0.0, 0.99), eps=1e-8):
        group0, group1 = list(group0), list(group1)
        super().__init__([{"params": group0, "lr": lr0, "weight_decay": weight_decay0}, {"params": group1, "lr": lr1, "weight_decay": weight_decay1}],
                         dict(lr=lr0, betas=betas, eps=eps, weight_decay=weight_decay0))
        params = group0 + group1
        dev = params[0].device
        if dev.type != "cuda" or any(p.dtype != torch.float32 for p in params):
            raise RuntimeError("FlatAdamW: float32 CUDA parameters only")
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        for p in params:                                  # re-point the parameter storage into the flat buffer (values preserved)
            view = self.flat[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            off += p.numel()
        self.n, self.n0 = n, sum(p.numel() for p in group0)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)
        self.flat_grad = None                              # set by attach_grads()

    def attach_grads(self, flat_grad):
        """`flat_grad`: the FlatGradBuffer's buffer, laid out in the same parameter order."""
        if flat_grad.numel() != self.n:
            raise RuntimeError("FlatAdamW: gradient buffer does not match the parameter buffer")
        self.flat_grad = flat_grad

    @torch.no_grad()
    def step(self, closure=None):
        if self.flat_grad is None:
            raise RuntimeError("FlatAdamW: call attach_grads(FlatGradBuffer.flat) first")
        g0, g1 = self.param_groups
        b1, b2 = g0["betas"]
        scale, found = getattr(self, "grad_scale", None), getattr(self, "found_inf", None)
        lib().call("b2n_adamw_flat", self.flat.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.n, self.n0,
                   float(g0["lr"]), float(g0["weight_decay"]), float(g1["lr"]), float(g1["weight_decay"]), float(b1), float(b2), float(g0["eps"]),
                   self.step_count.data_ptr(), None if scale is None else scale.data_ptr(), None if found is None else found.data_ptr(),
                   torch.cuda.current_stream().cuda_stream)
        return None

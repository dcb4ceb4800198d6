"""AdamW over one flat parameter buffer (csrc/optim.cu) — the optimizer lines of the reference's training loop: train.py:274
AdamW(model.get_params(lr, lr_net), betas=(0, 0.99), eps=1e-8) with network.py:332-356's parameter groups, TrainerUtil.py:1040-1056
scaler.step / scaler.update / lr_scheduler.step / ema.update.

Every parameter's storage is re-pointed into one contiguous fp32 buffer (group after group), so the whole step — GradScaler unscale, overflow skip,
decoupled weight decay, moments, update and (when due) the weight EMA — is ONE kernel over 683 509 floats instead of ~70 tensors through a multi-tensor
apply.  It plugs into torch.amp.GradScaler through the `_step_supports_amp_scaling` protocol: the scaler hands over its device-side scale and
found-inf flag, so the step never synchronises.  torch LR schedulers work unchanged (they rewrite param_groups[k]["lr"], read at every step).
"""
import ctypes

import torch

from ._lib import lib


class _AdamGroupsC(ctypes.Structure):
    """b2n_adam_groups (include/b2nerf_fused.h)"""
    _fields_ = [("n_groups", ctypes.c_uint32), ("end", ctypes.c_uint32 * 8), ("lr", ctypes.c_float * 8), ("weight_decay", ctypes.c_float * 8)]


def reference_param_groups(model, lr=1e-2, lr_net=1e-3, wd=0.0):
    """NeRFNetwork.get_params for the head stage (network.py:332-356) on a HeadModel: [tables | networks + individual codes | audio_att_net]."""
    enc = [model.encoder_xy.embeddings, model.encoder_yz.embeddings, model.encoder_xz.embeddings]
    att = list(model.audio_att_net.parameters()) if getattr(model, "att", 0) > 0 else []
    skip = {id(p) for p in enc + att}
    net = [p for p in model.parameters() if id(p) not in skip]
    groups = [{"params": enc, "lr": lr, "weight_decay": 0.01},               # no weight_decay key in the reference => AdamW's default 0.01
              {"params": net, "lr": lr_net, "weight_decay": wd}]
    if att:
        groups.append({"params": att, "lr": lr_net * 5, "weight_decay": 0.0001})
    return groups


class FlatAdamW(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, param_groups, betas=(0.0, 0.99), eps=1e-8, ema_decay=None, ema_update_interval=1000):
        """param_groups: list of {"params", "lr", "weight_decay"} (<= 8).  ema_decay: keep torch_ema-style shadow weights (`self.ema`, flat), updated every
        `ema_update_interval` optimizer steps inside the AdamW kernel (TrainerUtil.py:98-99: decay 0.95, interval 1000)."""
        param_groups = [dict(g, params=list(g["params"])) for g in param_groups]
        if not 1 <= len(param_groups) <= 8:
            raise RuntimeError("FlatAdamW: 1..8 parameter groups")
        super().__init__(param_groups, dict(lr=1e-3, betas=betas, eps=eps, weight_decay=0.01))
        params = [p for g in self.param_groups for p in g["params"]]
        dev = params[0].device
        if dev.type != "cuda" or any(p.dtype != torch.float32 for p in params):
            raise RuntimeError("FlatAdamW: float32 CUDA parameters only")
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        off, self.ends = 0, []
        for g in self.param_groups:
            for p in g["params"]:                          # re-point the parameter storage into the flat buffer (values preserved)
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                off += p.numel()
            self.ends.append(off)
        self.n = n
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)
        self.flat_grad = None                              # set by attach_grads()
        self.ema_decay, self.ema_update_interval = ema_decay, int(ema_update_interval)
        self.ema = self.flat.clone() if ema_decay is not None else None
        self.ema_updates, self.host_steps = 0, 0

    def ordered_params(self):
        return [p for g in self.param_groups for p in g["params"]]

    def attach_grads(self, flat_grad):
        """`flat_grad`: the FlatGradBuffer's buffer, laid out in the same parameter order (ordered_params())."""
        if flat_grad.numel() != self.n:
            raise RuntimeError("FlatAdamW: gradient buffer does not match the parameter buffer")
        self.flat_grad = flat_grad

    @torch.no_grad()
    def step(self, closure=None, scaler=None):
        """scaler: a FlatGradScaler — the whole GradScaler protocol then runs on the device inside this call (b2n_adamw_flat_scaled).  Without it the optimizer
        follows torch.amp.GradScaler's `_step_supports_amp_scaling` hand-off (grad_scale / found_inf attributes), or steps unscaled."""
        if self.flat_grad is None:
            raise RuntimeError("FlatAdamW: call attach_grads(FlatGradBuffer.flat) first")
        gs = _AdamGroupsC()
        gs.n_groups = len(self.param_groups)
        for k, g in enumerate(self.param_groups):
            gs.end[k], gs.lr[k], gs.weight_decay[k] = self.ends[k], float(g["lr"]), float(g["weight_decay"])
        b1, b2 = self.param_groups[0]["betas"]
        scale, found = getattr(self, "grad_scale", None), getattr(self, "found_inf", None)
        self.host_steps += 1
        ema, decay = None, 0.0
        if self.ema is not None and self.host_steps % self.ema_update_interval == 0:
            # torch_ema: decay = min(decay, (1 + num_updates) / (10 + num_updates)), num_updates counted from 1
            self.ema_updates += 1
            ema, decay = self.ema, min(self.ema_decay, (1 + self.ema_updates) / (10 + self.ema_updates))
        if scaler is not None:
            lib().call("b2n_adamw_flat_scaled", self.flat.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.n,
                       ctypes.byref(gs), float(b1), float(b2), float(self.param_groups[0]["eps"]), self.step_count.data_ptr(), scaler.state.data_ptr(),
                       float(scaler.growth_factor), float(scaler.backoff_factor), int(scaler.growth_interval), None if ema is None else ema.data_ptr(), float(decay),
                       torch.cuda.current_stream().cuda_stream)
            return None
        lib().call("b2n_adamw_flat_groups", self.flat.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.n,
                   ctypes.byref(gs), float(b1), float(b2), float(self.param_groups[0]["eps"]), self.step_count.data_ptr(),
                   None if scale is None else scale.data_ptr(), None if found is None else found.data_ptr(), None if ema is None else ema.data_ptr(),
                   float(decay), torch.cuda.current_stream().cuda_stream)
        return None

    # ---- EMA weights for evaluation (TrainerUtil.py:928-939: ema.store(); ema.copy_to(); ...; ema.restore()) -----------------------------
    @torch.no_grad()
    def ema_swap_in(self):
        if self.ema is None:
            return
        self._stored = self.flat.clone()
        self.flat.copy_(self.ema)

    @torch.no_grad()
    def ema_restore(self):
        if self.ema is not None and getattr(self, "_stored", None) is not None:
            self.flat.copy_(self._stored)
            self._stored = None

    # ---- checkpointing (TrainerUtil.py:1222-1345 saves optimizer.state_dict()): the moments live outside Optimizer.state --------------------
    def state_dict(self):
        sd = super().state_dict()
        sd["flat"] = {"exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(), "step": self.step_count.clone(), "host_steps": self.host_steps,
                      "ema": None if self.ema is None else self.ema.clone(), "ema_updates": self.ema_updates}
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)
        flat = sd.pop("flat", None)
        super().load_state_dict(sd)
        if flat is not None:
            self.exp_avg.copy_(flat["exp_avg"]); self.exp_avg_sq.copy_(flat["exp_avg_sq"]); self.step_count.copy_(flat["step"])
            self.host_steps, self.ema_updates = int(flat["host_steps"]), int(flat["ema_updates"])
            if self.ema is not None and flat["ema"] is not None:
                self.ema.copy_(flat["ema"])


class FlatGradScaler:
    """torch.amp.GradScaler for a FlatAdamW (same scale / growth / backoff rules and defaults; TrainerUtil.py:1045-1047 scaler.scale(loss).backward();
    scaler.step(optimizer); scaler.update()), entirely on the device: `state` = [scale, growth tracker, found_inf, -].  step() hands the state to the optimizer's
    kernel chain (non-finite check over the flat gradient buffer, unscale + skip inside AdamW, scale update), so update() has nothing left to do and nothing ever
    synchronises."""

    def __init__(self, device, init_scale=65536.0, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000, enabled=True):
        self.enabled = bool(enabled)
        self.growth_factor, self.backoff_factor, self.growth_interval = growth_factor, backoff_factor, growth_interval
        self.state = torch.tensor([init_scale if enabled else 1.0, 0.0, 0.0, 0.0], dtype=torch.float32, device=device)

    def scale(self, loss):
        return loss * self.state[0] if self.enabled else loss

    def step(self, optimizer):
        return optimizer.step(scaler=self) if self.enabled else optimizer.step()

    def update(self):
        return None

    def get_scale(self):
        return float(self.state[0].item())

    def state_dict(self):
        return {"state": self.state.clone(), "growth_factor": self.growth_factor, "backoff_factor": self.backoff_factor, "growth_interval": self.growth_interval}

    def load_state_dict(self, sd):
        self.state.copy_(sd["state"])
        self.growth_factor, self.backoff_factor, self.growth_interval = sd["growth_factor"], sd["backoff_factor"], sd["growth_interval"]

"""Training step of the head model on the B200-native ops (the role of TrainerUtil.train_step + the optimizer lines of
train_one_epoch, TrainerUtil.py:188-367, 1040-1056; renderer.py:279-304).

    near/far -> march_rays_train (mean_count estimate, 128-aligned) -> network (autocast fp16) -> composite_rays_train_triplane
    -> MSE on the composited colour (+ the reference's entropy / ambient regularisers) -> backward through the drop-in autograd
    Functions (composite backward, grid_encode backward with vector reductions) -> one flat-buffer all-reduce (DP) -> AdamW.

The per-op graph is the reference's; what is B200-specific is underneath (csrc/) and in the data-parallel plumbing (dist.py).
"""
import torch

import raymarching

from .dist import FlatGradBuffer, broadcast_occupancy
from .model import HeadModel


class Trainer:
    def __init__(self, model: HeadModel, lr=1e-2, lr_net=1e-3, fp16=True, max_steps=16, dt_gamma=1.0 / 256, min_near=0.05, lambda_amb=1e-4):
        self.m = model
        self.fp16, self.max_steps, self.dt_gamma, self.min_near, self.lambda_amb = fp16, max_steps, dt_gamma, min_near, lambda_amb
        enc = [model.encoder_xy.embeddings, model.encoder_yz.embeddings, model.encoder_xz.embeddings]
        enc_ids = {id(p) for p in enc}
        net = [p for p in model.parameters() if id(p) not in enc_ids]
        # AdamW(betas=(0.0, 0.99), eps=1e-8) with lr for the tables and lr_net for the networks (train.py:274, network.py:315-357)
        self.opt = torch.optim.AdamW([{"params": enc, "lr": lr}, {"params": net, "lr": lr_net, "weight_decay": 0}], betas=(0.0, 0.99), eps=1e-8)
        self.grads = FlatGradBuffer(list(model.parameters()))
        self.scaler = torch.amp.GradScaler("cuda", enabled=fp16)
        self.local_step = 0
        self.mean_count = 0

    def render_train(self, rays_o, rays_d, auds, index, eye, bg_color, perturb=True):
        """run_cuda's training branch (renderer.py:279-304).  Returns dict(image, weights_sum, ambient_aud, ambient_eye, uncertainty, n_samples_buffer)."""
        m = self.m
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, m.aabb_train, self.min_near)
        enc_a = m.encode_audio(auds)
        ind_code = m.individual_codes[index]
        counter = m.step_counter[self.local_step % 16]
        counter.zero_()
        self.local_step += 1
        xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, m.bound, m.density_bitfield, m.cascade, m.grid_size, nears, fars, counter,
                                                                self.mean_count, perturb, 128, False, self.dt_gamma, self.max_steps)
        sigmas, rgbs, amb_aud, amb_eye, unc = m.forward_unfused(xyzs, dirs, enc_a, ind_code, eye)
        ws, aud_sum, eye_sum, unc_sum, depth, image = raymarching.composite_rays_train_triplane(
            sigmas, rgbs, amb_aud.abs().sum(-1), amb_eye.abs().sum(-1), unc, deltas, rays)
        image = (image + (1 - ws).unsqueeze(-1) * bg_color).clamp(0, 1)
        return dict(image=image, weights_sum=ws, ambient_aud=aud_sum, ambient_eye=eye_sum, uncertainty=unc_sum, n_samples_buffer=xyzs.shape[0])

    def loss(self, out, gt_rgb):
        # TrainerUtil.py:238-300 (head branch): MSE, entropy of the alpha channel, ambient regularisers
        mse = ((out["image"] - gt_rgb) ** 2).mean(-1).mean()
        alphas = out["weights_sum"].clamp(1e-5, 1 - 1e-5)
        entropy = (-alphas * torch.log2(alphas) - (1 - alphas) * torch.log2(1 - alphas)).mean()
        amb = self.lambda_amb * (out["ambient_aud"].mean() + out["ambient_eye"].mean())
        return mse + 1e-3 * entropy + amb

    def train_step(self, rays_o, rays_d, auds, gt_rgb, index=0, eye=None, bg_color=None, perturb=True):
        m = self.m
        if eye is None:
            eye = torch.full((1, 1), 0.4, device=rays_o.device)
        if bg_color is None:
            bg_color = torch.ones(1, 3, device=rays_o.device)
        self.grads.zero_()
        with torch.autocast("cuda", dtype=torch.float16, enabled=self.fp16):
            out = self.render_train(rays_o, rays_d, auds, index, eye, bg_color, perturb)
            loss = self.loss(out, gt_rgb)
        self.scaler.scale(loss).backward()
        self.grads.all_reduce_mean()
        self.scaler.step(self.opt)
        self.scaler.update()
        return loss.detach(), out["n_samples_buffer"]

    def update_mean_count(self):
        """update_extra_state's step-counter part (renderer.py:812-815): one D2H read every 16 steps."""
        total = min(16, self.local_step)
        if total > 0:
            self.mean_count = int(self.m.step_counter[:total, 0].sum().item() / total)
        self.local_step = 0

    def sync_occupancy(self):
        broadcast_occupancy(self.m)

"""Training step of the head model on the B200-native ops (the role of TrainerUtil.train_step + the optimizer lines of
train_one_epoch, TrainerUtil.py:188-367, 1040-1056; renderer.py:279-304).

    near/far -> march_rays_train (mean_count estimate, 128-aligned) -> network (autocast fp16) -> composite_rays_train_triplane
    -> MSE on the composited colour (+ the reference's entropy / ambient regularisers) -> backward through the drop-in autograd
    Functions (composite backward, grid_encode backward) -> one flat-buffer all-reduce (DP) -> fused AdamW.

`train_step` is the eager path (what the reference's loop does, op by op).  `train_step_graphed` replays forward + backward + all-reduce from
ONE CUDA graph (the step is launch-bound in eager mode: ~350 small kernels, 8.2 ms on the host vs 5.6 ms of device work) and then runs the
fused optimizer; no host synchronisation anywhere in the step (the GradScaler hands its found-inf flag to the fused AdamW on the device).

The per-op graph is the reference's; what is B200-specific is underneath (csrc/) and in the data-parallel plumbing (dist.py).
"""
import torch

import raymarching

from .dist import FlatGradBuffer, broadcast_occupancy
from .model import HeadModel


class Trainer:
    def __init__(self, model: HeadModel, lr=1e-2, lr_net=1e-3, fp16=True, max_steps=16, dt_gamma=1.0 / 256, min_near=0.05, lambda_amb=1e-4, fused_optimizer=True, fused_head=False):
        self.m = model
        self.fused_head = fused_head
        self.fp16, self.max_steps, self.dt_gamma, self.min_near, self.lambda_amb = fp16, max_steps, dt_gamma, min_near, lambda_amb
        enc = [model.encoder_xy.embeddings, model.encoder_yz.embeddings, model.encoder_xz.embeddings]
        enc_ids = {id(p) for p in enc}
        net = [p for p in model.parameters() if id(p) not in enc_ids]
        # AdamW(betas=(0.0, 0.99), eps=1e-8) with lr for the tables and lr_net for the networks (train.py:274, network.py:315-357)
        if enc[0].is_cuda and fused_optimizer:
            from .optim import FlatAdamW
            self.opt = FlatAdamW(enc, net, lr, lr_net, weight_decay0=0.01, weight_decay1=0.0, betas=(0.0, 0.99), eps=1e-8)    # one kernel, no sync
        else:
            self.opt = torch.optim.AdamW([{"params": enc, "lr": lr}, {"params": net, "lr": lr_net, "weight_decay": 0}], betas=(0.0, 0.99), eps=1e-8)
        self.grads = FlatGradBuffer(enc + net)           # same parameter order as the optimizer's flat buffers
        if hasattr(self.opt, "attach_grads"):
            self.opt.attach_grads(self.grads.flat)
            if getattr(model, "_handle", None) is not None:
                model.pack()                              # the parameter storage moved: refresh the pointers baked into the packed model
        # the fused backward accumulates every gradient straight into these .grad views (fused_train._direct_targets): no per-parameter add kernels
        model._direct_grads = bool(fused_head and enc[0].is_cuda)
        self.scaler = torch.amp.GradScaler("cuda", enabled=fp16)
        self.local_step = 0
        self.mean_count = 0
        self._graphs = {}                # M bucket -> (CUDAGraph, static buffers)
        self._g_counter_sum = None

    def render_train(self, rays_o, rays_d, auds, index, eye, bg_color, perturb=True, counter=None, mean_count=None):
        """run_cuda's training branch (renderer.py:279-304).  Returns dict(image, weights_sum, ambient_aud, ambient_eye, uncertainty, n_samples_buffer).
        `index` may be an int or a 1-element device tensor (graph mode); `counter` overrides the step-counter slot."""
        m = self.m
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, m.aabb_train, self.min_near)
        fused_audio = self.fused_head and m.att > 0
        enc_a = None if fused_audio else m.encode_audio(auds)       # fused: the audio nets are part of the head's autograd node
        ind_code = m.individual_codes.index_select(0, index)[0] if torch.is_tensor(index) else m.individual_codes[index]
        if counter is None:
            counter = m.step_counter[self.local_step % 16]
            self.local_step += 1
        counter.zero_()
        xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, m.bound, m.density_bitfield, m.cascade, m.grid_size, nears, fars, counter,
                                                                self.mean_count if mean_count is None else mean_count, perturb, 128, False, self.dt_gamma,
                                                                self.max_steps)
        if self.fused_head:
            from .fused_train import fused_head_train
            m.pack()                                   # the optimizer moved the weights: refresh the operand images (three small kernels)
            if fused_audio:
                from .fused_train import fused_head_audio_train
                sigmas, rgbs, amb_aud, amb_eye, unc = fused_head_audio_train(m, xyzs, dirs, auds, ind_code, eye)
            else:
                sigmas, rgbs, amb_aud, amb_eye, unc = fused_head_train(m, xyzs, dirs, enc_a, ind_code, eye)
        else:
            sigmas, rgbs, amb_aud, amb_eye, unc = m.forward_unfused(xyzs, dirs, enc_a, ind_code, eye)
        if self.fused_head:
            # ambient.abs().sum(-1) (renderer.py:294-295) over a trailing dimension of 1 of values that are >= 0 by construction (an L2 norm and a
            # sigmoid): the identity, forward and backward — skip the eight abs / sum / sign / mul kernels
            amb_a, amb_e = amb_aud.reshape(-1), amb_eye.reshape(-1)
        else:
            amb_a, amb_e = amb_aud.abs().sum(-1), amb_eye.abs().sum(-1)
        ws, aud_sum, eye_sum, unc_sum, depth, image = raymarching.composite_rays_train_triplane(sigmas, rgbs, amb_a, amb_e, unc, deltas, rays)
        raw = image
        image = None if self.fused_head else (image + (1 - ws).unsqueeze(-1) * bg_color).clamp(0, 1)      # the fused loss blends the background itself
        return dict(image=image, raw_image=raw, bg_color=bg_color, weights_sum=ws, ambient_aud=aud_sum, ambient_eye=eye_sum, uncertainty=unc_sum,
                    n_samples_buffer=xyzs.shape[0])

    def loss(self, out, gt_rgb):
        # TrainerUtil.py:238-300 (head branch): MSE, entropy of the alpha channel, ambient regularisers
        if out["image"] is None:
            from .fused_train import fused_head_loss
            return fused_head_loss(out["raw_image"], out["weights_sum"], out["ambient_aud"], out["ambient_eye"], gt_rgb, out["bg_color"], 1e-3, self.lambda_amb)
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

    # ---- graph mode -----------------------------------------------------------------------------------------------------------------
    M_BUCKET = 8192      # sample-buffer sizes are rounded up to this in graph mode so that a changing mean_count rarely forces a re-capture

    def _capture(self, n_rays, bucket, perturb):
        m, dev = self.m, next(self.m.parameters()).device
        st = dict(rays_o=torch.zeros(n_rays, 3, device=dev), rays_d=torch.zeros(n_rays, 3, device=dev),
                  auds=torch.zeros(8, m.audio_in_dim, 2 if m.audio_in_dim == 1024 else 16, device=dev), gt=torch.zeros(n_rays, 3, device=dev),
                  index=torch.zeros(1, dtype=torch.long, device=dev), eye=torch.full((1, 1), 0.4, device=dev), bg=torch.ones(1, 3, device=dev),
                  counter=torch.zeros(2, dtype=torch.int32, device=dev))
        if self._g_counter_sum is None:
            self._g_counter_sum = torch.zeros(2, dtype=torch.int64, device=dev)
        st["rays_d"][:, 2] = 1.0

        def body():
            self.grads.zero_()
            with torch.autocast("cuda", dtype=torch.float16, enabled=self.fp16):
                out = self.render_train(st["rays_o"], st["rays_d"], st["auds"], st["index"], st["eye"], st["bg"], perturb, counter=st["counter"], mean_count=bucket)
                loss = self.loss(out, st["gt"])
            self.scaler.scale(loss).backward()
            self.grads.all_reduce_mean()
            self._g_counter_sum += st["counter"]
            return loss.detach(), out["n_samples_buffer"]

        # warm-up on a side stream with real-looking inputs (lazy cuDNN/cuBLAS init, geometry caches), then restore the optimiser-visible state
        return st, body

    def train_step_graphed(self, rays_o, rays_d, auds, gt_rgb, index=0, eye=None, bg_color=None, perturb=True):
        """Same arithmetic as train_step, forward + backward + all-reduce replayed from a CUDA graph.  Falls back to the eager step while
        mean_count is still unknown (the reference's first 16 steps size their buffers for the worst case and read the count back)."""
        if self.mean_count <= 0:
            return self.train_step(rays_o, rays_d, auds, gt_rgb, index, eye, bg_color, perturb)
        n = rays_o.shape[0]
        bucket = -(-self.mean_count // self.M_BUCKET) * self.M_BUCKET
        key = (n, bucket, bool(perturb))
        dev = rays_o.device
        if key not in self._graphs:
            st, body = self._capture(n, bucket, perturb)
            for k, v in (("rays_o", rays_o), ("rays_d", rays_d), ("auds", auds), ("gt", gt_rgb)):
                st[k].copy_(v)
            saved_sum = self._g_counter_sum.clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    body()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss, m_buf = body()
            self._g_counter_sum.copy_(saved_sum)
            if len(self._graphs) >= 4:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = (g, st, loss, m_buf)
        g, st, loss, m_buf = self._graphs[key]
        st["rays_o"].copy_(rays_o, non_blocking=True); st["rays_d"].copy_(rays_d, non_blocking=True)
        st["auds"].copy_(auds, non_blocking=True); st["gt"].copy_(gt_rgb, non_blocking=True)
        st["index"].fill_(int(index))
        if eye is not None:
            st["eye"].copy_(eye)
        if bg_color is not None:
            st["bg"].copy_(bg_color)
        g.replay()
        self.local_step += 1
        self.scaler.step(self.opt)
        self.scaler.update()
        return loss, m_buf

    def update_mean_count(self):
        """update_extra_state's step-counter part (renderer.py:812-815): one D2H read every 16 steps."""
        total = min(16, self.local_step)
        if total > 0:
            if self._g_counter_sum is not None and int(self._g_counter_sum[0].item()) > 0:      # steps taken through the graph
                self.mean_count = int(self._g_counter_sum[0].item() / total)
                self._g_counter_sum.zero_()
            else:
                self.mean_count = int(self.m.step_counter[:total, 0].sum().item() / total)
        self.local_step = 0

    def sync_occupancy(self):
        broadcast_occupancy(self.m)

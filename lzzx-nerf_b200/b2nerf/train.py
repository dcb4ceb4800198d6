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
    def __init__(self, model: HeadModel, lr=1e-2, lr_net=1e-3, fp16=True, max_steps=16, dt_gamma=1.0 / 256, min_near=0.05, lambda_amb=1e-4, fused_optimizer=True,
                 fused_head=False, iters=200000, unc_loss=True, amb_aud_loss=True, amb_eye_loss=True, ema_decay=None, ema_update_interval=1000, lr_schedule=True,
                 peer_allreduce=True):
        """iters: opt.iters (train.py:24) — the denominator of step_factor = min(global_step / iters, 1) that ramps the uncertainty / ambient terms
        (TrainerUtil.py:236) and of the LambdaLR decay 0.5 ** (iter / iters) (train.py:287-288).  ema_decay: 0.95 in the reference (train.py:296).
        peer_allreduce: under torch.distributed with one process per GPU, exchange the gradients with the NVLink peer-memory kernel (csrc/peer_allreduce.cu)
        instead of NCCL; a COLLECTIVE set-up — every rank must construct its Trainer at the same point."""
        self.m = model
        self.fused_head = fused_head
        import os
        # fused path: the audio encoder (a latency-bound 8-CTA kernel) and the operand packing run on a side stream beside the marcher — a second branch of
        # the step's graph (0.858 -> 0.820 ms per step on B200).  B2N_AUDIO_AHEAD=0 / B2N_PACK_AHEAD=0 keep them on the main stream (A/B).
        self.audio_ahead = os.environ.get("B2N_AUDIO_AHEAD", "1") == "1"
        self.pack_ahead = os.environ.get("B2N_PACK_AHEAD", "1") == "1"
        self.fp16, self.max_steps, self.dt_gamma, self.min_near, self.lambda_amb = fp16, max_steps, dt_gamma, min_near, lambda_amb
        self.iters, self.global_step = int(iters), 0
        self.unc_loss, self.amb_aud_loss, self.amb_eye_loss = bool(unc_loss), bool(amb_aud_loss), bool(amb_eye_loss)
        model.unc_loss = self.unc_loss
        from .optim import FlatAdamW, reference_param_groups
        groups = reference_param_groups(model, lr, lr_net)            # network.py:332-356: tables | networks | audio_att_net (5 x lr_net, wd 1e-4)
        enc = groups[0]["params"]
        # AdamW(betas=(0.0, 0.99), eps=1e-8) (train.py:274)
        if enc[0].is_cuda and fused_optimizer:
            self.opt = FlatAdamW(groups, betas=(0.0, 0.99), eps=1e-8, ema_decay=ema_decay, ema_update_interval=ema_update_interval)    # one kernel, no sync
        else:
            self.opt = torch.optim.AdamW(groups, betas=(0.0, 0.99), eps=1e-8)
        self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lambda it: 0.5 ** (it / self.iters)) if lr_schedule else None
        ordered = [p for g in groups for p in g["params"]]
        self.grads = FlatGradBuffer(ordered, peer=peer_allreduce)           # same parameter order as the optimizer's flat buffers
        self.n_table_grads = sum(p.numel() for p in enc)
        if hasattr(self.opt, "attach_grads"):
            self.opt.attach_grads(self.grads.flat)
            if getattr(model, "_handle", None) is not None:
                model.pack()                              # the parameter storage moved: refresh the pointers baked into the packed model
        # the fused backward accumulates every gradient straight into these .grad views (fused_train._direct_targets): no per-parameter add kernels
        model._direct_grads = bool(fused_head and enc[0].is_cuda)
        if isinstance(self.opt, FlatAdamW):
            from .optim import FlatGradScaler
            self.scaler = FlatGradScaler(enc[0].device, enabled=fp16)      # GradScaler's rules, on the device, inside the optimizer's launch chain
        else:
            self.scaler = torch.amp.GradScaler("cuda", enabled=fp16)
        self.local_step = 0
        self.mean_count = 0
        self._graphs = {}                # M bucket -> (CUDAGraph, static buffers)
        self._g_counter_sum = None

    def step_factor(self):
        return min(self.global_step / self.iters, 1.0)

    def _after_step(self):
        if self.sched is not None:
            self.sched.step()                             # scheduler_update_every_step (TrainerUtil.py:1048-1049)
        self.global_step += 1

    def render_train(self, rays_o, rays_d, auds, index, eye, bg_color, perturb=True, counter=None, mean_count=None):
        """run_cuda's training branch (renderer.py:279-304).  Returns dict(image, weights_sum, ambient_aud, ambient_eye, uncertainty, n_samples_buffer).
        `index` may be an int or a 1-element device tensor (graph mode); `counter` overrides the step-counter slot."""
        m = self.m
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, m.aabb_train, self.min_near)
        fused_audio = self.fused_head and m.att > 0
        enc_a = None if fused_audio else m.encode_audio(auds)       # fused: the audio nets are part of the head's autograd node
        ind_code = m.individual_codes.index_select(0, index)[0] if torch.is_tensor(index) else m.individual_codes[index]
        pre = None
        if fused_audio and self.audio_ahead:
            # the audio code does not depend on the rays: its (latency-bound, 8-CTA) kernel runs on a side stream while the marcher runs
            from .fused_train import audio_encode_ahead
            pre = audio_encode_ahead(m, auds, torch.cuda.current_stream(rays_o.device))
            if self.pack_ahead:
                with torch.cuda.stream(pre[3]):
                    m.pack()                           # behind the audio kernel on the same side stream
        if counter is None:
            counter = m.step_counter[self.local_step % 16]
            self.local_step += 1
        counter.zero_()
        xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, m.bound, m.density_bitfield, m.cascade, m.grid_size, nears, fars, counter,
                                                                self.mean_count if mean_count is None else mean_count, perturb, 128, False, self.dt_gamma,
                                                                self.max_steps)
        if self.fused_head:
            from .fused_train import fused_head_train
            if not (pre is not None and self.pack_ahead):
                m.pack()                               # the optimizer moved the weights: refresh the operand images (three small kernels)
            if fused_audio:
                from .fused_train import fused_head_audio_train
                sigmas, rgbs, amb_aud, amb_eye, unc = fused_head_audio_train(m, xyzs, dirs, auds, ind_code, eye, ahead=pre)
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
        return dict(image=image, raw_image=raw, bg_color=bg_color, weights_sum=ws, ambient_aud=aud_sum, ambient_eye=eye_sum, uncertainty=unc_sum, depth=depth,
                    nears=nears, fars=fars, n_samples_buffer=xyzs.shape[0], rays=(xyzs, dirs, enc_a, ind_code, eye), auds=auds)

    def loss(self, out, gt_rgb, face_mask=None, step_factor=None):
        """TrainerUtil.py:238-334, head branch (no LPIPS patch / lips-finetune terms): uncertainty-weighted MSE + loss_u + static-uncertainty term, entropy of
        the alpha channel (1e-4), ambient regularisers masked by ~face_mask / face_mask and ramped by step_factor.  face_mask [N] bool (None: every ray on the face)."""
        from .fused_train import fused_head_loss, torch_head_loss
        sf = self.step_factor() if step_factor is None else step_factor
        unc = out["uncertainty"] if (self.unc_loss and not self.m.testing) else None
        kw = dict(unc_sum=unc, face_mask=face_mask, step_factor=sf, lambda_ent=1e-4, lambda_amb=self.lambda_amb, max_steps=self.max_steps,
                  amb_aud_loss=self.amb_aud_loss, amb_eye_loss=self.amb_eye_loss)
        if out["image"] is None:
            return fused_head_loss(out["raw_image"], out["weights_sum"], out["ambient_aud"], out["ambient_eye"], gt_rgb, out["bg_color"], **kw)
        return torch_head_loss(out["image"], out["weights_sum"], out["ambient_aud"], out["ambient_eye"], gt_rgb, **kw)

    def reg_loss(self, out, step_factor):
        """The smoothness regulariser of every 16th step (TrainerUtil.py:336-363): the network evaluated again on the step's samples (no grad) and on samples
        moved by +-1e-3 (with grad), MSE between the uncertainty / ambient outputs, weighted step_factor * 1e-5; enc_a / ind_code detached like the reference.
        Fused path: one inference launch of the head kernel + one more fused training forward / backward; otherwise the op-by-op graph."""
        xyzs, dirs, enc_a, ind_code, eye = out["rays"]
        m = self.m
        delta = (torch.rand_like(xyzs) * 2 - 1) * 1e-3
        if self.fused_head:
            from .fused_train import fused_head_train
            if enc_a is None:
                enc_a = m.encode_audio_fused(out["auds"])
            enc_a, code = enc_a.detach().float(), ind_code.detach().float().view(1, -1)
            _, _, aud0, eye0, unc0 = m(xyzs, dirs, enc_a, code, eye)
            _, _, aud1, eye1, unc1 = fused_head_train(m, xyzs + delta, dirs, enc_a, code, eye)
        else:
            with torch.no_grad():
                _, _, aud0, eye0, unc0 = m.forward_unfused(xyzs, dirs, enc_a.detach(), ind_code.detach(), eye)
            _, _, aud1, eye1, unc1 = m.forward_unfused(xyzs + delta, dirs, enc_a.detach(), ind_code.detach(), eye)
        reg = 0
        if self.unc_loss:
            reg = reg + ((unc0.float() - unc1.float()) ** 2).mean()
        if self.amb_aud_loss:
            reg = reg + ((aud0.float() - aud1.float()) ** 2).mean()
        if self.amb_eye_loss:
            reg = reg + ((eye0.float() - eye1.float()) ** 2).mean()
        return reg * (step_factor * 1e-5)

    def _reg_due(self):
        return bool(self.reg_every) and self.global_step % self.reg_every == 0 and self.global_step > 0

    def train_step(self, rays_o, rays_d, auds, gt_rgb, index=0, eye=None, bg_color=None, perturb=True, face_mask=None):
        m = self.m
        if eye is None:
            eye = torch.full((1, 1), 0.4, device=rays_o.device)
        if bg_color is None:
            bg_color = torch.ones(1, 3, device=rays_o.device)
        self.grads.zero_()
        with torch.autocast("cuda", dtype=torch.float16, enabled=self.fp16):
            out = self.render_train(rays_o, rays_d, auds, index, eye, bg_color, perturb)
            loss = self.loss(out, gt_rgb, face_mask)
            if self._reg_due():
                loss = loss + self.reg_loss(out, self.step_factor())
        self.scaler.scale(loss).backward()
        self.grads.all_reduce_mean()
        self.scaler.step(self.opt)
        self.scaler.update()
        self._after_step()
        return loss.detach(), out["n_samples_buffer"]

    reg_every = 16       # TrainerUtil.py:337 (`global_step % 16 == 0`); 0 disables the regulariser (at global_step 0 its weight step_factor * 1e-5 is 0)

    # ---- graph mode -----------------------------------------------------------------------------------------------------------------
    M_BUCKET = 8192      # sample-buffer sizes are rounded up to this in graph mode so that a changing mean_count rarely forces a re-capture

    def _capture(self, n_rays, bucket, perturb, do_reg=False):
        m, dev = self.m, next(self.m.parameters()).device
        st = dict(rays_o=torch.zeros(n_rays, 3, device=dev), rays_d=torch.zeros(n_rays, 3, device=dev),
                  auds=torch.zeros(8, m.audio_in_dim, 2 if m.audio_in_dim == 1024 else 16, device=dev), gt=torch.zeros(n_rays, 3, device=dev),
                  index=torch.zeros(1, dtype=torch.long, device=dev), eye=torch.full((1, 1), 0.4, device=dev), bg=torch.ones(1, 3, device=dev),
                  counter=torch.zeros(2, dtype=torch.int32, device=dev), face=torch.ones(n_rays, dtype=torch.bool, device=dev),
                  step_factor=torch.zeros(1, device=dev))
        if self._g_counter_sum is None:
            self._g_counter_sum = torch.zeros(2, dtype=torch.int64, device=dev)
        st["rays_d"][:, 2] = 1.0

        def body():
            self.grads.zero_()
            with torch.autocast("cuda", dtype=torch.float16, enabled=self.fp16):
                out = self.render_train(st["rays_o"], st["rays_d"], st["auds"], st["index"], st["eye"], st["bg"], perturb, counter=st["counter"], mean_count=bucket)
                loss = self.loss(out, st["gt"], st["face"], st["step_factor"])
                if do_reg:
                    loss = loss + self.reg_loss(out, st["step_factor"])
            self.scaler.scale(loss).backward()
            self.grads.all_reduce_mean()
            self._g_counter_sum += st["counter"]
            return loss.detach(), out["n_samples_buffer"]

        # warm-up on a side stream with real-looking inputs (lazy cuDNN/cuBLAS init, geometry caches), then restore the optimiser-visible state
        return st, body

    def train_step_graphed(self, rays_o, rays_d, auds, gt_rgb, index=0, eye=None, bg_color=None, perturb=True, face_mask=None):
        """Same arithmetic as train_step, forward + backward + all-reduce replayed from a CUDA graph.  Falls back to the eager step while
        mean_count is still unknown (the reference's first 16 steps size their buffers for the worst case and read the count back)."""
        if self.mean_count <= 0:
            return self.train_step(rays_o, rays_d, auds, gt_rgb, index, eye, bg_color, perturb, face_mask)
        n = rays_o.shape[0]
        bucket = -(-self.mean_count // self.M_BUCKET) * self.M_BUCKET
        do_reg = self._reg_due()
        key = (n, bucket, bool(perturb), do_reg)
        dev = rays_o.device
        if key not in self._graphs:
            st, body = self._capture(n, bucket, perturb, do_reg)
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
        st["step_factor"].fill_(self.step_factor())
        if face_mask is not None:
            st["face"].copy_(face_mask.view(-1), non_blocking=True)
        else:
            st["face"].fill_(True)
        if eye is not None:
            st["eye"].copy_(eye)
        if bg_color is not None:
            st["bg"].copy_(bg_color)
        g.replay()
        self.local_step += 1
        self.scaler.step(self.opt)
        self.scaler.update()
        self._after_step()
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

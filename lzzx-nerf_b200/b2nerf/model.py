"""Head model of the talking-head NeRF on the B200-native kernels.

`HeadModel` holds exactly the parameters/buffers of the reference's `NeRFNetwork(NeRFRenderer)` head branch
(nerf_triplane/network.py:97-152, renderer.py:86-160) under the SAME names and shapes, so a reference checkpoint's
`model` state_dict loads with `strict=False` (torso parameters are ignored):
    audio_net.*, audio_att_net.*, encoder_{xy,yz,xz}.embeddings / .offsets, sigma_net.net.{0,1,2}.weight,
    color_net.net.{0,1}.weight, unc_net.net.{0,1}.weight, aud_ch_att_net.net.{0,1}.weight, eye_att_net.net.{0,1}.weight,
    individual_codes, aabb_train, aabb_infer, density_grid, density_bitfield, step_counter.

Two evaluation paths, same math:
  * forward_unfused(...)  — the reference's op-by-op graph (network.py:252-311) on the drop-in encoders + torch Linear;
                            differentiable; used for training through autograd and as the on-GPU parity partner;
  * forward(...)          — the fused tcgen05 kernel (csrc/fused_head.cu) through the C ABI (inference).
"""
import ctypes
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from gridencoder import GridEncoder
from shencoder import SHEncoder

from ._lib import lib


class _HeadSavedC(ctypes.Structure):
    """b2n_head_saved (include/b2nerf_fused.h)"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("x36", "ha", "he", "hu", "att", "s_in", "h1", "h2", "c_in", "hc", "misc")]


class _TallLinear(torch.autograd.Function):
    """y = x W^T for tall activations (M = 10^5..10^6 samples, fan-in / fan-out <= 128) — the shape of every head-MLP layer in a training step.
    Forward and input gradient are plain GEMMs; the WEIGHT gradient dY^T X reduces over M, and the library heuristic runs it as one wave of
    4-6 CTAs (170-300 us per layer in the step profile, profiles/r1_train_step.md).  Under autocast (fp16 operands) it is computed by
    b2n_linear_wgrad (csrc/wgrad.cu: one pass over both activations, fp32 accumulation in tensor memory); in fp32 mode the reduction is
    split into S slabs with a batched GEMM and the partials are summed."""

    SLABS = 64

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float16)
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return x @ w.t()

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dy @ w if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            M, n_out, n_in = x.shape[0], w.shape[0], w.shape[1]
            if x.dtype == torch.float16 and dy.dtype == torch.float16 and n_out <= 128 and n_in <= 128:
                # hand-written kernel (csrc/wgrad.cu): both operands streamed once as MN-major tcgen05 operands, fp32 accumulation in TMEM
                acc = torch.zeros(n_out, n_in, dtype=torch.float32, device=x.device)
                lib().call("b2n_linear_wgrad", dy.data_ptr(), x.data_ptr(), M, n_out, n_in, acc.data_ptr(), torch.cuda.current_stream().cuda_stream)
                dw = acc.to(w.dtype)
            else:
                S = _TallLinear.SLABS
                part = torch.bmm(dy.view(S, M // S, -1).transpose(1, 2), x.view(S, M // S, -1))       # [S, out, in]
                dw = part.float().sum(0).to(w.dtype)
        return dx, dw


class MLP(nn.Module):
    """Bias-free ReLU MLP (network.py:73-94)."""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers):
        super().__init__()
        self.dim_in, self.dim_out, self.dim_hidden, self.num_layers = dim_in, dim_out, dim_hidden, num_layers
        dims = [dim_in] + [dim_hidden] * (num_layers - 1) + [dim_out]
        self.net = nn.ModuleList(nn.Linear(dims[i], dims[i + 1], bias=False) for i in range(num_layers))

    tall_linear = True       # class switch: False = plain nn.Linear everywhere (the reference's MLP, used by profiles/reference_on_b200.py)

    def forward(self, x):
        tall = (MLP.tall_linear and x.is_cuda and x.dim() == 2 and x.shape[0] >= 16384 and x.shape[0] % _TallLinear.SLABS == 0 and torch.is_grad_enabled()
                and any(l.weight.requires_grad for l in self.net))
        for i, layer in enumerate(self.net):
            x = _TallLinear.apply(x.contiguous(), layer.weight) if tall else layer(x)
            if i != self.num_layers - 1:
                x = F.relu(x, inplace=True)
        return x


class AudioAttNet(nn.Module):
    """Attention over the 8-frame audio window (network.py:9-36)."""

    def __init__(self, dim_aud=64, seq_len=8):
        super().__init__()
        self.seq_len, self.dim_aud = seq_len, dim_aud
        chans = [dim_aud, 16, 8, 4, 2, 1]
        layers = []
        for cin, cout in zip(chans[:-1], chans[1:]):
            layers += [nn.Conv1d(cin, cout, kernel_size=3, stride=1, padding=1, bias=True), nn.LeakyReLU(0.02, True)]
        self.attentionConvNet = nn.Sequential(*layers)
        self.attentionNet = nn.Sequential(nn.Linear(seq_len, seq_len, bias=True), nn.Softmax(dim=1))

    def forward(self, x):                      # x [1, seq_len, dim_aud]
        y = self.attentionConvNet(x.permute(0, 2, 1))
        y = self.attentionNet(y.view(1, self.seq_len)).view(1, self.seq_len, 1)
        return torch.sum(y * x, dim=1)


class AudioNet(nn.Module):
    """Per-frame audio feature extractor (network.py:40-70)."""

    def __init__(self, dim_in=29, dim_aud=64, win_size=16):
        super().__init__()
        self.win_size, self.dim_aud = win_size, dim_aud
        chans = [dim_in, 32, 32, 64, 64]
        layers = []
        for cin, cout in zip(chans[:-1], chans[1:]):
            layers += [nn.Conv1d(cin, cout, kernel_size=3, stride=2, padding=1, bias=True), nn.LeakyReLU(0.02, True)]
        self.encoder_conv = nn.Sequential(*layers)
        self.encoder_fc1 = nn.Sequential(nn.Linear(64, 64), nn.LeakyReLU(0.02, True), nn.Linear(64, dim_aud))

    def forward(self, x):
        half_w = self.win_size // 2
        x = x[:, :, 8 - half_w:8 + half_w]
        return self.encoder_fc1(self.encoder_conv(x).squeeze(-1))


class _HeadWeightsC(ctypes.Structure):      # mirrors b2n_head_weights (include/b2nerf_fused.h)
    _fields_ = [("table_xy", ctypes.c_void_p), ("table_yz", ctypes.c_void_p), ("table_xz", ctypes.c_void_p), ("offsets", ctypes.c_void_p),
                ("S", ctypes.c_float), ("H", ctypes.c_uint32), ("bound", ctypes.c_float),
                ("aud_att_w0", ctypes.c_void_p), ("aud_att_w1", ctypes.c_void_p), ("eye_att_w0", ctypes.c_void_p), ("eye_att_w1", ctypes.c_void_p),
                ("sigma_w0", ctypes.c_void_p), ("sigma_w1", ctypes.c_void_p), ("sigma_w2", ctypes.c_void_p),
                ("color_w0", ctypes.c_void_p), ("color_w1", ctypes.c_void_p), ("unc_w0", ctypes.c_void_p), ("unc_w1", ctypes.c_void_p)]


class _AudioWeightsC(ctypes.Structure):     # mirrors b2n_audio_weights
    _fields_ = [("conv_w", ctypes.c_void_p * 4), ("conv_b", ctypes.c_void_p * 4), ("fc_w", ctypes.c_void_p * 2), ("fc_b", ctypes.c_void_p * 2),
                ("att_conv_w", ctypes.c_void_p * 5), ("att_conv_b", ctypes.c_void_p * 5), ("att_fc_w", ctypes.c_void_p), ("att_fc_b", ctypes.c_void_p),
                ("dim_in", ctypes.c_uint32)]


class _RenderCfgC(ctypes.Structure):        # mirrors b2n_render_cfg
    _fields_ = [("bound", ctypes.c_float), ("dt_gamma", ctypes.c_float), ("min_near", ctypes.c_float), ("T_thresh", ctypes.c_float),
                ("density_scale", ctypes.c_float), ("max_steps", ctypes.c_uint32), ("cascade", ctypes.c_uint32), ("grid_size", ctypes.c_uint32),
                ("aabb", ctypes.c_float * 6), ("head_ctas", ctypes.c_uint32), ("image_width", ctypes.c_uint32)]


class HeadModel(nn.Module):
    def __init__(self, bound=1.0, audio_in_dim=1024, audio_dim=32, ind_dim=4, ind_num=10000, att=2, grid_size=128):
        super().__init__()
        self.bound, self.audio_dim, self.att, self.grid_size = float(bound), audio_dim, att, grid_size
        self.cascade = 1 + math.ceil(math.log2(bound))
        self.audio_in_dim = audio_in_dim
        self.audio_net = AudioNet(audio_in_dim, audio_dim)
        if att > 0:
            self.audio_att_net = AudioAttNet(audio_dim)
        grid = dict(input_dim=2, num_levels=12, level_dim=1, base_resolution=64, log2_hashmap_size=14, desired_resolution=512 * bound)
        self.encoder_xy, self.encoder_yz, self.encoder_xz = GridEncoder(**grid), GridEncoder(**grid), GridEncoder(**grid)
        self.in_dim = 36
        self.eye_att_net = MLP(self.in_dim, 1, 16, 2)
        self.sigma_net = MLP(self.in_dim + audio_dim + 1, 1 + 64, 64, 3)
        self.encoder_dir = SHEncoder(input_dim=3, degree=4)
        self.color_net = MLP(16 + 64 + ind_dim, 3, 64, 2)
        self.unc_net = MLP(self.in_dim, 1, 32, 2)
        self.aud_ch_att_net = MLP(self.in_dim, audio_dim, 64, 2)
        self.individual_codes = nn.Parameter(torch.randn(ind_num, ind_dim) * 0.1)
        aabb = torch.tensor([-bound, -bound / 2, -bound, bound, bound / 2, bound], dtype=torch.float32)
        self.register_buffer("aabb_train", aabb.clone())
        self.register_buffer("aabb_infer", aabb.clone())
        self.register_buffer("density_grid", torch.zeros(self.cascade, grid_size ** 3))
        self.register_buffer("density_bitfield", torch.zeros(self.cascade * grid_size ** 3 // 8, dtype=torch.uint8))
        self.register_buffer("step_counter", torch.zeros(16, 2, dtype=torch.int32))
        self.testing = False
        self.unc_loss = True
        self._handle = None

    # ---- audio prologue (network.py:226-240) --------------------------------------------------------------------------
    def encode_audio(self, a):
        if a is None:
            return None
        enc_a = self.audio_net(a)
        if self.att > 0:
            enc_a = self.audio_att_net(enc_a.unsqueeze(0))
        return enc_a

    def audio_weights_struct(self):
        p = lambda t: t.detach().data_ptr()
        conv = [self.audio_net.encoder_conv[i] for i in (0, 2, 4, 6)]
        fc = [self.audio_net.encoder_fc1[i] for i in (0, 2)]
        att = [self.audio_att_net.attentionConvNet[i] for i in (0, 2, 4, 6, 8)]
        lin = self.audio_att_net.attentionNet[0]
        return _AudioWeightsC((ctypes.c_void_p * 4)(*[p(c.weight) for c in conv]), (ctypes.c_void_p * 4)(*[p(c.bias) for c in conv]),
                              (ctypes.c_void_p * 2)(*[p(c.weight) for c in fc]), (ctypes.c_void_p * 2)(*[p(c.bias) for c in fc]),
                              (ctypes.c_void_p * 5)(*[p(c.weight) for c in att]), (ctypes.c_void_p * 5)(*[p(c.bias) for c in att]),
                              p(lin.weight), p(lin.bias), self.audio_in_dim)

    @torch.no_grad()
    def encode_audio_fused(self, a, out=None):
        """encode_audio as one cluster kernel (csrc/fused_audio.cu); a [8, dim_in, L] fp32 -> [1, 32] fp32 (inference, att > 0)."""
        if self.att <= 0:
            raise RuntimeError("encode_audio_fused needs the attention net (att > 0)")
        a = a.float().contiguous()
        if a.dim() != 3 or a.shape[0] != 8 or a.shape[1] != self.audio_in_dim:
            raise RuntimeError(f"encode_audio_fused: expected auds [8, {self.audio_in_dim}, L], got {tuple(a.shape)}")
        w = self.audio_weights_struct()
        enc = out if out is not None else torch.empty(1, 32, device=a.device)
        self._keep_audio = (a, w)
        lib().call("b2n_audio_encode", ctypes.byref(w), a.data_ptr(), a.shape[2], enc.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return enc

    # ---- reference graph on the drop-in ops (network.py:215-223, 252-311) ---------------------------------------------
    def encode_x(self, xyz):
        xy, yz, xz = xyz[:, :-1], xyz[:, 1:], torch.cat([xyz[:, :1], xyz[:, -1:]], dim=-1)
        return torch.cat([self.encoder_xy(xy, bound=self.bound), self.encoder_yz(yz, bound=self.bound), self.encoder_xz(xz, bound=self.bound)], dim=-1)

    def density(self, x, enc_a, e=None, enc_x=None):
        if enc_x is None:
            enc_x = self.encode_x(x)
        enc_a = enc_a.repeat(enc_x.shape[0], 1)
        aud_ch_att = self.aud_ch_att_net(enc_x)
        enc_w = enc_a * aud_ch_att
        eye_att = torch.sigmoid(self.eye_att_net(enc_x))
        h = torch.cat([enc_x, enc_w, e * eye_att], dim=-1)
        h = self.sigma_net(h)
        return {"sigma": torch.exp(h[..., 0]), "geo_feat": h[..., 1:], "ambient_aud": aud_ch_att.norm(dim=-1, keepdim=True), "ambient_eye": eye_att}

    def forward_unfused(self, x, d, enc_a, c, e):
        enc_x = self.encode_x(x)
        r = self.density(x, enc_a, e, enc_x)
        h = torch.cat([self.encoder_dir(d), r["geo_feat"], c.repeat(x.shape[0], 1)], dim=-1)
        color = torch.sigmoid(self.color_net(h)) * (1 + 2 * 0.001) - 0.001
        unc = torch.zeros_like(enc_x) if (self.testing or not self.unc_loss) else self.unc_net(enc_x.detach())
        unc = torch.log(1 + torch.exp(unc))
        return r["sigma"], color, r["ambient_aud"], r["ambient_eye"], unc[..., None]

    # ---- fused path -----------------------------------------------------------------------------------------------------
    def pack(self, with_unc=None):
        """(Re)pack the MLP weights into the tensor-core operand image; call after every optimizer step / load_state_dict."""
        L = lib()
        if self._handle is None:
            h = ctypes.c_void_p()
            L.call("b2n_model_create", ctypes.byref(h), None)
            self._handle = h
        with_unc = (not self.testing and self.unc_loss) if with_unc is None else with_unc
        p = lambda t: t.detach().data_ptr()
        for t in (self.encoder_xy.embeddings, self.sigma_net.net[0].weight):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("HeadModel.pack: parameters must be contiguous float32 CUDA tensors")
        w = _HeadWeightsC(p(self.encoder_xy.embeddings), p(self.encoder_yz.embeddings), p(self.encoder_xz.embeddings), p(self.encoder_xy.offsets),
                          float(math.log2(self.encoder_xy.per_level_scale)), self.encoder_xy.base_resolution, self.bound,
                          p(self.aud_ch_att_net.net[0].weight), p(self.aud_ch_att_net.net[1].weight), p(self.eye_att_net.net[0].weight),
                          p(self.eye_att_net.net[1].weight), p(self.sigma_net.net[0].weight), p(self.sigma_net.net[1].weight),
                          p(self.sigma_net.net[2].weight), p(self.color_net.net[0].weight), p(self.color_net.net[1].weight),
                          p(self.unc_net.net[0].weight) if with_unc else None, p(self.unc_net.net[1].weight) if with_unc else None)
        self._packed_weights = w            # keep the struct alive
        L.call("b2n_model_update", self._handle, ctypes.byref(w), torch.cuda.current_stream().cuda_stream)
        return self

    @property
    def handle(self):
        if self._handle is None:
            self.pack()
        return self._handle

    @torch.no_grad()
    def forward(self, x, d, enc_a, c, e, n_valid=None, out=None):
        """Fused head: returns sigma [M], color [M,3], ambient_aud [M,1], ambient_eye [M,1], uncertainty [M,1,1] (fp32).
        `out` may pass a previous return tuple to write into (no allocation)."""
        M = x.shape[0]
        x, d = x.float().contiguous(), d.float().contiguous()
        dev = x.device
        if out is not None:
            sig, rgb, aud, eye_o, unc = out[0], out[1], out[2].view(-1), out[3].view(-1), out[4].view(-1)
        else:
            sig, rgb = torch.empty(M, device=dev), torch.empty(M, 3, device=dev)
            aud, eye_o, unc = torch.empty(M, device=dev), torch.empty(M, device=dev), torch.empty(M, device=dev)
        f = lambda t: None if t is None else t.detach().float().contiguous().view(-1)
        enc_a, c, e = f(enc_a), f(c), f(e)
        self._keep = (enc_a, c, e)
        lib().call("b2n_head_forward", self.handle, x.data_ptr(), d.data_ptr(), M, enc_a.data_ptr(), None if c is None else c.data_ptr(),
                   None if e is None else e.data_ptr(), None if n_valid is None else n_valid.data_ptr(),
                   sig.data_ptr(), rgb.data_ptr(), aud.data_ptr(), eye_o.data_ptr(), unc.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return sig, rgb, aud[:, None], eye_o[:, None], unc[:, None, None]

    # ---- training forward with saved activations (csrc/fused_head.cu, SAVE instantiation) --------------------------------
    SAVED_WIDTHS = dict(x36=40, ha=64, he=16, hu=32, att=32, s_in=80, h1=64, h2=64, c_in=88, hc=64, misc=8)

    @torch.no_grad()
    def forward_train_fused(self, x, d, enc_a, c, e):
        """Fused head on ALL rows of x (no n_valid), keeping the fp16 activations the backward needs.  Call pack() after every weight update.
        Returns (sigma [M], rgb [M,3], ambient_aud [M], ambient_eye [M], unc [M], saved: dict of fp16 [M, width] tensors)."""
        M, dev = x.shape[0], x.device
        x, d = x.float().contiguous(), d.float().contiguous()
        with_unc = (not self.testing) and self.unc_loss
        saved = {k: torch.empty(M, w, dtype=torch.float16, device=dev) for k, w in self.SAVED_WIDTHS.items() if k != "hu" or with_unc}
        sv = _HeadSavedC(*[saved[k].data_ptr() if k in saved else None for k, _ in _HeadSavedC._fields_])
        sig, rgb = torch.empty(M, device=dev), torch.empty(M, 3, device=dev)
        aud, eye_o, unc = torch.empty(M, device=dev), torch.empty(M, device=dev), torch.empty(M, device=dev)
        f = lambda t: None if t is None else t.detach().float().contiguous().view(-1)
        enc_a, c, e = f(enc_a), f(c), f(e)
        self._keep_train = (enc_a, c, e, sv)
        lib().call("b2n_head_forward_train", self.handle, x.data_ptr(), d.data_ptr(), M, enc_a.data_ptr(), None if c is None else c.data_ptr(),
                   None if e is None else e.data_ptr(), sig.data_ptr(), rgb.data_ptr(), aud.data_ptr(), eye_o.data_ptr(), unc.data_ptr(),
                   ctypes.byref(sv), torch.cuda.current_stream().cuda_stream)
        return sig, rgb, aud, eye_o, unc, saved

    # ---- occupancy-grid maintenance (renderer.py:633-768, head branch) ----------------------------------------------------------
    @torch.no_grad()
    def mark_untrained_grid(self, poses, intrinsic, S=64):
        """NeRFRenderer.mark_untrained_grid (renderer.py:633-697; TrainerUtil.py:475 calls it before the first step): density_grid cells that none of the
        training cameras `poses` [B,4,4] (c2w; numpy or tensor) with `intrinsic` (fx, fy, cx, cy) sees are set to -1, which freezes them out of every later
        update_extra_state.  One kernel over the Morton-ordered grid per 1024 poses (csrc/occupancy.cu); `S` is accepted for signature compatibility."""
        import numpy as np
        if isinstance(poses, np.ndarray):
            poses = torch.from_numpy(poses)
        dev = self.density_grid.device
        poses = poses.to(dev).float().contiguous().view(-1, 4, 4)
        fx, fy, cx, cy = (float(v) for v in intrinsic)
        B = poses.shape[0]
        if B <= 1024:
            lib().call("b2n_mark_untrained_grid", poses.data_ptr(), B, fx, fy, cx, cy, self.cascade, self.grid_size, self.bound, self.density_grid.data_ptr(),
                       torch.cuda.current_stream().cuda_stream)
            return
        # more poses than one launch stages in shared memory: a cell is untrained iff EVERY chunk marks it
        seen = torch.zeros_like(self.density_grid, dtype=torch.bool)
        for h in range(0, B, 1024):
            probe = torch.zeros_like(self.density_grid)
            chunk = poses[h:h + 1024].contiguous()
            lib().call("b2n_mark_untrained_grid", chunk.data_ptr(), chunk.shape[0], fx, fy, cx, cy, self.cascade, self.grid_size, self.bound, probe.data_ptr(),
                       torch.cuda.current_stream().cuda_stream)
            seen |= probe >= 0
        self.density_grid[~seen] = -1

    @torch.no_grad()
    def update_extra_state(self, auds, eye=None, decay=0.95, density_thresh=10.0, density_scale=1.0, fused=True):
        """NeRFRenderer.update_extra_state for the head (renderer.py:699-766): evaluate sigma on the jittered 128^3 lattice of every cascade, dilate,
        EMA-max into density_grid, re-pack the bitfield (threshold min(mean_density, density_thresh)).
        fused=True: per cascade ONE point-generation kernel that writes the lattice in Morton order (jitter = torch.rand_like in the reference's
        point order, so the same seed gives the reference's points) and ONE fused head launch over the 2.1 M points whose sigma output is tmp_grid itself, then
        dilation + EMA-max + mean in one kernel and packbits against the device-side threshold in another (csrc/occupancy.cu) — the reference runs ~45 kernels
        per MLP call plus index_put / dilation / masked max / mean / packbits with a host read-back in between.
        fused=False keeps the reference's op-by-op graph on the drop-in ops (parity partner).  Returns mean_density (one D2H read, like the reference)."""
        import raymarching
        dev, G = self.density_grid.device, self.grid_size
        enc_a = self.encode_audio(auds)
        st = torch.cuda.current_stream().cuda_stream
        if fused:
            self.pack()
            n = G ** 3
            sig_all = torch.empty(self.cascade, n, device=dev)
            xyzs = torch.empty(n, 3, device=dev)
            if getattr(self, "_grid_dirs", None) is None or self._grid_dirs.device != dev:
                self._grid_dirs = torch.zeros(n, 3, device=dev); self._grid_dirs[:, 2] = 1.0
                self._grid_stats = torch.zeros(4, device=dev)
            f = lambda t: None if t is None else t.detach().float().contiguous().view(-1)
            enc_f, code_f, eye_f = f(enc_a), f(self.individual_codes[0:1]), f(eye)
            for cas in range(self.cascade):
                rnd = torch.rand(n, 3, device=dev)                          # == torch.rand_like(cas_xyzs) of the reference (same shape / dtype / device)
                lib().call("b2n_density_grid_points", rnd.data_ptr(), G, cas, self.bound, xyzs.data_ptr(), st)
                lib().call("b2n_head_forward", self.handle, xyzs.data_ptr(), self._grid_dirs.data_ptr(), n, enc_f.data_ptr(), code_f.data_ptr(),
                           None if eye_f is None else eye_f.data_ptr(), None, sig_all[cas].data_ptr(), None, None, None, None, st)
            lib().call("b2n_density_grid_update", sig_all.data_ptr(), float(density_scale), self.density_grid.data_ptr(), self.cascade, G, float(decay),
                       float(density_thresh), self.density_bitfield.data_ptr(), self._grid_stats.data_ptr(), st)
            self.mean_density = float(self._grid_stats[2].item())
            return self.mean_density
        if not hasattr(self, "_grid_coords") or self._grid_coords.device != dev:
            ar = torch.arange(G, dtype=torch.int32, device=dev)
            xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing="ij")
            self._grid_coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1).contiguous()        # [G^3, 3]
            self._grid_indices = raymarching.morton3D(self._grid_coords).long()
        coords, indices = self._grid_coords, self._grid_indices
        xyzs = 2 * coords.float() / (G - 1) - 1
        tmp_grid = torch.zeros_like(self.density_grid)
        for cas in range(self.cascade):
            bound = min(2 ** cas, self.bound)
            half = bound / G
            cas_xyzs = xyzs * (bound - half)
            cas_xyzs += (torch.rand_like(cas_xyzs) * 2 - 1) * half
            sig = self.density(cas_xyzs, enc_a, eye)["sigma"].reshape(-1).float()
            tmp_grid[cas, indices] = sig * density_scale
        tmp_grid = raymarching.morton3D_dilation(tmp_grid)
        valid = (self.density_grid >= 0) & (tmp_grid >= 0)
        self.density_grid[valid] = torch.maximum(self.density_grid[valid] * decay, tmp_grid[valid])
        self.mean_density = float(torch.mean(self.density_grid.clamp(min=0)).item())
        self.density_bitfield = raymarching.packbits(self.density_grid, min(self.mean_density, density_thresh), self.density_bitfield)
        return self.mean_density

    # ---- whole-frame inference (renderer.py:406-570) ----------------------------------------------------------------------
    @torch.no_grad()
    def render_frame(self, rays_o, rays_d, enc_a, ind_code=None, eye=None, bg_color=None, dt_gamma=1.0 / 256, max_steps=16, min_near=0.05,
                     T_thresh=1e-4, density_scale=1.0, out=None, head_ctas=0, workspace=None, aux=None, image_width=0):
        """run_cuda_for_inference without host syncs: returns image [N,3] (clamped, background-blended), weights_sum [N], depth [N].
        workspace: a caller-owned uint8 buffer of b2n_render_frame_workspace_bytes(N) — frames rendered concurrently on several streams need one each (default: one
        per model, i.e. calls must be stream-ordered); aux = (weights_sum [N], depth [N]) output buffers to reuse."""
        rays_o, rays_d = rays_o.float().contiguous().view(-1, 3), rays_d.float().contiguous().view(-1, 3)
        N, dev = rays_o.shape[0], rays_o.device
        L = lib()
        if workspace is None:
            need = int(L.raw("b2n_render_frame_workspace_bytes")(N))
            if getattr(self, "_ws", None) is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            workspace = self._ws
        image = out if out is not None else torch.empty(N, 3, device=dev)
        ws, depth = aux if aux is not None else (torch.empty(N, device=dev), torch.empty(N, device=dev))
        cfg = _RenderCfgC(self.bound, dt_gamma, min_near, T_thresh, density_scale, max_steps, self.cascade, self.grid_size,
                          (ctypes.c_float * 6)(*[float(v) for v in self.aabb_infer.tolist()]) if not hasattr(self, "_aabb_host") else self._aabb_host, int(head_ctas), int(image_width))
        f = lambda t: None if t is None else t.detach().float().contiguous().view(-1)
        enc_a, ind_code, eye, bg = f(enc_a), f(ind_code), f(eye), f(bg_color)
        self._keep = (enc_a, ind_code, eye, bg, cfg)
        L.call("b2n_render_frame", self.handle, ctypes.byref(cfg), rays_o.data_ptr(), rays_d.data_ptr(), N, self.density_bitfield.data_ptr(),
               enc_a.data_ptr(), None if ind_code is None else ind_code.data_ptr(), None if eye is None else eye.data_ptr(),
               None if bg is None else bg.data_ptr(), workspace.data_ptr(), image.data_ptr(), ws.data_ptr(), depth.data_ptr(),
               torch.cuda.current_stream().cuda_stream)
        return image, ws, depth

    def load_state_dict(self, *args, **kwargs):
        """The packed model (operand images + corner-quad image of the tables) is a snapshot: refresh it when new weights arrive."""
        r = super().load_state_dict(*args, **kwargs)
        if self._handle is not None and next(self.parameters()).is_cuda:
            self.pack()
        return r

    def cache_host_constants(self):
        """Read aabb_infer once (a D2H copy) so render_frame never synchronises."""
        self._aabb_host = (ctypes.c_float * 6)(*[float(v) for v in self.aabb_infer.tolist()])
        return self

    def __del__(self):
        try:
            if self._handle is not None:
                lib().raw("b2n_model_destroy")(self._handle)
        except Exception:
            pass

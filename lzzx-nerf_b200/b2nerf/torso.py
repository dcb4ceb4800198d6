"""Torso branch of the talking-head NeRF on the B200-native kernels.

`TorsoModel` holds the torso parameters / buffers of the reference's `NeRFNetwork(NeRFRenderer)` under the SAME names and shapes
(nerf_triplane/network.py:156-167, renderer.py:123-149), so the torso part of a reference checkpoint loads with `strict=False`:
    anchor_points [3,4], torso_deform_net.net.{0,1,2}.weight, torso_encoder.embeddings / .offsets, torso_net.net.{0,1,2}.weight,
    individual_codes_torso [ind_num, 8], density_grid_torso [grid_size^2]  (+ the python float mean_density_torso).

Two evaluation paths, same math (run_torso, renderer.py:572-631 with forward_torso, network.py:170-205):
  * run_torso(...)        — the reference's op-by-op graph on the drop-in encoders (freqencoder / gridencoder kernels) + torch Linear; the parity partner;
  * run_torso_fused(...)  — ONE kernel (csrc/fused_torso.cu, b2n_torso_forward) through the C ABI.
Both return the reference's dict: bg_color [N,3] (torso over background — what the head composite takes as bg_color), torso_alpha [N,1], torso_color = bg_color.
"""
import ctypes
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from freqencoder import FreqEncoder
from gridencoder import GridEncoder

from ._lib import lib
from .model import MLP


class _TorsoWeightsC(ctypes.Structure):     # mirrors b2n_torso_weights (include/b2nerf_fused.h)
    _fields_ = [("deform_w0", ctypes.c_void_p), ("deform_w1", ctypes.c_void_p), ("deform_w2", ctypes.c_void_p),
                ("torso_w0", ctypes.c_void_p), ("torso_w1", ctypes.c_void_p), ("torso_w2", ctypes.c_void_p),
                ("table", ctypes.c_void_p), ("offsets", ctypes.c_void_p), ("S", ctypes.c_float), ("H", ctypes.c_uint32), ("torso_shrink", ctypes.c_float)]


def get_bg_coords(H, W, device):
    """utils.py:218-223: [1, H*W, 2] in [-1, 1] (first coordinate runs over rows)."""
    X = torch.arange(H, device=device) / (H - 1) * 2 - 1
    Y = torch.arange(W, device=device) / (W - 1) * 2 - 1
    xs, ys = torch.meshgrid(X, Y, indexing="ij")
    return torch.cat([xs.reshape(-1, 1), ys.reshape(-1, 1)], dim=-1).unsqueeze(0)


class _TorsoOperandsC(ctypes.Structure):    # mirrors b2n_torso_operands (include/b2nerf_fused.h)
    _fields_ = [(n, ctypes.c_void_p) for n in ("x_t0", "x_t1", "x_t2", "x_d0", "x_d1", "x_d2", "dy_t2", "dy_t1", "dy_t0", "dy_d2", "dy_d1", "dy_d0")]


class _FusedTorso(torch.autograd.Function):
    """forward_torso + run_torso's blend (network.py:170-205, renderer.py:572-631) as one autograd node on csrc/fused_torso.cu."""

    OPERAND_WIDTHS = dict(x_t0=120, x_t1=32, x_t2=32, x_d0=88, x_d1=32, x_d2=32, dy_t2=8, dy_t1=32, dy_t0=32, dy_d2=8, dy_d1=32, dy_d0=32)

    @staticmethod
    def forward(ctx, model, bg_coords, h_const, bg_color, wd0, wd1, wd2, wt0, wt1, wt2, table):
        N, dev = bg_coords.shape[0], bg_coords.device
        per_ray = int(bg_color is not None and bg_color.numel() == 3 * N and N > 1)
        out, alpha = torch.empty(N, 3, device=dev), torch.empty(N, device=dev)
        w = model.weights_struct()
        ws = torch.empty(int(lib().raw("b2n_torso_workspace_bytes")()), dtype=torch.uint8, device=dev)
        hc = h_const.detach().float().contiguous().view(-1)
        lib().call("b2n_torso_forward", ctypes.byref(w), bg_coords.data_ptr(), N, model.density_grid_torso.data_ptr(), model.grid_size, float(model.density_thresh()),
                   hc.data_ptr(), None if bg_color is None else bg_color.data_ptr(), per_ray, out.data_ptr(), alpha.data_ptr(), None, ws.data_ptr(),
                   torch.cuda.current_stream().cuda_stream)
        ctx.model, ctx.per_ray, ctx.thresh = model, per_ray, float(model.density_thresh())
        ctx.save_for_backward(bg_coords, hc, bg_color if bg_color is not None else torch.empty(0, device=dev))
        ctx.has_bg = bg_color is not None
        return out, alpha

    @staticmethod
    def backward(ctx, g_out, g_alpha):
        from .fused_train import _wgrad_all
        model = ctx.model
        bg_coords, hc, bg_color = ctx.saved_tensors
        N, dev = bg_coords.shape[0], bg_coords.device
        Np = -(-N // 256) * 256
        ops = {k: torch.empty(Np, wd, dtype=torch.float16, device=dev) for k, wd in _FusedTorso.OPERAND_WIDTHS.items()}
        if Np > N:
            for t in ops.values():
                t[N:].zero_()
        oc = _TorsoOperandsC(*[ops[k].data_ptr() for k, _ in _TorsoOperandsC._fields_])
        table = model.torso_encoder.embeddings
        g_table = torch.zeros(table.shape, dtype=torch.float32, device=dev)
        w = model.weights_struct()
        ws = torch.empty(int(lib().raw("b2n_torso_workspace_bytes")()), dtype=torch.uint8, device=dev)
        g_out = g_out.float().contiguous()
        g_alpha = None if g_alpha is None else g_alpha.float().contiguous()
        lib().call("b2n_torso_backward", ctypes.byref(w), bg_coords.data_ptr(), N, model.density_grid_torso.data_ptr(), model.grid_size, ctx.thresh, hc.data_ptr(),
                   bg_color.data_ptr() if ctx.has_bg else None, ctx.per_ray, g_out.data_ptr(), None if g_alpha is None else g_alpha.data_ptr(), g_table.data_ptr(),
                   ctypes.byref(oc), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
        dw = _wgrad_all([(ops["dy_d0"], ops["x_d0"]), (ops["dy_d1"], ops["x_d1"]), (ops["dy_d2"], ops["x_d2"]),
                         (ops["dy_t0"], ops["x_t0"]), (ops["dy_t1"], ops["x_t1"]), (ops["dy_t2"], ops["x_t2"])])
        g_wd0, g_wd1, g_wd2 = dw[0][:, :84].contiguous(), dw[1], dw[2][:2].contiguous()
        g_wt0, g_wt1, g_wt2 = dw[3][:, :116].contiguous(), dw[4], dw[5][:4].contiguous()
        # the 50 per-frame constant inputs of both first layers: d h_const = W[:, const]^T colsum(dY) (the weights as the kernels see them: rounded to fp16)
        d, t = model.torso_deform_net.net, model.torso_net.net
        s_t, s_d = ops["dy_t0"][:N].float().sum(0), ops["dy_d0"][:N].float().sum(0)
        g_hc = (t[0].weight.detach()[:, 66:116].half().float().t() @ s_t + d[0].weight.detach()[:, 34:84].half().float().t() @ s_d).view(1, -1)
        return None, None, g_hc, None, g_wd0, g_wd1, g_wd2, g_wt0, g_wt1, g_wt2, g_table


class TorsoModel(nn.Module):
    def __init__(self, ind_dim_torso=8, ind_num=10000, grid_size=128, torso_shrink=0.8, density_thresh_torso=0.01):
        super().__init__()
        self.grid_size, self.torso_shrink, self.density_thresh_torso = grid_size, float(torso_shrink), float(density_thresh_torso)
        self.individual_dim_torso = ind_dim_torso
        self.anchor_points = nn.Parameter(torch.tensor([[0.01, 0.01, 0.1, 1], [-0.1, -0.1, 0.1, 1], [0.1, -0.1, 0.1, 1]]))
        self.torso_deform_encoder = FreqEncoder(input_dim=2, degree=8)          # 34
        self.anchor_encoder = FreqEncoder(input_dim=6, degree=3)                # 42
        self.torso_deform_net = MLP(34 + 42 + ind_dim_torso, 2, 32, 3)
        self.torso_encoder = GridEncoder(input_dim=2, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16, desired_resolution=2048,
                                         gridtype="tiled")                    # 32
        self.torso_net = MLP(32 + 34 + 42 + ind_dim_torso, 4, 32, 3)
        self.individual_codes_torso = nn.Parameter(torch.randn(ind_num, ind_dim_torso) * 0.1)
        self.register_buffer("density_grid_torso", torch.zeros(grid_size ** 2))
        self.mean_density_torso = 0.0

    # ---- per-frame constants ------------------------------------------------------------------------------------------------------------
    def frame_constants(self, poses, index=0):
        """[1, 50] = [anchor_encoder(wrapped anchors) | individual code]: the inputs of both MLPs that do not depend on the pixel (network.py:179-190)."""
        wrapped = self.anchor_points[None, ...] @ poses.permute(0, 2, 1).inverse()
        wrapped = (wrapped[:, :, :2] / wrapped[:, :, 3, None] / wrapped[:, :, 2, None]).view(1, -1)
        enc_anchor = self.anchor_encoder(wrapped.float())
        c = self.individual_codes_torso[index if self.training else 0].view(1, -1)
        return torch.cat([enc_anchor, c], dim=-1)

    # ---- op-by-op path (the reference graph) ------------------------------------------------------------------------------------------------
    def forward_torso(self, x, h_const):
        """network.py:170-205 with [enc_anchor | c] precomputed.  x [M,2] in [-1,1] -> alpha [M,1], color [M,3], dx [M,2]."""
        x = x * self.torso_shrink
        enc_x = self.torso_deform_encoder(x)
        h = torch.cat([enc_x, h_const.repeat(x.shape[0], 1)], dim=-1)
        dx = self.torso_deform_net(h)
        x = (x + dx).clamp(-1, 1)
        x = self.torso_encoder(x, bound=1)
        h = torch.cat([x, h], dim=-1)
        h = self.torso_net(h)
        alpha = torch.sigmoid(h[..., :1]) * (1 + 2 * 0.001) - 0.001
        color = torch.sigmoid(h[..., 1:]) * (1 + 2 * 0.001) - 0.001
        return alpha, color, dx

    def density_thresh(self):
        return min(self.density_thresh_torso, self.mean_density_torso)

    def run_torso(self, bg_coords, poses, index=0, bg_color=None):
        """renderer.py:572-631 on the drop-in ops (call under torch.autocast like the reference does)."""
        bg_coords = bg_coords.contiguous().view(-1, 2)
        N, dev = bg_coords.shape[0], bg_coords.device
        if bg_color is None:
            bg_color = 1
        occupancy = F.grid_sample(self.density_grid_torso.view(1, 1, self.grid_size, self.grid_size), bg_coords.view(1, -1, 1, 2), align_corners=True).view(-1)
        mask = occupancy > self.density_thresh()
        torso_alpha, torso_color = torch.zeros([N, 1], device=dev), torch.zeros([N, 3], device=dev)
        results = {}
        if mask.any():
            a, c, deform = self.forward_torso(bg_coords[mask], self.frame_constants(poses, index))
            torso_alpha[mask] = a.float()
            torso_color[mask] = c.float()
            results["deform"] = deform
        bg_color = torso_color * torso_alpha + bg_color * (1 - torso_alpha)
        results.update(torso_alpha=torso_alpha, torso_color=bg_color, bg_color=bg_color, mask=mask)
        return results

    # ---- fused path -------------------------------------------------------------------------------------------------------------------------
    def weights_struct(self):
        p = lambda t: t.detach().data_ptr()
        for t in (self.torso_encoder.embeddings, self.torso_net.net[0].weight, self.torso_deform_net.net[0].weight):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError("TorsoModel: parameters must be contiguous float32 CUDA tensors")
        d, t, e = self.torso_deform_net.net, self.torso_net.net, self.torso_encoder
        return _TorsoWeightsC(p(d[0].weight), p(d[1].weight), p(d[2].weight), p(t[0].weight), p(t[1].weight), p(t[2].weight), p(e.embeddings), p(e.offsets),
                              float(math.log2(e.per_level_scale)), e.base_resolution, self.torso_shrink)

    @torch.no_grad()
    def run_torso_fused(self, bg_coords, poses, index=0, bg_color=None, h_const=None, want_deform=False, out=None, workspace=None):
        """Same results as run_torso under autocast(fp16), one kernel.  bg_color: None (white), [3] / [1,3] or [N,3]; out: optional [N,3] result buffer
        (FrameRenderer passes the buffer its frame graph reads as bg_color)."""
        bg_coords = bg_coords.contiguous().view(-1, 2)
        if not (bg_coords.is_cuda and bg_coords.dtype == torch.float32):
            raise RuntimeError("run_torso_fused: bg_coords must be a float32 CUDA tensor (there is no CPU path)")
        N, dev = bg_coords.shape[0], bg_coords.device
        if h_const is None:
            h_const = self.frame_constants(poses, index)
        h_const = h_const.float().contiguous().view(-1)
        per_ray = 0
        if bg_color is not None:
            bg_color = torch.as_tensor(bg_color, dtype=torch.float32, device=dev).contiguous()
            if bg_color.numel() == 1:
                bg_color = bg_color.expand(3).contiguous()
            per_ray = int(bg_color.numel() == 3 * N and N > 1)
            if not per_ray and bg_color.numel() != 3:
                raise RuntimeError("run_torso_fused: bg_color must have 1, 3 or N*3 elements")
        out = torch.empty(N, 3, device=dev) if out is None else out
        if not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and out.numel() == 3 * N):
            raise RuntimeError("run_torso_fused: out must be a contiguous float32 CUDA tensor [N,3]")
        alpha = torch.empty(N, device=dev)
        deform = torch.empty(N, 2, device=dev) if want_deform else None
        w = self.weights_struct()
        if workspace is None:          # packed weight image + tile counter; concurrent callers (frames in flight on several streams) pass their own
            workspace = torch.empty(int(lib().raw("b2n_torso_workspace_bytes")()), dtype=torch.uint8, device=dev)
        lib().call("b2n_torso_forward", ctypes.byref(w), bg_coords.data_ptr(), N, self.density_grid_torso.data_ptr(), self.grid_size, float(self.density_thresh()),
                   h_const.data_ptr(), None if bg_color is None else bg_color.data_ptr(), per_ray, out.data_ptr(), alpha.data_ptr(),
                   None if deform is None else deform.data_ptr(), workspace.data_ptr(), torch.cuda.current_stream().cuda_stream)
        res = dict(torso_alpha=alpha.view(N, 1), torso_color=out, bg_color=out)
        if deform is not None:
            res["deform"] = deform
        return res

    # ---- fused training path -------------------------------------------------------------------------------------------------------
    def run_torso_train_fused(self, bg_coords, poses, index=0, bg_color=None):
        """run_torso for the torso TRAINING stage (TrainerUtil.py:188-236 with opt.torso) on two kernels + the head's weight-gradient kernel: forward =
        k_torso_frame, backward = k_torso_backward (recompute + backward-data + table gradients) and ONE b2n_linear_wgrad_batch launch for the six weight
        matrices, instead of autograd over ~25 forward and ~60 backward launches of the op-by-op graph.  Differentiable w.r.t. every torso parameter
        (anchor_points and individual_codes_torso through frame_constants).  Returns the reference's dict (bg_color / torso_color [N,3], torso_alpha [N,1])."""
        bg_coords = bg_coords.contiguous().view(-1, 2).float()
        h_const = self.frame_constants(poses, index).float()                  # [1, 50], differentiable (anchor encoder + individual code)
        N, dev = bg_coords.shape[0], bg_coords.device
        if bg_color is not None:
            bg_color = torch.as_tensor(bg_color, dtype=torch.float32, device=dev).contiguous()
            if bg_color.numel() == 1:
                bg_color = bg_color.expand(3).contiguous()
            if bg_color.numel() not in (3, 3 * N):
                raise RuntimeError("run_torso_train_fused: bg_color must have 1, 3 or N*3 elements")
        d, t = self.torso_deform_net.net, self.torso_net.net
        out, alpha = _FusedTorso.apply(self, bg_coords, h_const, bg_color, d[0].weight, d[1].weight, d[2].weight, t[0].weight, t[1].weight, t[2].weight,
                                       self.torso_encoder.embeddings)
        return dict(torso_alpha=alpha.view(N, 1), torso_color=out, bg_color=out)

    # ---- 2-D occupancy refresh ----------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def update_extra_state(self, pose, index=0, decay=0.95, fused=True, noise=None):
        """The torso part of NeRFRenderer.update_extra_state (renderer.py:768-809): alpha of the jittered grid_size^2 lattice for one (pose, individual
        code), 5x5 max-pool dilation, EMA-max into density_grid_torso, mean -> mean_density_torso (one .item(), as in the reference).  fused=True evaluates
        the lattice with the fused kernel (threshold -inf: every point is evaluated), fused=False through forward_torso under autocast.  `noise` [G*G, 2] in
        [0, 1) replaces torch.rand_like (tests)."""
        G, dev = self.grid_size, self.density_grid_torso.device
        ar = torch.arange(G, dtype=torch.int32, device=dev)
        xx, yy = torch.meshgrid(ar, ar, indexing="ij")
        coords = torch.cat([xx.reshape(-1, 1), yy.reshape(-1, 1)], dim=-1)                        # [N, 2] in [0, G)
        indices = (coords[:, 1] * G + coords[:, 0]).long()                                        # NOTE: xy transposed (renderer.py:789)
        half = 1 / G
        xys = (2 * coords.float() / (G - 1) - 1) * (1 - half)
        xys = xys + ((torch.rand_like(xys) if noise is None else noise.to(dev)) * 2 - 1) * half
        ind_was, self.training = self.training, True                                              # the refresh uses the frame's own code (renderer.py:777)
        try:
            h_const = self.frame_constants(pose.to(dev), index)
        finally:
            self.training = ind_was
        if fused:
            thresh_was, mean_was = self.density_thresh_torso, self.mean_density_torso
            self.density_thresh_torso = self.mean_density_torso = -1e30                           # occupancy > -inf: evaluate every lattice point
            try:
                alphas = self.run_torso_fused(xys.contiguous(), None, 0, None, h_const=h_const)["torso_alpha"].view(-1)
            finally:
                self.density_thresh_torso, self.mean_density_torso = thresh_was, mean_was
        else:
            with torch.autocast("cuda", dtype=torch.float16, enabled=xys.is_cuda):
                alphas = self.forward_torso(xys, h_const)[0].squeeze(1).float()
        tmp = torch.zeros_like(self.density_grid_torso)
        tmp[indices] = alphas
        tmp = F.max_pool2d(tmp.view(1, 1, G, G), kernel_size=5, stride=1, padding=2).view(-1)
        self.density_grid_torso.copy_(torch.maximum(self.density_grid_torso * decay, tmp))
        self.mean_density_torso = torch.mean(self.density_grid_torso).item()
        return self.mean_density_torso


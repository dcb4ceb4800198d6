"""Frame-level public API: what a serving loop calls per 512x512 frame (the role of TrainerUtil.test_step ->
NeRFRenderer.run_cuda_for_inference, TrainerUtil.py:408-460 / renderer.py:406-570).

    r = FrameRenderer(model, n_rays=512 * 512)
    r.render_host(rays_o_pinned, rays_d_pinned, auds_pinned, out_pinned)       # H2D -> audio encode -> frame -> D2H

The device work of one frame (audio encoder + the fixed launch sequence of b2n_render_frame) is captured once into a CUDA
graph and replayed, so the host issues three copies and one graph launch per frame and never synchronises inside a frame.
"""
import torch

from .model import HeadModel


class FrameRenderer:
    def __init__(self, model: HeadModel, n_rays, eye=0.4, ind_index=0, dt_gamma=1.0 / 256, max_steps=16, T_thresh=1e-4, use_graph=True, fused_audio=True):
        self.m = model
        self.dev = next(model.parameters()).device
        self.N = int(n_rays)
        self.kw = dict(dt_gamma=dt_gamma, max_steps=max_steps, T_thresh=T_thresh)
        d = self.dev
        self.rays_o = torch.empty(self.N, 3, device=d)
        self.rays_d = torch.empty(self.N, 3, device=d)
        self.auds = torch.empty(8, model.audio_in_dim, 2 if model.audio_in_dim == 1024 else 16, device=d)
        self.eye = torch.tensor([[float(eye)]], device=d)
        self.ind_code = model.individual_codes[ind_index:ind_index + 1].detach().clone()
        self.image = torch.empty(self.N, 3, device=d)
        self.enc_a = torch.empty(1, 32, device=d)
        self.fused_audio = fused_audio and model.att > 0
        self.graph = None
        self.launches_per_frame = None
        model.cache_host_constants()
        model.pack()
        if use_graph:
            self._capture()

    def _device_frame(self):
        if self.fused_audio:
            enc_a = self.m.encode_audio_fused(self.auds, out=self.enc_a)      # one cluster kernel (csrc/fused_audio.cu)
        else:
            with torch.autocast("cuda", dtype=torch.float16):
                enc_a = self.m.encode_audio(self.auds).float()              # AudioNet + AudioAttNet through torch (network.py:226-240)
        self.m.render_frame(self.rays_o, self.rays_d, enc_a, self.ind_code, self.eye, out=self.image, **self.kw)

    @torch.no_grad()
    def _capture(self):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):                                   # warm-up: lazy inits (cuDNN plans, func attributes, workspace growth)
                self._device_frame()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        from ._lib import lib
        g = torch.cuda.CUDAGraph()
        n0 = lib().launch_count()
        with torch.cuda.graph(g):
            self._device_frame()
        self.launches_per_frame = lib().launch_count() - n0      # libb2nerf kernel nodes replayed per frame (torch's audio-net kernels not counted)
        self.graph = g

    @torch.no_grad()
    def render_device(self, rays_o=None, rays_d=None, auds=None):
        """Inputs already on the device (copied into the renderer's static buffers); returns the static image buffer [N,3]."""
        if rays_o is not None:
            self.rays_o.copy_(rays_o.view(-1, 3), non_blocking=True)
            self.rays_d.copy_(rays_d.view(-1, 3), non_blocking=True)
        if auds is not None:
            self.auds.copy_(auds, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._device_frame()
        return self.image

    @torch.no_grad()
    def render_host(self, rays_o_host, rays_d_host, auds_host, out_host):
        """Host (pinned) buffers in and out: 2 x N x 12 B + audio window up, N x 12 B down, per frame."""
        self.rays_o.copy_(rays_o_host.view(-1, 3), non_blocking=True)
        self.rays_d.copy_(rays_d_host.view(-1, 3), non_blocking=True)
        self.auds.copy_(auds_host, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._device_frame()
        out_host.copy_(self.image, non_blocking=True)
        return out_host

    def h2d_bytes(self):
        return self.rays_o.numel() * 4 * 2 + self.auds.numel() * 4

    def d2h_bytes(self):
        return self.image.numel() * 4

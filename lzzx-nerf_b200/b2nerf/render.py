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
    def __init__(self, model: HeadModel, n_rays, eye=0.4, ind_index=0, dt_gamma=1.0 / 256, max_steps=16, T_thresh=1e-4, use_graph=True, fused_audio=True,
                 camera=None, torso=None, bg_coords=None, smooth_lips=False, lips_state=None, lips_lambda=0.35, head_ctas=0, image_width=0):
        """image_width: b2n_render_cfg.image_width — the rays handed to render_device / render_host are the row-major pixels of an image of this width (0: unknown;
        taken from `camera` when that is given).  head_ctas: b2n_render_cfg.head_ctas — 0 when this renderer has the GPU to itself; FramePipeline sets it for its slots.
        smooth_lips: the reference's opt.smooth_lips (renderer.py:456-460, on in the serving config HubertInferenceMQ.py): the audio code of frame k is
        0.35 * (smoothed code of frame k-1) + 0.65 * (its own); the state is a device float[33] (`lips_state`, shared by the slots of a FramePipeline) and the
        audio kernel runs in frame order in front of the frame graph instead of inside it.  reset_lips() starts a new sequence.
        camera = (H, W, fx, fy, cx, cy): also build the device-side prologue / epilogue (rays from a 4x4 pose, RGB24 output), so that
        render_host_pose() moves a pose + the audio window up and one uint8 frame down.
        torso = a TorsoModel (+ bg_coords [N,2], utils.py:218-223): every frame first runs the fused torso kernel (csrc/fused_torso.cu) over the background into
        the per-ray bg_color buffer the head frame reads (renderer.py:572-631 then :559-561); set the head pose with set_torso_pose()."""
        self.m = model
        self.head_ctas = int(head_ctas)
        self.image_width = int(camera[1]) if camera is not None else int(image_width)
        self.torso = torso
        self.camera = camera
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
        self.smooth_lips = bool(smooth_lips)
        if self.smooth_lips:
            if not self.fused_audio:
                raise RuntimeError("FrameRenderer: smooth_lips needs the fused audio encoder (att > 0)")
            self.lips_state = lips_state if lips_state is not None else torch.zeros(33, device=d)
            self.lips_lambda = float(lips_lambda)
            self._audio_w = model.audio_weights_struct()
        self._audio_done = False        # smooth_lips: a FramePipeline already ran the audio kernel for the coming frame (in frame order, on its audio stream)
        self.graph = None               # torch.cuda.CUDAGraph of the fixed launch sequence (fallback)
        self.loop_graph = None          # libb2nerf frame graph with a device-controlled WHILE loop (preferred)
        self.launches_per_frame = None
        self.ws = torch.empty(self.N, device=d)
        self.depth = torch.empty(self.N, device=d)
        self.bg = None
        if torso is not None:
            if bg_coords is None or bg_coords.numel() != 2 * self.N:
                raise RuntimeError("FrameRenderer: torso needs bg_coords [n_rays, 2]")
            self.bg_coords = bg_coords.to(d).float().contiguous().view(-1, 2)
            self.bg = torch.ones(self.N, 3, device=d)                       # torso over white; the frame reads it as its per-ray bg_color
            self.torso_bg_color = None
            self.torso_h_const = torch.zeros(50, device=d)
            from ._lib import lib as _lib
            self.torso_ws = torch.empty(int(_lib().raw("b2n_torso_workspace_bytes")()), dtype=torch.uint8, device=d)
            self.set_torso_pose(torch.eye(4, device=d)[None])
        if camera is not None:
            if int(camera[0]) * int(camera[1]) != self.N:
                raise RuntimeError("FrameRenderer: camera H*W must equal n_rays")
            self.pose = torch.eye(4, device=d).contiguous()
            self.rgb8 = torch.empty(self.N, 3, dtype=torch.uint8, device=d)
            self.pose_graph = None
        model.cache_host_constants()
        model.pack()
        if use_graph:
            try:
                import os
                if os.environ.get("B2N_FRAME_NO_WHILE") == "1":       # measurement switch: the fixed 16-iteration sequence captured as an ordinary graph
                    raise RuntimeError("WHILE-node graph disabled by B2N_FRAME_NO_WHILE")
                self._build_loop_graph()
            except RuntimeError as e:          # conditional graph nodes unavailable: capture the fixed sequence instead (still a GPU path)
                self.loop_graph_error = str(e)
                self._capture()

    @torch.no_grad()
    def set_torso_pose(self, poses, index=0):
        """Head pose [1,4,4] of the coming frames -> the 50 per-frame constant inputs of the torso MLPs (network.py:179-190)."""
        self.torso_h_const.copy_(self.torso.frame_constants(poses.to(self.dev), index).view(-1))

    def reset_lips(self):
        """Start a new frame sequence (the reference's `self.enc_a = None`, renderer.py:151-152)."""
        if self.smooth_lips:
            self.lips_state.zero_()

    def encode_audio_smoothed(self):
        """k_audio_encode + the smooth_lips EMA on the current stream: self.auds -> self.enc_a, lips_state updated in place."""
        import ctypes
        from ._lib import lib
        lib().call("b2n_audio_encode_smooth", ctypes.byref(self._audio_w), self.auds.data_ptr(), self.auds.shape[2], self.enc_a.data_ptr(),
                   self.lips_state.data_ptr(), self.lips_lambda, torch.cuda.current_stream(self.dev).cuda_stream)

    def _torso(self):
        if self.torso is not None:
            self.torso.run_torso_fused(self.bg_coords, None, 0, self.torso_bg_color, h_const=self.torso_h_const, out=self.bg, workspace=self.torso_ws)

    def _device_frame(self):
        self._torso()
        if self.smooth_lips:
            if not self._audio_done:
                self.encode_audio_smoothed()
            enc_a = self.enc_a
        elif self.fused_audio:
            enc_a = self.m.encode_audio_fused(self.auds, out=self.enc_a)      # one cluster kernel (csrc/fused_audio.cu)
        else:
            with torch.autocast("cuda", dtype=torch.float16):
                enc_a = self.m.encode_audio(self.auds).float()              # AudioNet + AudioAttNet through torch (network.py:226-240)
        # eager frames use this renderer's own loop workspace and output buffers: several renderers may run concurrently on their own streams
        if getattr(self, "_eager_ws", None) is None:
            from ._lib import lib as _lib
            self._eager_ws = torch.empty(int(_lib().raw("b2n_render_frame_workspace_bytes")(self.N)), dtype=torch.uint8, device=self.dev)
        self.m.render_frame(self.rays_o, self.rays_d, enc_a, self.ind_code, self.eye, bg_color=self.bg, out=self.image, head_ctas=self.head_ctas,
                            workspace=self._eager_ws, aux=(self.ws, self.depth), image_width=self.image_width, **self.kw)

    @torch.no_grad()
    def _build_loop_graph(self):
        import ctypes
        from ._lib import lib
        from .model import _RenderCfgC
        L, m = lib(), self.m
        need = int(L.raw("b2n_render_frame_workspace_bytes")(self.N))
        self._graph_ws = torch.empty(need, dtype=torch.uint8, device=self.dev)
        cfg = _RenderCfgC(m.bound, self.kw["dt_gamma"], 0.05, self.kw["T_thresh"], 1.0, self.kw["max_steps"], m.cascade, m.grid_size, m._aabb_host, self.head_ctas, self.image_width)
        aw = m.audio_weights_struct() if (self.fused_audio and not self.smooth_lips) else None        # smooth_lips: the audio kernel runs in frame order, in front of the graph
        self._graph_keep = (cfg, aw, self.ind_code.float().contiguous().view(-1), self.eye.float().contiguous().view(-1))
        h = ctypes.c_void_p()
        torch.cuda.synchronize(self.dev)
        L.call("b2n_frame_graph_create", ctypes.byref(h), m.handle, ctypes.byref(cfg), ctypes.byref(aw) if aw is not None else None,
               self.auds.data_ptr(), self.auds.shape[2], self.enc_a.data_ptr(), self.rays_o.data_ptr(), self.rays_d.data_ptr(), self.N,
               m.density_bitfield.data_ptr(), self._graph_keep[2].data_ptr(), self._graph_keep[3].data_ptr(), None if self.bg is None else self.bg.data_ptr(),
               self._graph_ws.data_ptr(),
               self.image.data_ptr(), self.ws.data_ptr(), self.depth.data_ptr())
        self.loop_graph = h
        kf, kb = ctypes.c_uint64(), ctypes.c_uint64()
        L.call("b2n_frame_graph_info", h, ctypes.byref(kf), ctypes.byref(kb))
        self.kernels_fixed, self.kernels_per_iteration = int(kf.value), int(kb.value)

    @torch.no_grad()
    def _build_pose_graph(self):
        """Second graph: pose -> rays (k_frame_rays) -> the same frame -> float image + RGB24 (k_frame_finish)."""
        import ctypes
        from ._lib import lib
        L, m = lib(), self.m
        H, W, fx, fy, cx, cy = self.camera
        cfg, aw, ind, eye = self._graph_keep
        io = _FrameIoC(self.pose.data_ptr(), float(fx), float(fy), float(cx), float(cy), int(H), int(W), self.rgb8.data_ptr())
        self._pose_keep = io
        h = ctypes.c_void_p()
        torch.cuda.synchronize(self.dev)
        L.call("b2n_frame_graph_create_io", ctypes.byref(h), m.handle, ctypes.byref(cfg), ctypes.byref(aw) if aw is not None else None,
               self.auds.data_ptr(), self.auds.shape[2], self.enc_a.data_ptr(), self.rays_o.data_ptr(), self.rays_d.data_ptr(), self.N,
               m.density_bitfield.data_ptr(), ind.data_ptr(), eye.data_ptr(), None if self.bg is None else self.bg.data_ptr(), self._graph_ws.data_ptr(),
               self.image.data_ptr(), self.ws.data_ptr(), self.depth.data_ptr(), ctypes.byref(io))
        self.pose_graph = h

    @torch.no_grad()
    def render_host_pose(self, pose_host, auds_host, out_u8_host):
        """Pinned host pose [4,4] + audio window in, pinned uint8 [N,3] frame out: 64 B + the audio window up, N x 3 bytes down per frame."""
        from ._lib import lib
        if self.camera is None or self.loop_graph is None or not self.fused_audio:
            raise RuntimeError("render_host_pose needs FrameRenderer(camera=...), the loop graph and the fused audio encoder")
        if self.pose_graph is None:
            self._build_pose_graph()
        self.pose.copy_(pose_host.view(4, 4), non_blocking=True)
        if auds_host is not None:
            self.auds.copy_(auds_host, non_blocking=True)
        self._torso()
        if self.smooth_lips and not self._audio_done:
            self.encode_audio_smoothed()
        lib().call("b2n_frame_graph_launch", self.pose_graph, torch.cuda.current_stream(self.dev).cuda_stream)
        out_u8_host.copy_(self.rgb8, non_blocking=True)
        return out_u8_host

    def last_iterations(self):
        """Loop iterations the last frame executed (synchronises)."""
        import ctypes
        from ._lib import lib
        ws = self._graph_ws if self.loop_graph is not None else self._eager_ws
        it = ctypes.c_int32()
        lib().call("b2n_frame_iterations", ws.data_ptr(), self.N, ctypes.byref(it), torch.cuda.current_stream(self.dev).cuda_stream)
        return int(it.value)

    def _launch(self):
        if self.loop_graph is not None:
            from ._lib import lib
            self._torso()
            if self.smooth_lips and not self._audio_done:
                self.encode_audio_smoothed()
            if not self.fused_audio:
                with torch.autocast("cuda", dtype=torch.float16):
                    self.enc_a.copy_(self.m.encode_audio(self.auds).float())
            lib().call("b2n_frame_graph_launch", self.loop_graph, torch.cuda.current_stream(self.dev).cuda_stream)
        elif self.graph is not None:
            self.graph.replay()
        else:
            self._device_frame()

    def __del__(self):
        try:
            if self.loop_graph is not None:
                from ._lib import lib
                lib().raw("b2n_frame_graph_destroy")(self.loop_graph)
                if getattr(self, "pose_graph", None) is not None:
                    lib().raw("b2n_frame_graph_destroy")(self.pose_graph)
        except Exception:
            pass

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
        self._launch()
        return self.image

    @torch.no_grad()
    def render_host(self, rays_o_host, rays_d_host, auds_host, out_host):
        """Host (pinned) buffers in and out: 2 x N x 12 B + audio window up, N x 12 B down, per frame."""
        self.rays_o.copy_(rays_o_host.view(-1, 3), non_blocking=True)
        self.rays_d.copy_(rays_d_host.view(-1, 3), non_blocking=True)
        if auds_host is not None:
            self.auds.copy_(auds_host, non_blocking=True)
        self._launch()
        out_host.copy_(self.image, non_blocking=True)
        return out_host

    def h2d_bytes(self):
        return self.rays_o.numel() * 4 * 2 + self.auds.numel() * 4

    def d2h_bytes(self):
        return self.image.numel() * 4


import ctypes as _ct


class _FrameIoC(_ct.Structure):
    """b2n_frame_io (include/b2nerf_fused.h)"""
    _fields_ = [("pose", _ct.c_void_p), ("fx", _ct.c_float), ("fy", _ct.c_float), ("cx", _ct.c_float), ("cy", _ct.c_float), ("H", _ct.c_uint32),
                ("W", _ct.c_uint32), ("rgb8_out", _ct.c_void_p)]


class FramePipeline:
    """`depth` frames in flight on one GPU: frame k runs on renderer/stream k % depth, so the host<->device copies and the kernel
    tails / dependent-launch gaps of one frame overlap the next frame's work (frames are independent — the reference's test loop,
    TrainerUtil.py:408-460, renders them one after another).  Results are returned in submission order by `drain()` / the per-slot
    events; the per-frame arithmetic is exactly FrameRenderer's."""

    def __init__(self, model: HeadModel, n_rays, depth=2, **kw):
        self.depth = int(depth)
        self.smooth = bool(kw.get("smooth_lips", False))
        if self.smooth and kw.get("lips_state") is None:                 # ONE smoothing state for the whole frame sequence
            kw["lips_state"] = torch.zeros(33, device=next(model.parameters()).device)
        if self.depth >= 2 and "head_ctas" not in kw:
            # several frames share the GPU: each frame's network launches take 40 % of the SMs, so two of them run side by side and the small march /
            # composite launches of the other frames find free SMs (3397 -> 3550 frames/s at depth 5 on B200 when introduced; 50..100 of 148 SMs are within 2 %)
            kw["head_ctas"] = max(1, int(round(0.4 * torch.cuda.get_device_properties(next(model.parameters()).device).multi_processor_count)))
        self.slots = [FrameRenderer(model, n_rays, **kw) for _ in range(self.depth)]
        self.dev = self.slots[0].dev
        # smooth_lips makes the audio code of frame k depend on frame k-1: the (tiny) audio kernels run in frame order on their own stream, the frames
        # themselves still overlap on the slot streams
        self.audio_stream = torch.cuda.Stream(device=self.dev) if self.smooth else None
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(self.depth)]
        self.done = [torch.cuda.Event() for _ in range(self.depth)]
        self.k = 0

    def _slot(self):
        i = self.k % self.depth
        self.k += 1
        self.streams[i].wait_stream(torch.cuda.current_stream(self.dev))      # inputs produced on the caller's stream are ready
        return i

    def _ordered_audio(self, i, auds):
        """smooth_lips: copy the window and run audio encode + EMA for slot i on the audio stream (frame order), then let the slot's stream wait for it."""
        sl, a = self.slots[i], self.audio_stream
        a.wait_stream(torch.cuda.current_stream(self.dev))
        a.wait_event(self.done[i])                                         # the slot's previous frame no longer reads its auds / enc_a buffers
        with torch.cuda.stream(a):
            sl.auds.copy_(auds, non_blocking=True)
            sl.encode_audio_smoothed()
        self.streams[i].wait_stream(a)
        sl._audio_done = True

    def submit_device(self, rays_o, rays_d, auds):
        """Enqueue one frame whose inputs are on the device; returns (slot index, the slot's static image buffer)."""
        i = self._slot()
        if self.smooth:
            self._ordered_audio(i, auds); auds = None
        with torch.cuda.stream(self.streams[i]):
            img = self.slots[i].render_device(rays_o, rays_d, auds)
            self.done[i].record()
        self.slots[i]._audio_done = False
        return i, img

    def submit_host(self, rays_o_host, rays_d_host, auds_host, out_host):
        """Enqueue one frame with pinned host inputs / output (out_host must stay untouched until the slot's event or drain())."""
        i = self._slot()
        if self.smooth:
            self._ordered_audio(i, auds_host); auds_host = None
        with torch.cuda.stream(self.streams[i]):
            self.slots[i].render_host(rays_o_host, rays_d_host, auds_host, out_host)
            self.done[i].record()
        self.slots[i]._audio_done = False
        return i

    def submit_host_pose(self, pose_host, auds_host, out_u8_host):
        """Enqueue one frame from a pinned host pose + audio window; the uint8 frame lands in out_u8_host (renderers built with camera=...)."""
        i = self._slot()
        if self.smooth:
            self._ordered_audio(i, auds_host); auds_host = None
        with torch.cuda.stream(self.streams[i]):
            self.slots[i].render_host_pose(pose_host, auds_host, out_u8_host)
            self.done[i].record()
        self.slots[i]._audio_done = False
        return i

    def drain(self):
        """Make the caller's stream wait for every frame in flight."""
        cur = torch.cuda.current_stream(self.dev)
        for s in self.streams:
            cur.wait_stream(s)

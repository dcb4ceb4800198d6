#!/usr/bin/env python
"""Training composites alone (4 M rays, 0..16 samples each) — for sweeping the staging capacity:  B2N_COMP_CAP=1024 python profiles/bench_composite.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from raymarching.backend import _backend as rb
from profiles.kernel_rooflines import timeit, HBM

dev = torch.device("cuda")
torch.manual_seed(0)
N = 4 * 1024 * 1024
hi = int(os.environ.get("COUNT_MAX", "16"))
counts = torch.randint(0, hi + 1, (N,), device=dev, dtype=torch.int32)
offs = torch.cumsum(counts, 0, dtype=torch.int32) - counts
rays = torch.stack([torch.arange(N, device=dev, dtype=torch.int32), offs, counts], 1).contiguous()
M = int(counts.sum())
sig = torch.rand(M, device=dev) * 5; rgb = torch.rand(M, 3, device=dev); aud = torch.rand(M, device=dev); eye = torch.rand(M, device=dev)
unc = torch.rand(M, device=dev); dl = torch.rand(M, 2, device=dev) * 0.03 + 0.01
ws, a0, a1, us, dep = (torch.empty(N, device=dev) for _ in range(5)); img = torch.empty(N, 3, device=dev)
f = lambda: rb.composite_rays_train_triplane_forward(sig, rgb, aud, eye, unc, dl, rays, M, N, 1e-4, ws, a0, a1, us, dep, img)
t = timeit(f); print(f"cap={os.environ.get('B2N_COMP_CAP', 'default')} fwd {t * 1e6:.1f} us  {(36 * M + 44 * N) / t / 1e9 / HBM:.3f} of HBM")
g_ws, g_a0, g_a1, g_u = (torch.randn(N, device=dev) for _ in range(4)); g_img = torch.randn(N, 3, device=dev)
gs, ga0, ga1, gu = (torch.zeros(M, device=dev) for _ in range(4)); grgb = torch.zeros(M, 3, device=dev)
f = lambda: rb.composite_rays_train_triplane_backward(g_ws, g_a0, g_a1, g_u, g_img, sig, rgb, aud, eye, unc, dl, rays, ws, a0, a1, us, img, M, N, 1e-4, gs, grgb, ga0, ga1, gu)
t = timeit(f); print(f"  bwd {t * 1e6:.1f} us  {((36 + 28) * M + 72 * N) / t / 1e9 / HBM:.3f} of HBM")
win = (torch.arange(N, device=dev) // 4096) * 4096
local = (win.float() + torch.rand(N, device=dev) * 4096).argsort()
rays_loc = rays[local].contiguous()
f = lambda: rb.composite_rays_train_triplane_forward(sig, rgb, aud, eye, unc, dl, rays_loc, M, N, 1e-4, ws, a0, a1, us, dep, img)
t = timeit(f, reps=3); print(f"  fwd scrambled(4096) {t * 1e6:.1f} us  {(36 * M + 44 * N) / t / 1e9 / HBM:.3f} of HBM")
f = lambda: rb.composite_rays_train_triplane_backward(g_ws, g_a0, g_a1, g_u, g_img, sig, rgb, aud, eye, unc, dl, rays_loc, ws, a0, a1, us, img, M, N, 1e-4, gs, grgb, ga0, ga1, gu)
t = timeit(f, reps=3); print(f"  bwd scrambled(4096) {t * 1e6:.1f} us  {((36 + 28) * M + 72 * N) / t / 1e9 / HBM:.3f} of HBM")

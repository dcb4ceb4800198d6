#!/usr/bin/env python
"""march_rays_train at 1 M random training rays of the synthetic head scene, a few launches (run under `ncu --metrics gpu__time_duration.sum` for the per-kernel split)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
import raymarching
from raymarching.backend import _backend as rb
from b2nerf import scene
dev = torch.device("cuda")
bf = torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).to(dev)
Nm = 1024 * 1024
oo, dd = [], []
for s in range(16):
    a, b = scene.train_rays(s, 65536)
    oo.append(torch.from_numpy(a)); dd.append(torch.from_numpy(b))
ro, rd = torch.cat(oo).to(dev), torch.cat(dd).to(dev)
aabb = torch.from_numpy(scene.AABB).to(dev)
nears, fars = raymarching.near_far_from_aabb(ro, rd, aabb, 0.05)
Mm = Nm * 6
xyzs, dirs, dls = torch.zeros(Mm, 3, device=dev), torch.zeros(Mm, 3, device=dev), torch.zeros(Mm, 2, device=dev)
rr = torch.empty(Nm, 3, device=dev, dtype=torch.int32); counter = torch.zeros(2, device=dev, dtype=torch.int32); noises = torch.rand(Nm, device=dev)
for n in (Nm, 65536):
    for _ in range(3):
        counter.zero_()
        rb.march_rays_train(ro[:n], rd[:n], bf, 1.0, 1 / 256, 16, n, 1, 128, Mm, nears[:n], fars[:n], xyzs, dirs, dls, rr[:n], counter, noises[:n])
    torch.cuda.synchronize()
    print(n, "rays ->", int(counter[0]), "samples")

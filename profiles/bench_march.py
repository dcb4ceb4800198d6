#!/usr/bin/env python
"""march_rays_train alone at 1 M and 65 536 rays of the synthetic head scene (count + emit), CUDA events.  python profiles/bench_march.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
import raymarching
from raymarching.backend import _backend as rb
from b2nerf import scene
from profiles.kernel_rooflines import timeit, HBM

dev = torch.device("cuda")
bf = torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).to(dev)
aabb = torch.from_numpy(scene.AABB).to(dev)
oo, dd = [], []
for s in range(16):
    a, b = scene.train_rays(s, 65536)
    oo.append(torch.from_numpy(a)); dd.append(torch.from_numpy(b))
for Nm in (1024 * 1024, 65536):
    ro, rd = torch.cat(oo)[:Nm].to(dev), torch.cat(dd)[:Nm].to(dev)
    nears, fars = raymarching.near_far_from_aabb(ro, rd, aabb, 0.05)
    Mm = Nm * 6
    xyzs, dirs, dls = torch.zeros(Mm, 3, device=dev), torch.zeros(Mm, 3, device=dev), torch.zeros(Mm, 2, device=dev)
    rr = torch.empty(Nm, 3, device=dev, dtype=torch.int32); counter = torch.zeros(2, device=dev, dtype=torch.int32); noises = torch.rand(Nm, device=dev)

    def march():
        counter.zero_()
        rb.march_rays_train(ro, rd, bf, 1.0, 1 / 256, 16, Nm, 1, 128, Mm, nears, fars, xyzs, dirs, dls, rr, counter, noises)

    sec = timeit(march)
    tot = int(counter[0])
    print(f"N={Nm}: {sec * 1e6:.1f} us, {tot} samples, {(52 * Nm + 32 * tot) / sec / 1e9 / HBM:.3f} of HBM, {Nm / sec / 1e6:.0f} M rays/s")

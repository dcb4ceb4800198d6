#!/usr/bin/env python
"""morton3D_dilation alone on 8 x 256^3 cells.  python profiles/bench_dilation.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from raymarching.backend import _backend as rb
from profiles.kernel_rooflines import timeit, HBM
dev = torch.device("cuda")
grid = torch.rand(8, 256 ** 3, device=dev) * 20
gd = torch.empty_like(grid)
t = timeit(lambda: rb.morton3D_dilation(grid, 8, 256, gd))
print(f"morton3D_dilation {t * 1e6:.1f} us  {8 * grid.numel() / t / 1e9 / HBM:.3f} of HBM")

#!/usr/bin/env python
"""grid_encode_backward (tri-plane config) at the training step's size, privatised kernel with different slice counts vs the direct scatter."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np, torch
from gridencoder import GridEncoder
from gridencoder.backend import _backend as gb
enc = GridEncoder(input_dim=2, num_levels=12, level_dim=1, base_resolution=64, log2_hashmap_size=14, desired_resolution=512).cuda()
B = 327040
g = torch.Generator(device="cuda").manual_seed(0)
x = (0.5 + 0.1 * torch.randn(B, 2, device="cuda", generator=g)).clamp(0, 1)       # head-sized cluster
grad = torch.randn(12, B, 1, device="cuda", generator=g)
ge = torch.zeros_like(enc.embeddings.data)
S = float(np.log2(enc.per_level_scale))
def t(reps=20):
    f = lambda: gb.grid_encode_backward(grad, x, enc.embeddings.data, enc.offsets, ge, B, 2, 1, 12, S, 64, None, None, 0, False)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
os.environ["B2N_GRID_BWD_DIRECT"] = "1"; print("direct", round(t(), 1), "us"); del os.environ["B2N_GRID_BWD_DIRECT"]
for s in (2, 4, 6, 8, 12, 16, 24, 36):
    os.environ["B2N_GRID_BWD_SLICES"] = str(s); print("slices", s, round(t(), 1), "us")

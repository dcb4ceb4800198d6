#!/usr/bin/env python
"""Which torch ops (with shapes) are left in the fused training step: eager step under torch.profiler(record_shapes=True), grouped by op + shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from b2nerf import scene
from b2nerf.train import Trainer
dev = torch.device("cuda")
model = bench.build_model(dev); model.testing = False
model.density_bitfield.copy_(torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).to(dev))
tr = Trainer(model, fp16=True, fused_head=True)
n = 65536
batches = []
for s in range(4):
    o, d = scene.train_rays(step=s, n=n)
    batches.append((torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev), torch.from_numpy(scene.audio_window(s)).to(dev), torch.rand(n, 3, device=dev)))
for s in range(20):
    b = batches[s % 4]; tr.train_step(*b, index=s)
    if s == 15: tr.update_mean_count()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    for s in range(3):
        b = batches[s % 4]; tr.train_step(*b, index=s)
    torch.cuda.synchronize()
ka = prof.key_averages(group_by_input_shape=True)
rows = [(k.key, str(k.input_shapes)[:90], k.self_device_time_total / 3, k.count // 3) for k in ka if k.self_device_time_total > 0 and k.key.startswith("aten::")]
rows.sort(key=lambda r: -r[2])
for name, shp, us, cnt in rows[:45]:
    print(f"{us:8.1f} us x{cnt:<3d} {name:34s} {shp}")

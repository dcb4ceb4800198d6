#!/usr/bin/env python
"""Kernel-time breakdown of one data-parallel training step (BASELINE configs[2]: 65 536 rays) with torch.profiler.
Run on a B200:  python profiles/profile_train_step.py > gpurun_out/train_profile.txt"""
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
tr = Trainer(model, fp16=True, fused_head="--unfused" not in sys.argv)
n = 65536
batches = []
for s in range(4):
    o, d = scene.train_rays(step=s, n=n)
    batches.append((torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev), torch.from_numpy(scene.audio_window(s)).to(dev), torch.rand(n, 3, device=dev)))
for s in range(20):
    b = batches[s % 4]; tr.train_step(*b, index=s)
    if s == 15: tr.update_mean_count()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
step = tr.train_step if "--eager" in sys.argv else tr.train_step_graphed
for s in range(4):
    b = batches[s % 4]; step(*b, index=s)
torch.cuda.synchronize()
e0.record()
for s in range(10):
    b = batches[s % 4]; step(*b, index=s)
e1.record(); torch.cuda.synchronize()
print("ms/step (events, 10 steps):", e0.elapsed_time(e1) / 10)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for s in range(5):
        b = batches[s % 4]; step(*b, index=s)
    torch.cuda.synchronize()
ka = prof.key_averages()
tot = sum(k.device_time_total for k in ka if k.device_type == torch.autograd.DeviceType.CUDA) if False else None
rows = [(k.key, k.self_device_time_total / 5, k.count // 5) for k in ka if k.self_device_time_total > 0]
rows.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"device time per step: {tot:.1f} us in {sum(r[2] for r in rows)} kernels")
for name, us, cnt in rows[:60]:
    print(f"{us:9.1f} us {100 * us / tot:5.1f}%  x{cnt:<4d} {name[:150]}")

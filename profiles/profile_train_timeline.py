#!/usr/bin/env python
"""Kernel TIMELINE of one graphed training step (BASELINE configs[2]: 65 536 rays): start offset, duration and stream of every kernel of one replay, so the
overlap between the branches of the step's graph (weight gradients | table gradients | audio backward; marcher | operand packing | audio encode) can be read.
Run on a B200:  python profiles/profile_train_timeline.py > gpurun_out/train_timeline.txt"""
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
o, d = scene.train_rays(step=0, n=n)
b = (torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev), torch.from_numpy(scene.audio_window(0)).to(dev), torch.rand(n, 3, device=dev))
for s in range(20):
    tr.train_step(*b, index=s)
    if s == 15: tr.update_mean_count()
torch.cuda.synchronize()
for s in range(40):
    tr.train_step_graphed(*b, index=s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for s in range(32):
    tr.train_step_graphed(*b, index=s)
e1.record(); torch.cuda.synchronize()
print("ms/step (events, 32 graphed steps):", e0.elapsed_time(e1) / 32)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for s in range(3):
        tr.train_step_graphed(*b, index=s)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# the last replay: kernels after the last big gap
starts = [e.time_range.start for e in ev]
cut = 0
for i in range(1, len(ev)):
    if "k_near_far" in ev[i].name:
        cut = i
t0 = ev[cut].time_range.start
print(f"{'start us':>9} {'dur us':>8}  stream  kernel")
for e in ev[cut:]:
    print(f"{e.time_range.start - t0:9.1f} {e.time_range.end - e.time_range.start:8.1f}  {getattr(e, 'device_index', 0)}:{getattr(e, 'device_resource_id', '?')}  {e.name[:110]}")

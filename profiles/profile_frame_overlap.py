#!/usr/bin/env python
"""Timeline of the frame pipeline (6 frames in flight, WHILE-graph path) with torch.profiler / CUPTI: per-kernel busy time vs wall time.
    python profiles/profile_frame_overlap.py > gpurun_out/frame_overlap.txt"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from b2nerf import scene
from b2nerf.render import FramePipeline

dev = torch.device("cuda")
model = bench.build_model(dev); model.testing = True
model.density_bitfield.copy_(torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).to(dev))
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 6
pipe = FramePipeline(model, bench.N_RAYS, depth=depth)
frames = [tuple(torch.from_numpy(a).to(dev) for a in scene.frame_rays(frame=f)) for f in range(8)]
auds = [torch.from_numpy(scene.audio_window(frame=f)).to(dev) for f in range(8)]
def run(n):
    for k in range(n):
        pipe.submit_device(frames[k % 8][0], frames[k % 8][1], auds[k % 8])
    pipe.drain(); torch.cuda.synchronize()
run(24)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(48)
ev = [e for e in prof.events() if e.device_type is not None and "cuda" in str(e.device_type).lower() and e.time_range is not None]
ks = [(e.name.split("(")[0].replace("b2n::", "").replace("void ", ""), e.time_range.start, e.time_range.end) for e in ev]
if not ks:
    print("no kernel records (CUPTI does not see kernels inside the conditional graph)"); sys.exit(0)
t0, t1 = min(k[1] for k in ks), max(k[2] for k in ks)
wall = t1 - t0
busy = collections.Counter(); cnt = collections.Counter()
for n, a, b in ks:
    busy[n] += b - a; cnt[n] += 1
print(f"wall {wall:.1f} us for 48 frames = {wall / 48:.1f} us per frame; kernel records {len(ks)}")
for n, v in busy.most_common(12):
    print(f"  {n:40s} x{cnt[n]:5d}  sum {v:9.1f} us  = {v / 48:7.1f} us per frame  ({100 * v / wall:5.1f} % of wall)")
# union of head-kernel intervals: fraction of wall time with at least one head kernel running
iv = sorted((a, b) for n, a, b in ks if n.startswith("k_head_"))
u, cur_a, cur_b = 0.0, None, None
for a, b in iv:
    if cur_b is None or a > cur_b:
        if cur_b is not None: u += cur_b - cur_a
        cur_a, cur_b = a, b
    else:
        cur_b = max(cur_b, b)
if cur_b is not None: u += cur_b - cur_a
print(f"time with >= 1 head kernel running: {u:.1f} us = {100 * u / wall:.1f} % of wall")


def coverage(pred):
    """time with k = 0, 1, 2, ... kernels matching pred running"""
    pts = []
    for n, a, b in ks:
        if pred(n):
            pts.append((a, 1)); pts.append((b, -1))
    pts.sort()
    hist, level, last = collections.Counter(), 0, t0
    for t, d in pts:
        hist[level] += t - last
        level += d; last = t
    hist[level] += t1 - last
    return hist

for label, pred in (("head kernels", lambda n: n.startswith("k_head_")), ("any kernel", lambda n: True),
                    ("non-head kernels", lambda n: not n.startswith("k_head_"))):
    h = coverage(pred)
    print(label + " concurrently running: " + ", ".join(f"{k}: {100 * v / wall:.1f} %" for k, v in sorted(h.items())))

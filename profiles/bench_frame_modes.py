#!/usr/bin/env python
"""Frame pipeline throughput, WHILE-node graph vs fixed 16-iteration graph (B2N_FRAME_NO_WHILE=1), same process, CUDA events, inputs cycling through 24 frames.
    python profiles/bench_frame_modes.py [depth]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
import bench
from b2nerf import scene
from b2nerf.render import FramePipeline

dev = torch.device("cuda")
model = bench.build_model(dev); model.testing = True
model.density_bitfield.copy_(torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).to(dev))
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 5
frames = [tuple(torch.from_numpy(a).to(dev) for a in scene.frame_rays(frame=f)) for f in range(24)]
auds = [torch.from_numpy(scene.audio_window(frame=f)).to(dev) for f in range(24)]
for mode in ("0", "1", "0", "1"):
    os.environ["B2N_FRAME_NO_WHILE"] = mode
    t_build = time.time()
    pipe = FramePipeline(model, bench.N_RAYS, depth=depth)
    t_build = time.time() - t_build
    def run(n):
        for k in range(n):
            pipe.submit_device(frames[k % 24][0], frames[k % 24][1], auds[k % 24])
        pipe.drain()
    run(48); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(480); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"no_while={mode} depth={depth}: {480 / ms * 1e3:.0f} frames/s ({ms / 480 * 1e3:.1f} us per frame), build {t_build:.1f} s", flush=True)
    del pipe

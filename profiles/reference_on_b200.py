#!/usr/bin/env python
"""The "reference on B200" row of SURVEY §8(d): the reference's OWN CUDA extensions (oracle/_ref/*.so, compiled in place from /root/reference by
oracle/build_ref_ext.sh) driven the way the reference drives them — host loop with one alive-count read-back per iteration
(renderer.py:406-570), per-op kernels, torch MLPs under autocast, plain AdamW + GradScaler — on the same synthetic frames / batches as bench.py.

Test infrastructure (it executes oracle/_ref): the drop-in packages' `backend` modules are replaced by the reference's pybind modules BEFORE the
packages are imported, so the Python above the backend is identical on both sides and only the kernels differ.

    python profiles/reference_on_b200.py > gpurun_out/reference_on_b200.json
"""
import glob, importlib.util, json, os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch


def load_ref(name):
    hits = glob.glob(os.path.join(ROOT, "oracle", "_ref", name + ".*.so"))
    if not hits:
        print(json.dumps({"unavailable": f"oracle/_ref/{name} not built"})); sys.exit(0)
    spec = importlib.util.spec_from_file_location(name, hits[0])
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod


for pkg, ext in (("raymarching", "_ref_raymarching_face"), ("gridencoder", "_ref_grid_encoder"), ("shencoder", "_ref_sh_encoder"), ("freqencoder", "_ref_freqencoder")):
    stub = types.ModuleType(pkg + ".backend")
    stub._backend = load_ref(ext)
    sys.modules[pkg + ".backend"] = stub

import raymarching  # noqa: E402  (now on the reference kernels)
import bench  # noqa: E402
from b2nerf import scene  # noqa: E402
from b2nerf.model import MLP  # noqa: E402
from b2nerf.train import Trainer  # noqa: E402

MLP.tall_linear = False
dev = torch.device("cuda")
N, POOL = bench.N_RAYS, 8
model = bench.build_model(dev)
bitfield = torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).to(dev)
model.density_bitfield.copy_(bitfield)
frames = [tuple(torch.from_numpy(a).to(dev) for a in scene.frame_rays(frame=f)) for f in range(POOL)]
auds = [torch.from_numpy(scene.audio_window(frame=f)).to(dev) for f in range(POOL)]
eye = torch.tensor([[0.4]], device=dev)
ind_code = model.individual_codes[0:1].detach()


@torch.no_grad()
def render_frame(rays_o, rays_d, aud, max_steps=16, dt_gamma=1 / 256, T_thresh=1e-4):
    """run_cuda_for_inference (renderer.py:406-570): per-op kernels, unfused network under autocast, torch mask compaction."""
    with torch.autocast("cuda", dtype=torch.float16):
        enc_a = model.encode_audio(aud)
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, model.aabb_infer, 0.05)
        ws, depth, image = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(N, 3, device=dev)
        sa, se, su = torch.zeros(N, device=dev), torch.zeros(N, device=dev), torch.zeros(N, device=dev)
        alive = torch.arange(N, dtype=torch.int32, device=dev); rays_t = nears.clone()
        step, samples = 0, 0
        while step < max_steps:
            n_alive = alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, model.bound, model.density_bitfield, model.cascade,
                                                        model.grid_size, nears, fars, 128, False, dt_gamma, max_steps)
            sig, rgb, aa, ae, un = model.forward_unfused(xyzs, dirs, enc_a, ind_code, eye)
            raymarching.composite_rays_triplane(n_alive, n_step, alive, rays_t, sig, rgb, deltas, aa, ae, un, ws, depth, image, sa, se, su, T_thresh)
            alive = alive[alive >= 0]            # the reference's per-iteration host synchronisation (renderer.py:542)
            samples += n_alive * n_step
            step += n_step
        image = (image + (1 - ws).unsqueeze(-1)).clamp(0, 1)
    return image, samples


def timed(fn, steps, warmup):
    for s in range(warmup):
        fn(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        fn(warmup + s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


out = {"what": "reference CUDA extensions (oracle/_ref) + torch MLPs, reference-style host orchestration, same synthetic inputs as bench.py", "gpu": torch.cuda.get_device_name(0)}
smp = []
ms = timed(lambda s: smp.append(render_frame(frames[s % POOL][0], frames[s % POOL][1], auds[s % POOL])[1]), 20, 4)
out["infer_512x512"] = {"frames_per_sec": 1e3 / ms, "ms_per_frame": ms, "sample_slots_per_frame": smp[-1]}

# training step: eager, per-op reference kernels, nn.Linear MLPs, foreach AdamW, GradScaler (TrainerUtil.py:1040-1056)
model_t = bench.build_model(dev); model_t.testing = False
model_t.density_bitfield.copy_(bitfield)
tr = Trainer(model_t, fp16=True, fused_optimizer=False)
n = 65536
batches = []
for s in range(4):
    o, d = scene.train_rays(step=s, n=n)
    batches.append((torch.from_numpy(o).to(dev), torch.from_numpy(d).to(dev), torch.from_numpy(scene.audio_window(s)).to(dev), torch.rand(n, 3, device=dev)))
for s in range(18):
    b = batches[s % 4]; tr.train_step(*b, index=s)
    if s == 15:
        tr.update_mean_count()
ms = timed(lambda s: tr.train_step(*batches[s % 4], index=s), 20, 2)
out["train_65536_rays"] = {"rays_per_sec": n / (ms * 1e-3), "ms_per_step": ms, "mean_count": tr.mean_count}
print(json.dumps(out, indent=1))

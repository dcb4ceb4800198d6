#!/usr/bin/env python
"""The "reference on B200" row of SURVEY §8(d): the UNMODIFIED reference model code — nerf_triplane/network.py NeRFNetwork + renderer.py
(run_cuda_for_inference / run_cuda), its op wrappers and its own CUDA extensions — staged by oracle/stage_ref_py.sh + oracle/build_ref_ext.sh (only
-std=c++17 differs from its own build flags), driven the way the reference drives it:
  * inference: TrainerUtil.test_step (TrainerUtil.py:408-460): torch.no_grad + autocast(fp16), model.render(...) per frame, host loop with one alive-count
    read-back per iteration (renderer.py:503-545);
  * training:  TrainerUtil.train_one_epoch (TrainerUtil.py:1031-1056): autocast, model.render(training), loss, scaler.scale(loss).backward(), scaler.step(AdamW over
    model.get_params), scaler.update(), loss.item() — the head-branch loss of TrainerUtil.py:238-334 written inline (TrainerUtil itself imports the product's
    logging / LPIPS / video stack and is not staged).
Same synthetic scene, frames, batches and random-init weights as bench.py.  TEST / MEASUREMENT INFRASTRUCTURE: nothing of lzzx-nerf_b200/ is on the timed path
(only b2nerf.scene, numpy input synthesis).

    python profiles/reference_on_b200.py [--frames 20] [--steps 20]     -> one JSON object on stdout
"""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_PY = os.path.join(ROOT, "oracle", "_ref_py")
if not os.path.isdir(os.path.join(REF_PY, "nerf_triplane")):
    print(json.dumps({"unavailable": "oracle/_ref_py not staged (oracle/stage_ref_py.sh)"})); sys.exit(0)
sys.path[:0] = [REF_PY, os.path.join(ROOT, "tests"), os.path.join(ROOT, "lzzx-nerf_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

try:
    for _ in range(32):
        try:
            from nerf_triplane.network import NeRFNetwork  # noqa: E402
            import nerf_triplane.renderer as ref_renderer  # noqa: E402
            break
        except ModuleNotFoundError as e:
            sys.modules[e.name] = types.ModuleType(e.name)
            for k in [k for k in sys.modules if k.startswith("nerf_triplane")]:
                del sys.modules[k]
    import raymarching  # noqa: E402
    assert raymarching.__file__.startswith(REF_PY)
except ImportError as e:
    print(json.dumps({"unavailable": f"reference extensions not built: {e}"})); sys.exit(0)
import refcases as rc  # noqa: E402
from b2nerf import scene  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=20)
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
torch.backends.cuda.matmul.allow_tf32 = False      # train.py:11-13
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda")
HW, N, POOL = 512, 512 * 512, 8
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def build(training):
    torch.manual_seed(0)
    m = NeRFNetwork(rc.ref_opt(False, "hubert")).to(dev)       # the reference's own random init (tables +-1e-4, grid.py:132-134)
    m.density_bitfield.copy_(T(scene.bitfield_from_grid(scene.density_grid())))
    m.testing = not training
    return m.train() if training else m.eval()


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


out = {"what": "unmodified reference model code (oracle/_ref_py) on the reference's own CUDA extensions (oracle/_ref), reference-style host loop; same synthetic "
               "inputs as bench.py", "gpu": torch.cuda.get_device_name(0)}
# ---- inference: 512 x 512 frames ----------------------------------------------------------------------------------------------------------
m = build(False)
frames = [tuple(T(a) for a in scene.frame_rays(frame=f)) for f in range(POOL)]
auds = [T(scene.audio_window(frame=f)) for f in range(POOL)]
eye = torch.tensor([[0.4]], device=dev)
bgc = torch.zeros(1, N, 2, device=dev)


def frame(s):
    ro, rd = frames[s % POOL]
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        res, _ = m.render(ro[None], rd[None], auds[s % POOL], bgc, None, eye=eye, index=[0], bg_color=None, perturb=False, dt_gamma=1 / 256, max_steps=16, T_thresh=1e-4)
    return res["image"]


ms = timed(frame, args.frames, 4)
out["infer_512x512"] = {"frames_per_sec": 1e3 / ms, "ms_per_frame": ms, "frames_timed": args.frames}

# ---- training: 65 536 rays per step -------------------------------------------------------------------------------------------------------------
mt = build(True)
opt = torch.optim.AdamW(mt.get_params(1e-2, 1e-3), betas=(0.0, 0.99), eps=1e-8)        # train.py:274 (its integer 0 is rejected by torch 2.11's AdamW: the only edit)
scaler = torch.amp.GradScaler("cuda")
n = 65536
batches = []
for s in range(4):
    o, d = scene.train_rays(step=s, n=n)
    batches.append((T(o), T(d), T(scene.audio_window(s)), torch.rand(n, 3, device=dev), torch.rand(n, device=dev) < 0.5))
bgc_t, bg1 = torch.zeros(1, n, 2, device=dev), torch.ones(1, 3, device=dev)
state = {"global_step": 0}


def step(s):
    ro, rd, au, gt, face = batches[s % 4]
    opt.zero_grad()
    sf = min(state["global_step"] / 200000, 1.0)
    with torch.autocast("cuda", dtype=torch.float16):
        res, _ = mt.render(ro[None], rd[None], au, bgc_t, None, eye=eye, index=[s % 7], bg_color=bg1, perturb=True, force_all_rays=False, dt_gamma=1 / 256, max_steps=16,
                           T_thresh=1e-4)
        pred = res["image"].view(-1, 3)
        loss = ((pred - gt) ** 2).mean(-1)
        unc = res["uncertainty"]
        w = torch.softmax(unc, dim=-1) * n
        loss = loss * (0.2 + 0.8 * ((1 - sf) + sf * w.detach()).clamp(0, 10))
        beta = unc + 1
        loss = loss + sf * ((torch.norm(pred - gt, dim=-1).detach() / (2 * beta ** 2) + torch.log(beta) ** 2 / 2) * face) + 1e-3 * sf * (unc * ~face)
        loss = loss.mean()
        al = res["weights_sum"].clamp(1e-5, 1 - 1e-5)
        loss = loss + 1e-4 * (-al * torch.log2(al) - (1 - al) * torch.log2(1 - al)).mean()
        lam = sf * 1e-4
        loss = loss + lam * (res["ambient_aud"].view(-1) * ~face).mean() + lam * ((res["ambient_eye"].view(-1) / 16 * res["ambient_aud"].view(-1).detach()) * face).mean()
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    state["global_step"] += 1
    return loss.item()                                  # TrainerUtil.py:1051


for s in range(16):                                     # warm-up with worst-case buffers, then the mean_count estimate (update_extra_state's tail, renderer.py:812-815)
    step(s)
mt.mean_count = int(mt.step_counter[:16, 0].sum().item() / 16)
mt.local_step = 0
ms = timed(step, args.steps, 2)
out["train_65536_rays"] = {"rays_per_sec": n / (ms * 1e-3), "ms_per_step": ms, "mean_count": mt.mean_count, "steps_timed": args.steps,
                           "note": "without the every-16th-step smoothness regulariser (TrainerUtil.py:336-363)"}
print(json.dumps(out))

#!/usr/bin/env python
"""Debug helper: where do the GPU and CPU-port table gradients differ?  (run on a B200: python profiles/debug_train_grads.py)"""
import os, sys, copy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import oracle
from oracle import torch_port as tp
from b2nerf import scene
from test_gpu_train import _setup

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
n = 2048
m, Trainer, bf, o, d, auds, gt = _setup(n)
tr = Trainer(m, fp16=False)
rays_o, rays_d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
eye = torch.full((1, 1), 0.4, device="cuda"); bg = torch.ones(1, 3, device="cuda")
tr.grads.zero_()
out = tr.render_train(rays_o, rays_d, auds.cuda(), 3, eye, bg, perturb=False)
loss = tr.loss(out, gt.cuda()); loss.backward()
mc = copy.deepcopy(m).cpu()
nears, fars = oracle.near_far_from_aabb(o, d, scene.AABB, 0.05)
xyzs, dirs, deltas, rays, cnt = oracle.march_rays_train(o, d, bf, 1.0, 1 / 256, 16, 1, 128, n * 16, nears, fars, np.zeros(n, np.float32))
tot = int(cnt[0])
P = {k: v for k, v in mc.named_parameters()}
p = dict(P); p["encoder_xy.offsets"] = mc.encoder_xy.offsets; p["S"] = float(np.log2(mc.encoder_xy.per_level_scale)); p["H"] = 64; p["bound"] = 1.0
enc_a = mc.encode_audio(auds)[0]
sig, rgb, aa, ae, un = tp.head_forward(p, torch.from_numpy(xyzs[:tot]), torch.from_numpy(dirs[:tot]), enc_a, P["individual_codes"][3], torch.tensor([0.4]), testing=False)
sig.retain_grad(); rgb.retain_grad()
ws, s_a, s_e, s_u, dep, img = tp.composite_rays_train_triplane_ragged(sig, rgb, aa.abs(), ae.abs(), un, torch.from_numpy(deltas[:tot]), torch.from_numpy(rays.astype(np.int64)))
img = (img + (1 - ws)[:, None]).clamp(0, 1)
mse = ((img - gt) ** 2).mean(-1).mean()
al = ws.clamp(1e-5, 1 - 1e-5)
ent = (-al * torch.log2(al) - (1 - al) * torch.log2(1 - al)).mean()
cpu_loss = mse + 1e-3 * ent + 1e-4 * (s_a.mean() + s_e.mean())
cpu_loss.backward()
print("loss", float(loss), float(cpu_loss), "samples", tot)
print("image diff", float((out["image"].detach().cpu() - img.detach()).abs().max()), "ws diff", float((out["weights_sum"].detach().cpu() - ws.detach()).abs().max()))
offs = mc.encoder_xy.offsets.numpy()
for name in ("encoder_xy.embeddings", "encoder_yz.embeddings", "encoder_xz.embeddings", "sigma_net.net.0.weight", "color_net.net.1.weight"):
    g_gpu = dict(m.named_parameters())[name].grad.detach().cpu().flatten(); g_cpu = P[name].grad.flatten()
    err = (g_gpu - g_cpu).abs(); mx = float(g_cpu.abs().max())
    print(name, "max|g|", mx, "max err", float(err.max()), "rel", float(err.max()) / mx, "n(err>1e-3 max)", int((err > 1e-3 * mx).sum()), "of", err.numel(),
          "sum|err|/sum|g|", float(err.sum() / g_cpu.abs().sum()))
    if "embeddings" in name:
        top = torch.topk(err, 8).indices
        for i in top.tolist():
            lvl = int(np.searchsorted(offs, i, side="right") - 1)
            print(f"   idx {i} level {lvl} gpu {float(g_gpu[i]):+.6e} cpu {float(g_cpu[i]):+.6e}")

"""torch_port — TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product package).

Vectorised pure-PyTorch port of the reference's hot-path ops, runnable on the host cores: "the repo's ops run as pure-PyTorch
on the box's host cores" (BASELINE.json north_star / configs[0]).  The reference itself has NO CPU implementation of this path
(every wrapper forces CUDA: raymarching.py:33-34, freq.py:22; renderer.py:952-954 raises without cuda_ray), so this port is what
bench.py times as `cpu_baseline` (kind "port") and as the `--impl reference` arm.  It follows the same kernels as oracle.c:
    grid encode    gridencoder.cu:54-72, 124-175          march      raymarching.cu:857-928
    composite      raymarching.cu:1916-1974, 2177-2248    SH deg 4   shencoder.cu:44-67       head MLPs  network.py:252-311
It is fp32 without fused multiply-adds, so it is a throughput baseline and a loose numerical cross-check — the bit-faithful
checker is oracle.c.
"""
import math

import numpy as np
import torch

PRIME_Y = 2654435761


# ---------------------------------------------------------------------------------------------------------------------------
# encoders
# ---------------------------------------------------------------------------------------------------------------------------
def grid_encode_2d(u, table, offsets, S, H):
    """u [B,2] in [0,1]; table [sO] fp32; offsets list[L+1] -> [B,L] (hash gridtype, align_corners=False, C=1)."""
    B = u.shape[0]
    L = len(offsets) - 1
    out = torch.empty(B, L, dtype=torch.float32)
    inside = ((u >= 0) & (u <= 1)).all(dim=1)
    for lvl in range(L):
        size = offsets[lvl + 1] - offsets[lvl]
        scale = float(np.float32(np.exp2(np.float32(lvl) * np.float32(S))) * np.float32(H) - np.float32(1.0))
        res = int(math.ceil(scale)) + 1
        pos = u * scale + 0.5
        pg = pos.floor()
        fr = pos - pg
        x, y = pg[:, 0].long(), pg[:, 1].long()
        tab = table[offsets[lvl]:offsets[lvl + 1]]
        acc = torch.zeros(B, dtype=torch.float32)
        hashed = (res + 1) * (res + 1) > size
        for cy in (0, 1):
            for cx in (0, 1):
                xx, yy = x + cx, y + cy
                if hashed:
                    idx = (xx ^ ((yy * PRIME_Y) & 0xFFFFFFFF)) % size
                else:
                    idx = (xx + yy * (res + 1)) % size
                w = (fr[:, 0] if cx else 1 - fr[:, 0]) * (fr[:, 1] if cy else 1 - fr[:, 1])
                acc = acc + w * tab[idx.clamp_(0, size - 1)]
        out[:, lvl] = torch.where(inside, acc, torch.zeros(()))
    return out


def triplane_encode(xyz, tables, offsets, S, H, bound=1.0):
    """network.py:208-223: planes xy, yz, xz, each (x + bound) / (2 bound) -> 12 features; returns [B,36]."""
    u = (xyz + bound) / (2 * bound)
    return torch.cat([grid_encode_2d(u[:, [0, 1]], tables[0], offsets, S, H), grid_encode_2d(u[:, [1, 2]], tables[1], offsets, S, H),
                      grid_encode_2d(u[:, [0, 2]], tables[2], offsets, S, H)], dim=1)


def sh4(d):
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz, x2, y2, z2 = x * y, x * z, y * z, x * x, y * y, z * z
    return torch.stack([torch.full_like(x, 0.28209479177387814), -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
                        1.0925484305920792 * xy, -1.0925484305920792 * yz, 0.94617469575755997 * z2 - 0.31539156525251999, -1.0925484305920792 * xz,
                        0.54627421529603959 * (x2 - y2), 0.59004358992664352 * y * (-3 * x2 + y2), 2.8906114426405538 * xy * z,
                        0.45704579946446572 * y * (1 - 5 * z2), 0.3731763325901154 * z * (5 * z2 - 3), 0.45704579946446572 * x * (1 - 5 * z2),
                        1.4453057213202769 * z * (x2 - y2), 0.59004358992664352 * x * (-x2 + 3 * y2)], dim=1)


def _mlp(x, ws):
    for i, w in enumerate(ws):
        x = x @ w.t()
        if i != len(ws) - 1:
            x = torch.relu(x)
    return x


def head_forward(p, x, d, enc_a, c, e, testing=True):
    """NeRFNetwork.forward (network.py:252-311), fp32.  p: dict of fp32 CPU tensors with the reference's state_dict names."""
    tables = [p["encoder_xy.embeddings"][:, 0], p["encoder_yz.embeddings"][:, 0], p["encoder_xz.embeddings"][:, 0]]
    offsets = p["encoder_xy.offsets"].tolist()
    enc_x = triplane_encode(x, tables, offsets, p["S"], p["H"], p.get("bound", 1.0))
    att = _mlp(enc_x, [p["aud_ch_att_net.net.0.weight"], p["aud_ch_att_net.net.1.weight"]])
    enc_w = enc_a.view(1, -1) * att
    eye_att = torch.sigmoid(_mlp(enc_x, [p["eye_att_net.net.0.weight"], p["eye_att_net.net.1.weight"]]))
    h = _mlp(torch.cat([enc_x, enc_w, e.view(1, 1) * eye_att], dim=1), [p[f"sigma_net.net.{i}.weight"] for i in range(3)])
    sigma, geo = torch.exp(h[:, 0]), h[:, 1:]
    hc = torch.cat([sh4(d), geo, c.view(1, -1).expand(x.shape[0], -1)], dim=1)
    color = torch.sigmoid(_mlp(hc, [p["color_net.net.0.weight"], p["color_net.net.1.weight"]])) * 1.002 - 0.001
    if testing:
        unc = torch.full((x.shape[0],), math.log(2.0))
    else:
        unc = torch.log(1 + torch.exp(_mlp(enc_x.detach(), [p["unc_net.net.0.weight"], p["unc_net.net.1.weight"]])))[:, 0]      # unc_net(unc_inp.detach()), network.py:247
    return sigma, color, att.norm(dim=1), eye_att[:, 0], unc


# ---------------------------------------------------------------------------------------------------------------------------
# torso branch: NeRFRenderer.run_torso (renderer.py:572-631) + NeRFNetwork.forward_torso (network.py:170-205), fp32
# ---------------------------------------------------------------------------------------------------------------------------
def freq_encode(x, degree):
    """freqencoder.cu:48-57: [x | sin(2^f x), cos(2^f x) interleaved per frequency]; output index c >= D: col = c // D - 1, d = c % D, f = col // 2."""
    outs = [x]
    for f in range(degree):
        outs += [torch.sin(x * 2.0 ** f), torch.cos(x * 2.0 ** f)]
    return torch.cat(outs, dim=1)


def tiled_grid_encode_2d(u, table, offsets, S, H):
    """gridencoder.cu:54-72,124-175 with gridtype = tiled, D = 2, C = 2, align_corners = False; fp32.  u [B,2] in [0,1]; table [sO,2] -> [B, 2 L]."""
    B, L = u.shape[0], len(offsets) - 1
    out = torch.zeros(B, L, 2, dtype=torch.float32)
    for lvl in range(L):
        size = offsets[lvl + 1] - offsets[lvl]
        scale = float(np.float32(np.exp2(np.float32(lvl) * np.float32(S))) * np.float32(H) - np.float32(1.0))
        stride = int(math.ceil(scale)) + 2
        pos = u * scale + 0.5
        pg = pos.floor()
        fr = pos - pg
        x, y = pg[:, 0].long(), pg[:, 1].long()
        tab = table[offsets[lvl]:offsets[lvl + 1]]
        for cy in (0, 1):
            for cx in (0, 1):
                idx = (x + cx) + ((y + cy) * stride if stride <= size else 0)
                w = (fr[:, 0] if cx else 1 - fr[:, 0]) * (fr[:, 1] if cy else 1 - fr[:, 1])
                out[:, lvl] += w[:, None] * tab[idx % size]
    return out.reshape(B, 2 * L)


def torso_forward(p, bg_coords, h_const, density_grid, grid_size, thresh, bg_color=None, shrink=0.8):
    """p: fp32 CPU tensors under the reference's names (torso_deform_net.net.*.weight, torso_encoder.embeddings / .offsets, torso_net.net.*.weight) + "S", "H";
    bg_coords [N,2]; h_const [50] = [anchor encoding | individual code].  Returns (bg [N,3], alpha [N], mask [N])."""
    N = bg_coords.shape[0]
    occ = torch.nn.functional.grid_sample(density_grid.view(1, 1, grid_size, grid_size), bg_coords.view(1, -1, 1, 2), align_corners=True).view(-1)
    mask = occ > thresh
    x = bg_coords * shrink
    enc = freq_encode(x, 8)
    h = torch.cat([enc, h_const.view(1, -1).expand(N, -1)], dim=1)
    dx = _mlp(h, [p[f"torso_deform_net.net.{i}.weight"] for i in range(3)])
    u = ((x + dx).clamp(-1, 1) + 1) / 2
    g = tiled_grid_encode_2d(u, p["torso_encoder.embeddings"], p["torso_encoder.offsets"].tolist(), p["S"], p["H"])
    o = _mlp(torch.cat([g, h], dim=1), [p[f"torso_net.net.{i}.weight"] for i in range(3)])
    alpha = torch.where(mask, torch.sigmoid(o[:, 0]) * 1.002 - 0.001, torch.zeros(()))
    color = torch.where(mask[:, None], torch.sigmoid(o[:, 1:]) * 1.002 - 0.001, torch.zeros(()))
    bg = torch.ones(N, 3) if bg_color is None else bg_color.expand(N, 3)
    return color * alpha[:, None] + bg * (1 - alpha[:, None]), alpha, mask


# ---------------------------------------------------------------------------------------------------------------------------
# marching / compositing
# ---------------------------------------------------------------------------------------------------------------------------
def near_far_from_aabb(o, d, aabb, min_near):
    rd = 1.0 / d
    t0, t1 = (aabb[:3] - o) * rd, (aabb[3:] - o) * rd
    lo, hi = torch.minimum(t0, t1), torch.maximum(t0, t1)
    near, far = lo.max(dim=1).values, hi.min(dim=1).values
    miss = near > far
    near = torch.where(miss, torch.full_like(near, 3.4028234663852886e38), near.clamp_min(min_near))
    far = torch.where(miss, torch.full_like(far, 3.4028234663852886e38), far)
    return near, far


def _spread3(v):
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v


def march_rays(alive, rays_t, rays_o, rays_d, fars, bitfield, n_step, bound=1.0, dt_gamma=1 / 256, max_steps=16, H=128):
    """kernel_march_rays for one cascade (raymarching.cu:828-929), all alive rays at once.  Returns xyzs [n,n_step,3], dirs, deltas [n,n_step,2]."""
    n = alive.shape[0]
    o, d, far = rays_o[alive], rays_d[alive], fars[alive]
    rd = 1.0 / d
    t = rays_t[alive].clone()
    dt_max = 2 * math.sqrt(3) / H
    dt_min = min(dt_max, 2 * math.sqrt(3) / max_steps)
    xyzs, deltas = torch.zeros(n, n_step, 3), torch.zeros(n, n_step, 2)
    step = torch.zeros(n, dtype=torch.long)
    act = torch.nonzero((t < far) & (step < n_step)).squeeze(1)
    sgn = torch.where(d >= 0, torch.full_like(d, 0.5), torch.full_like(d, -0.5))
    while act.numel() > 0:
        ta = t[act]
        p = (o[act] + ta[:, None] * d[act]).clamp(-bound, bound)
        dt = (ta * dt_gamma).clamp(dt_min, dt_max)
        v = (0.5 * (p / bound + 1) * H).clamp(0, H - 1).long()
        m = _spread3(v[:, 0]) | (_spread3(v[:, 1]) << 1) | (_spread3(v[:, 2]) << 2)
        occ = ((bitfield[m >> 3].long() >> (m & 7)) & 1).bool()
        io = act[occ]
        if io.numel():
            k = step[io]
            xyzs[io, k] = p[occ]
            t[io] = ta[occ] + dt[occ]
            deltas[io, k, 0] = dt[occ]
            deltas[io, k, 1] = t[io]
            step[io] = k + 1
        ie = act[~occ]
        if ie.numel():
            pe, te = p[~occ], ta[~occ]
            tx = (((v[~occ].float() + 0.5 + sgn[ie]) / H * 2 - 1) * bound - pe) * rd[ie]
            tt = te + tx.min(dim=1).values.clamp_min(0)
            cur = te
            todo = torch.ones_like(cur, dtype=torch.bool)
            while todo.any():                                   # do { t += dt } while (t < tt)
                cur = torch.where(todo, cur + (cur * dt_gamma).clamp(dt_min, dt_max), cur)
                todo = todo & (cur < tt)
            t[ie] = cur
        act = act[(t[act] < far[act]) & (step[act] < n_step)]
    dirs = torch.where((deltas[:, :, :1] != 0), d[:, None, :].expand(-1, n_step, -1), torch.zeros(()))
    return xyzs, dirs, deltas, t


def composite_rays_train_triplane(sigmas, rgbs, aud, eye, unc, deltas, n_rays, n_per_ray, T_thresh=1e-4):
    """Equal-length segments (BASELINE configs[0]: 4096 rays x 64 samples): raymarching.cu:1916-1974 vectorised over rays."""
    s = sigmas.view(n_rays, n_per_ray); dl = deltas.view(n_rays, n_per_ray, 2)
    alpha = 1 - torch.exp(-s * dl[:, :, 0])
    T = torch.cumprod(torch.cat([torch.ones(n_rays, 1), 1 - alpha], dim=1), dim=1)
    stopped = (T[:, 1:] < T_thresh).long().cumsum(dim=1)
    live = torch.cat([torch.ones(n_rays, 1, dtype=torch.bool), stopped[:, :-1] == 0], dim=1)
    w = alpha * T[:, :-1] * live
    image = (w[:, :, None] * rgbs.view(n_rays, n_per_ray, 3)).sum(1)
    return w.sum(1), (aud.view(n_rays, -1) * live).sum(1), (eye.view(n_rays, -1) * live).sum(1), (w * unc.view(n_rays, -1)).sum(1), (w * dl[:, :, 1]).sum(1), image


def composite_rays_triplane(alive, rays_t, t_after, sig, rgb, deltas, aud, eye, unc, state, T_thresh):
    """kernel_composite_rays_triplane (raymarching.cu:2142-2249) over n_step samples per alive ray; updates `state` in place,
    returns the surviving ids."""
    n, n_step = deltas.shape[:2]
    ws, d, img, a0, a1, u = (state[k][alive] for k in ("ws", "depth", "image", "aud", "eye", "unc"))
    running = torch.ones(n, dtype=torch.bool)
    last_t = rays_t[alive].clone()
    for k in range(n_step):
        dl = deltas[:, k, 0]
        running = running & (dl != 0)
        alpha = 1 - torch.exp(-sig[:, k] * dl)
        T = 1 - ws
        w = torch.where(running, alpha * T, torch.zeros(()))
        ws = ws + w
        last_t = torch.where(running, deltas[:, k, 1], last_t)
        d = d + w * deltas[:, k, 1]
        img = img + w[:, None] * rgb[:, k]
        a0 = a0 + torch.where(running, aud[:, k], torch.zeros(()))
        a1 = a1 + torch.where(running, eye[:, k], torch.zeros(()))
        u = u + w * unc[:, k]
        running = running & ~(T < T_thresh)
    for key, val in (("ws", ws), ("depth", d), ("image", img), ("aud", a0), ("eye", a1), ("unc", u)):
        state[key][alive] = val
    keep = running
    rays_t[alive[keep]] = last_t[keep]
    return alive[keep]


def render_frame(p, rays_o, rays_d, bitfield, enc_a, c, e, aabb, max_steps=16, dt_gamma=1 / 256, min_near=0.05, T_thresh=1e-4, bound=1.0):
    """run_cuda_for_inference (renderer.py:406-570) in pure PyTorch on the host.  Returns image [N,3], weights_sum, depth, #samples."""
    N = rays_o.shape[0]
    nears, fars = near_far_from_aabb(rays_o, rays_d, aabb, min_near)
    state = dict(ws=torch.zeros(N), depth=torch.zeros(N), image=torch.zeros(N, 3), aud=torch.zeros(N), eye=torch.zeros(N), unc=torch.zeros(N))
    alive = torch.arange(N)
    rays_t = nears.clone()
    step, n_samples = 0, 0
    while step < max_steps and alive.numel() > 0:
        n_step = max(min(N // alive.numel(), 8), 1)
        xyzs, dirs, deltas, t_after = march_rays(alive, rays_t, rays_o, rays_d, fars, bitfield, n_step, bound, dt_gamma, max_steps)
        flat = lambda a: a.reshape(-1, a.shape[-1])
        sig, rgb, aud, eye, unc = head_forward(p, flat(xyzs), flat(dirs), enc_a, c, e, testing=True)
        n_samples += int((deltas[:, :, 0] != 0).sum())
        n = alive.numel()
        alive = composite_rays_triplane(alive, rays_t, t_after, sig.view(n, n_step), rgb.view(n, n_step, 3), deltas, aud.view(n, n_step), eye.view(n, n_step),
                                        unc.view(n, n_step), state, T_thresh)
        step += n_step
    image = (state["image"] + (1 - state["ws"])[:, None]).clamp(0, 1)
    return image, state["ws"], state["depth"], n_samples


def params_from_state_dict(sd, bound=1.0):
    """Head-model parameters as fp32 CPU tensors + grid geometry (from a HeadModel / reference NeRFNetwork state_dict)."""
    p = {k: v.detach().float().cpu() for k, v in sd.items() if torch.is_tensor(v) and v.dtype.is_floating_point}
    p["encoder_xy.offsets"] = sd["encoder_xy.offsets"].cpu()
    p["S"] = float(np.float32(np.log2(np.exp2(np.log2(512 * bound / 64) / 11))))
    p["H"] = 64
    p["bound"] = bound
    return p


def composite_rays_train_triplane_ragged(sigmas, rgbs, aud, eye, unc, deltas, rays, max_steps=16, T_thresh=1e-4):
    """Variable-length segments (rays [N,3] = (id, offset, count)), differentiable: pads every ray to `max_steps` samples with sigma = 0
    (alpha = 0, no contribution) and reuses the equal-length formulation.  Outputs are indexed by ray id like the kernel (raymarching.cu:1967-1974)."""
    n = rays.shape[0]
    ids, off, cnt = rays[:, 0].long(), rays[:, 1].long(), rays[:, 2].long()
    k = torch.arange(max_steps)[None, :]
    valid = k < cnt[:, None]
    idx = (off[:, None] + k).clamp_max(max(sigmas.shape[0] - 1, 0))
    pick = lambda a: torch.where(valid, a[idx], torch.zeros(()))
    s, a0, a1, u = pick(sigmas), pick(aud), pick(eye), pick(unc)
    dl0, dl1 = pick(deltas[:, 0]), pick(deltas[:, 1])
    c = torch.where(valid[:, :, None], rgbs[idx], torch.zeros(()))
    alpha = 1 - torch.exp(-s * dl0)
    T = torch.cumprod(torch.cat([torch.ones(n, 1), 1 - alpha], dim=1), dim=1)
    stopped = (T[:, 1:] < T_thresh).long().cumsum(dim=1)
    live = torch.cat([torch.ones(n, 1, dtype=torch.bool), stopped[:, :-1] == 0], dim=1) & valid
    w = alpha * T[:, :-1] * live
    outs = [w.sum(1), (a0 * live).sum(1), (a1 * live).sum(1), (w * u).sum(1), (w * dl1).sum(1), (w[:, :, None] * c).sum(1)]
    res = []
    for o in outs:                                   # scatter by ray id
        z = torch.zeros_like(o)
        res.append(z.index_copy(0, ids, o))
    return tuple(res)

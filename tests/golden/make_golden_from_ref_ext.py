#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CUDA EXTENSIONS on a B200.

    gpurun -- 'python tests/golden/make_golden_from_ref_ext.py gpurun_out/golden'     (then copy into tests/golden/)

The extensions are the reference's unmodified .cu files compiled for sm_100a by oracle/build_ref_ext.sh
(oracle/_ref/*.so).  Inputs are the seeded cases of tests/cases.py; outputs are what the reference kernels wrote,
re-ordered only where the reference itself is order-nondeterministic (march_rays_train's atomic allocation ->
canonical ray order, tests/refext.py:canonical).  These files pin the CPU oracle (tests/test_oracle_golden.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import conftest  # noqa: E402,F401  (sets sys.path)
import cases  # noqa: E402
import refext  # noqa: E402
from refext import T  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE)
os.makedirs(out_dir, exist_ok=True)
assert refext.available(), "oracle/_ref/*.so missing: run oracle/build_ref_ext.sh where /root/reference exists"
rm, ge, sh, fq = (refext.mod(n) for n in ("_ref_raymarching_face", "_ref_grid_encoder", "_ref_sh_encoder", "_ref_freqencoder"))
npy = lambda t: t.detach().cpu().numpy()


def save(name, **arrs):
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrs)
    print(f"[golden] {name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in arrs.items()))


# ---- utils ------------------------------------------------------------------------------------------------------
rng = np.random.default_rng(7)
coords = rng.integers(0, 128, (4096, 3)).astype(np.int32)
idx = torch.empty(4096, dtype=torch.int32, device="cuda"); rm.morton3D(T(coords), 4096, idx)
back = torch.empty(4096, 3, dtype=torch.int32, device="cuda"); rm.morton3D_invert(idx, 4096, back)
g32 = (rng.random((1, 32 ** 3)) * 20).astype(np.float32)
bits = torch.empty(32 ** 3 // 8, dtype=torch.uint8, device="cuda"); rm.packbits(T(g32), 32 ** 3 // 8, 10.0, bits)
g16 = rng.standard_normal((2, 16 ** 3)).astype(np.float32)
dil = torch.empty(2, 16 ** 3, device="cuda"); rm.morton3D_dilation(T(g16), 2, 16, dil)
so = (rng.standard_normal((256, 3)) * 0.3).astype(np.float32); sd = cases.dirs_case(256, 9)
sph = torch.empty(256, 2, device="cuda"); rm.sph_from_ray(T(so), T(sd), 2.0, 256, sph)
save("utils", morton_in=coords, morton=npy(idx), morton_back=npy(back), pack_in=g32, pack=npy(bits), dil_in=g16, dil=npy(dil),
     sph_o=so, sph_d=sd, sph=npy(sph))

# ---- marching -----------------------------------------------------------------------------------------------------
head = None
for name in ("head16", "cascade2", "nogamma"):
    c = cases.march_case(name)
    nears, fars = refext.near_far(c["rays_o"], c["rays_d"], c["aabb"], c["min_near"])
    xyzs, dirs, deltas, rays, counter = refext.march_train(c, nears, fars)
    counts, cx, cd, cl = refext.canonical(xyzs, dirs, deltas, rays)
    save("march_" + name, nears=npy(nears), fars=npy(fars), counts=counts, xyzs=cx, dirs=cd, deltas=cl, counter=npy(counter))
    if name == "head16":
        head = (c, nears, fars, counts, cx, cd, cl)
    # capped buffer: rays whose segment would overflow M are dropped (raymarching.cu:457); which ones depends on the
    # atomic order, so only the invariant totals are stored
    # inference march: first 512 rays, 4 steps each, resuming from nears
    n_alive, n_step = min(256, c["rays_o"].shape[0]), 4
    alive = torch.arange(n_alive, dtype=torch.int32, device="cuda")
    M = n_alive * n_step; M += 128 - M % 128
    ox, od, ol = (torch.zeros(M, k, device="cuda") for k in (3, 3, 2))
    noises = T(c["noises"][:n_alive])
    rm.march_rays(n_alive, n_step, alive, nears.clone(), T(c["rays_o"]), T(c["rays_d"]), c["bound"], c["dt_gamma"], c["max_steps"], c["C"], c["H"],
                  T(c["bitfield"]), nears, fars, ox, od, ol, noises)
    save("march_infer_" + name, xyzs=npy(ox), dirs=npy(od), deltas=npy(ol))

# ---- composites (on the head16 canonical segments) -----------------------------------------------------------------------
c, nears, fars, counts, cx, cd, cl = head
n = len(counts); offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
rays = np.stack([np.arange(n, dtype=np.int32), offs, counts.astype(np.int32)], 1)
m = int(counts.sum())
f = cases.sample_fields(m, 21); g = cases.ray_grads(n, 22)
tr = {k: T(v) for k, v in f.items()}; tg = {k: T(v) for k, v in g.items()}
t_rays, t_deltas = T(rays), T(cl)
Z = lambda *s: torch.zeros(*s, device="cuda"); E = lambda *s: torch.empty(*s, device="cuda")
for thresh_name, Tt in (("", 1e-4), ("_T1e-1", 1e-1)):
    ws, a0, a1, us, dep, img = E(n), E(n), E(n), E(n), E(n), E(n, 3)
    rm.composite_rays_train_forward(tr["sigmas"], tr["rgbs"], tr["amb_aud"], t_deltas, t_rays, m, n, Tt, ws, a0, dep, img)
    gs, gr, ga = Z(m), Z(m, 3), Z(m)
    rm.composite_rays_train_backward(tg["g_ws"], tg["g_aud"], tg["g_img"], tr["sigmas"], tr["rgbs"], tr["amb_aud"], t_deltas, t_rays, ws, a0, img, m, n, Tt, gs, gr, ga)
    save("composite_plain" + thresh_name, ws=npy(ws), a0=npy(a0), depth=npy(dep), image=npy(img), gs=npy(gs), grgb=npy(gr), ga0=npy(ga))
    ws, a0, dep, img = E(n), E(n), E(n), E(n, 3)
    rm.composite_rays_train_sigma_forward(tr["sigmas"], tr["rgbs"], tr["amb_aud"], t_deltas, t_rays, m, n, Tt, ws, a0, dep, img)
    gs, gr, ga = Z(m), Z(m, 3), Z(m)
    rm.composite_rays_train_sigma_backward(tg["g_ws"], tg["g_aud"], tg["g_img"], tr["sigmas"], tr["rgbs"], tr["amb_aud"], t_deltas, t_rays, ws, a0, img, m, n, Tt, gs, gr, ga)
    save("composite_sigma" + thresh_name, ws=npy(ws), a0=npy(a0), depth=npy(dep), image=npy(img), gs=npy(gs), grgb=npy(gr), ga0=npy(ga))
    ws, a0, us, dep, img = E(n), E(n), E(n), E(n), E(n, 3)
    rm.composite_rays_train_uncertainty_forward(tr["sigmas"], tr["rgbs"], tr["amb_aud"], tr["unc"], t_deltas, t_rays, m, n, Tt, ws, a0, us, dep, img)
    gs, gr, ga, gu = Z(m), Z(m, 3), Z(m), Z(m)
    rm.composite_rays_train_uncertainty_backward(tg["g_ws"], tg["g_aud"], tg["g_unc"], tg["g_img"], tr["sigmas"], tr["rgbs"], tr["amb_aud"], tr["unc"], t_deltas,
                                                 t_rays, ws, a0, us, img, m, n, Tt, gs, gr, ga, gu)
    save("composite_uncertainty" + thresh_name, ws=npy(ws), a0=npy(a0), us=npy(us), depth=npy(dep), image=npy(img), gs=npy(gs), grgb=npy(gr), ga0=npy(ga), gu=npy(gu))
    ws, a0, a1, us, dep, img = E(n), E(n), E(n), E(n), E(n), E(n, 3)
    rm.composite_rays_train_triplane_forward(tr["sigmas"], tr["rgbs"], tr["amb_aud"], tr["amb_eye"], tr["unc"], t_deltas, t_rays, m, n, Tt, ws, a0, a1, us, dep, img)
    gs, gr, ga, ge_, gu = Z(m), Z(m, 3), Z(m), Z(m), Z(m)
    rm.composite_rays_train_triplane_backward(tg["g_ws"], tg["g_aud"], tg["g_eye"], tg["g_unc"], tg["g_img"], tr["sigmas"], tr["rgbs"], tr["amb_aud"], tr["amb_eye"],
                                              tr["unc"], t_deltas, t_rays, ws, a0, a1, us, img, m, n, Tt, gs, gr, ga, ge_, gu)
    save("composite_triplane" + thresh_name, ws=npy(ws), a0=npy(a0), a1=npy(a1), us=npy(us), depth=npy(dep), image=npy(img), gs=npy(gs), grgb=npy(gr),
         ga0=npy(ga), ga1=npy(ge_), gu=npy(gu))

# inference composites: 3 chained iterations of (march n_step, composite) for each variant on the first 512 head16 rays
cH = head[0]
for variant, fn, n_amb, unc in (("rgb", rm.composite_rays, 0, False), ("plain", rm.composite_rays_ambient, 1, False),
                                ("sigma", rm.composite_rays_ambient_sigma, 1, False), ("uncertainty", rm.composite_rays_uncertainty, 1, True),
                                ("triplane", rm.composite_rays_triplane, 2, True)):
    N0 = 512
    alive = torch.arange(N0, dtype=torch.int32, device="cuda"); rays_t = nears.clone()
    ws, dep, img, s0, s1, su = Z(N0 * 2)[:N0 * 2], Z(N0 * 2), Z(N0 * 2, 3), Z(N0 * 2), Z(N0 * 2), Z(N0 * 2)
    trace = {}
    for it, n_step in enumerate((1, 2, 4)):
        n_alive = alive.shape[0]
        M = n_alive * n_step; M += 128 - M % 128
        ox, od, ol = (torch.zeros(M, k, device="cuda") for k in (3, 3, 2))
        rm.march_rays(n_alive, n_step, alive, rays_t, T(cH["rays_o"]), T(cH["rays_d"]), cH["bound"], cH["dt_gamma"], cH["max_steps"], cH["C"], cH["H"],
                      T(cH["bitfield"]), nears, fars, ox, od, ol, Z(n_alive))
        fs = {k: T(v) for k, v in cases.sample_fields(M, 300 + it, scale_sigma=30.0).items()}
        ex = [fs["amb_aud"], fs["amb_eye"]][:n_amb] + ([fs["unc"]] if unc else [])
        sums = [s0, s1][:n_amb] + ([su] if unc else [])
        fn(n_alive, n_step, 1e-2, alive, rays_t, fs["sigmas"], fs["rgbs"], ol, *ex, ws, dep, img, *sums)
        trace[f"alive{it}"] = npy(alive).copy(); trace[f"rays_t{it}"] = npy(rays_t).copy()
        alive = alive[alive >= 0]
    save("composite_infer_" + variant, ws=npy(ws), depth=npy(dep), image=npy(img), s0=npy(s0), s1=npy(s1), su=npy(su), **trace)

# ---- grid encoder -------------------------------------------------------------------------------------------------
for name in cases.GRID_CASES:
    c = cases.grid_case(name)
    B, D, L, C = c["inputs"].shape[0], c["D"], c["L"], c["C"]
    dt = torch.float16 if c["half"] else torch.float32
    emb, x, offs = T(c["embeddings"]), T(c["inputs"]), T(c["offsets"])
    out = torch.empty(L, B, C, device="cuda", dtype=dt); dy = torch.empty(B, L * D * C, device="cuda", dtype=dt)
    ge.grid_encode_forward(x, emb, offs, out, B, D, C, L, c["S"], c["H"], dy, c["gridtype"], c["align_corners"])
    gemb = torch.zeros_like(emb); gin = torch.zeros(B, D, device="cuda", dtype=dt)
    ge.grid_encode_backward(T(c["grad"]), x, emb, offs, gemb, B, D, C, L, c["S"], c["H"], dy, gin, c["gridtype"], c["align_corners"])
    save("grid_" + name, outputs=npy(out), dy_dx=npy(dy), grad_embeddings=npy(gemb), grad_inputs=npy(gin))

# per-level scales as the GPU computes them (through libb2nerf's diagnostic entry point; same ex2.approx + fma as the reference kernel —
# verified by the bit-exact grid outputs above) so the CPU oracle can reproduce the fp32 encodings bit for bit
from gridencoder.backend import grid_level_scales  # noqa: E402
sc = {}
for name in cases.GRID_CASES:
    c = cases.grid_case(name)
    sc[f"S{np.float32(c['S']).view(np.uint32):08x}_H{c['H']}_L{c['L']}"] = npy(grid_level_scales(c["S"], c["H"], c["L"]))
save("level_scales", **sc)

# ---- SH / freq ----------------------------------------------------------------------------------------------------
dirs = cases.dirs_case(256, 3); arrs = {}
raw = (np.random.default_rng(4).standard_normal((64, 3)) * 0.7).astype(np.float32)       # un-normalised inputs: polynomials, not unit-sphere values
for deg in range(1, 9):
    for tag, v in (("unit", dirs), ("raw", raw)):
        B = v.shape[0]
        o = torch.empty(B, deg * deg, device="cuda"); dy = torch.empty(B, 3 * deg * deg, device="cuda")
        sh.sh_encode_forward(T(v), o, B, 3, deg, dy)
        arrs[f"{tag}_out{deg}"] = npy(o); arrs[f"{tag}_dy{deg}"] = npy(dy)
        if deg in (4, 8):
            gi = torch.zeros(B, 3, device="cuda"); gr = T(np.random.default_rng(deg).standard_normal((B, deg * deg)).astype(np.float32))
            sh.sh_encode_backward(gr, T(v), B, 3, deg, dy, gi); arrs[f"{tag}_gin{deg}"] = npy(gi)
save("sh", **arrs)
fx = (np.random.default_rng(5).random((256, 6)) * 2 - 1).astype(np.float32); arrs = {}
for D, deg in ((2, 8), (6, 3), (3, 10)):
    Cc = D + 2 * D * deg; v = np.ascontiguousarray(fx[:, :D])
    o = torch.empty(256, Cc, device="cuda"); fq.freq_encode_forward(T(v), 256, D, deg, Cc, o)
    gr = T(np.random.default_rng(D).standard_normal((256, Cc)).astype(np.float32)); gi = torch.zeros(256, D, device="cuda")
    fq.freq_encode_backward(gr, o, 256, D, deg, Cc, gi)
    arrs[f"out_{D}_{deg}"] = npy(o); arrs[f"gin_{D}_{deg}"] = npy(gi)
save("freq", **arrs)
torch.cuda.synchronize()
print("[golden] done ->", out_dir)

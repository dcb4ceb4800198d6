"""Thin numpy-in / numpy-out drivers of the REFERENCE's own CUDA extensions (oracle/_ref/*.so, built from
/root/reference in place by oracle/build_ref_ext.sh).  GPU-only test infrastructure: used to generate golden vectors
and as the on-GPU parity oracle.  Allocation/zero-fill follows the reference's Python wrappers (raymarching.py,
grid.py, sphere_harmonics.py, freq.py)."""
import numpy as np
import torch

from conftest import load_ref_ext

_mods = {}


def mod(name):
    if name not in _mods:
        _mods[name] = load_ref_ext(name)
    return _mods[name]


def available():
    return all(mod(n) is not None for n in ("_ref_raymarching_face", "_ref_grid_encoder", "_ref_sh_encoder", "_ref_freqencoder"))


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def near_far(rays_o, rays_d, aabb, min_near):
    rm = mod("_ref_raymarching_face")
    o, d = T(rays_o), T(rays_d)
    n = o.shape[0]
    nears, fars = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    rm.near_far_from_aabb(o, d, T(aabb), n, min_near, nears, fars)
    return nears, fars


def march_train(case, nears, fars, M=None):
    """Runs the reference kernel; returns its raw outputs (atomic order) as torch tensors."""
    rm = mod("_ref_raymarching_face")
    o, d = T(case["rays_o"]), T(case["rays_d"])
    n = o.shape[0]
    M = n * case["max_steps"] if M is None else M
    xyzs, dirs, deltas = (torch.zeros(M, k, device="cuda") for k in (3, 3, 2))
    rays = torch.empty(n, 3, dtype=torch.int32, device="cuda")
    counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    rm.march_rays_train(o, d, T(case["bitfield"]), case["bound"], case["dt_gamma"], case["max_steps"], n, case["C"], case["H"], M,
                        nears, fars, xyzs, dirs, deltas, rays, counter, T(case["noises"]))
    torch.cuda.synchronize()
    return xyzs, dirs, deltas, rays, counter


def canonical(xyzs, dirs, deltas, rays, M=None):
    """Re-order a march_rays_train result into ray order with contiguous segments (numpy): the canonical layout the
    deterministic allocators (oracle, libb2nerf) produce.  Rays dropped for exceeding M keep their count."""
    xyzs, dirs, deltas, rays = (t.cpu().numpy() if torch.is_tensor(t) else t for t in (xyzs, dirs, deltas, rays))
    order = np.argsort(rays[:, 0], kind="stable")
    r = rays[order]
    assert (r[:, 0] == np.arange(len(r))).all()
    M = xyzs.shape[0] if M is None else M
    segs_x, segs_d, segs_l = [], [], []
    for _, off, cnt in r:
        if cnt > 0 and off + cnt <= M:
            segs_x.append(xyzs[off:off + cnt]); segs_d.append(dirs[off:off + cnt]); segs_l.append(deltas[off:off + cnt])
    cat = lambda s, k: np.concatenate(s, 0) if s else np.zeros((0, k), np.float32)
    return r[:, 2].copy(), cat(segs_x, 3), cat(segs_d, 3), cat(segs_l, 2)

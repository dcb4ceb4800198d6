"""Seeded input cases shared by the golden generator (tests/golden/make_golden_from_ref_ext.py, runs the reference's
own CUDA extension on a B200) and the parity tests (CPU oracle vs golden, CUDA path vs oracle / vs reference ext)."""
import numpy as np

from b2nerf import scene


def march_case(name):
    """Returns dict of numpy inputs for march_rays_train / march_rays."""
    if name == "head16":           # the real run config: bound 1, 1 cascade, max_steps 16, dt_gamma 1/256 (train.py:35,71,75)
        n, bound, C, max_steps, dt_gamma, seed = 1024, 1.0, 1, 16, 1.0 / 256, 0
        bitfield = scene.bitfield_from_grid(scene.density_grid())
        aabb = scene.AABB
    elif name == "cascade2":       # exercises the mip-level logic: bound 2 -> 2 cascades, longer rays, growing dt
        n, bound, C, max_steps, dt_gamma, seed = 384, 2.0, 2, 64, 1.0 / 128, 1
        g = scene.density_grid()
        g2 = scene.density_grid(sigma=0.6, peak=14.0)      # coarser blob for cascade 1
        bitfield = scene.bitfield_from_grid(np.concatenate([g, g2], 0))
        aabb = np.array([-2, -1, -2, 2, 1, 2], np.float32)
    elif name == "nogamma":        # dt_gamma = 0 and a large step budget: fixed dt_min stepping
        n, bound, C, max_steps, dt_gamma, seed = 256, 1.0, 1, 1024, 0.0, 2
        bitfield = scene.bitfield_from_grid(scene.density_grid())
        aabb = scene.AABB
    else:
        raise KeyError(name)
    o, d = scene.train_rays(step=seed, n=n, seed=seed)
    rng = np.random.default_rng(100 + seed)
    # degenerate rays: axis-parallel (a zero direction component -> 1/d = inf), a ray that misses the box
    d = d.copy(); o = o.copy()
    d[0] = [0.0, 0.0, -1.0]; o[0] = [0.1, 0.05, 3.0]
    d[1] = [0.0, -0.6, -0.8]; o[1] = [0.0, 2.0, 3.0]
    d[2] = [1.0, 0.0, 0.0]; o[2] = [-3.0, 0.9, 0.0]            # misses (|y| > 0.5)
    noises = rng.random(n).astype(np.float32)
    return dict(rays_o=o, rays_d=d, bitfield=bitfield, aabb=aabb, bound=bound, C=C, H=128, max_steps=max_steps,
                dt_gamma=dt_gamma, min_near=scene.MIN_NEAR, noises=noises)


def sample_fields(m, seed, scale_sigma=8.0):
    """Random per-sample network outputs for the composite cases."""
    rng = np.random.default_rng(seed)
    return dict(sigmas=(rng.random(m) ** 2 * scale_sigma * 40).astype(np.float32), rgbs=rng.random((m, 3)).astype(np.float32),
                amb_aud=rng.random(m).astype(np.float32), amb_eye=rng.random(m).astype(np.float32),
                unc=(rng.random(m) * 0.5).astype(np.float32))


def ray_grads(n, seed):
    rng = np.random.default_rng(seed)
    return dict(g_ws=rng.standard_normal(n).astype(np.float32), g_aud=rng.standard_normal(n).astype(np.float32),
                g_eye=rng.standard_normal(n).astype(np.float32), g_unc=rng.standard_normal(n).astype(np.float32),
                g_img=rng.standard_normal((n, 3)).astype(np.float32))


GRID_CASES = {
    # name: (D, L, C, base_res, log2_hashmap, desired_res, gridtype, align_corners, B, half)
    "triplane": (2, 12, 1, 64, 14, 512, 0, False, 1024, False),      # network.py:129-133
    "hash3d": (3, 8, 2, 16, 12, 256, 0, False, 512, False),
    "torso_tiled_half": (2, 16, 2, 16, 16, 2048, 1, False, 512, True),   # network.py:166 under autocast
    "tiled_ac": (2, 6, 4, 8, 10, 64, 1, True, 256, False),
    "d1c8": (1, 4, 8, 16, 8, 128, 0, False, 256, False),
}


def grid_case(name):
    from gridencoder.grid import level_table
    D, L, C, H, log2T, desired, gridtype, ac, B, half = GRID_CASES[name]
    pls = np.exp2(np.log2(desired / H) / (L - 1))
    offsets = np.array(level_table(D, L, pls, H, log2T, ac), np.int32)
    rng = np.random.default_rng({"triplane": 11, "hash3d": 12, "torso_tiled_half": 13, "tiled_ac": 14, "d1c8": 15}[name])
    x = rng.random((B, D)).astype(np.float32)
    x[0] = 0.0; x[1] = 1.0; x[2] = -0.25; x[3, 0] = 1.5            # edges + out-of-range rows (gridencoder.cu:98-122)
    emb = rng.uniform(-1, 1, (int(offsets[-1]), C)).astype(np.float32)
    if half:
        emb = emb.astype(np.float16)
    grad = rng.standard_normal((L, B, C)).astype(np.float16 if half else np.float32)
    return dict(inputs=x, embeddings=emb, offsets=offsets, S=float(np.log2(pls)), H=H, D=D, L=L, C=C, gridtype=gridtype,
                align_corners=ac, half=half, grad=grad)


def dirs_case(n=512, seed=3):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    v[0] = [0, 0, 1]; v[1] = [1, 0, 0]; v[2] = [0, -1, 0]
    return v.astype(np.float32)

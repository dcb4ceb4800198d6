"""GPU parity tests proper: the CUDA path, called through the drop-in operator surface / C ABI, against
  (a) the CPU oracle (oracle/oracle.c) on the seeded cases of tests/cases.py, and
  (b) the reference's OWN extension built for sm_100a (oracle/_ref/*.so) when it travelled to the box.
Integer / index / sample-position results must be bit-exact; float tolerances are written at each assert."""
import numpy as np
import pytest
import torch

import cases
import oracle
from b2nerf import scene

pytestmark = pytest.mark.gpu


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def npy(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def rm():
    import raymarching
    return raymarching


@pytest.fixture(scope="module")
def ref():
    import refext
    return refext if refext.available() else None


# ---------------------------------------------------------------------------------------------------------------
# utils
# ---------------------------------------------------------------------------------------------------------------
def test_morton_packbits_dilation_exact(rm):
    rng = np.random.default_rng(0)
    coords = rng.integers(0, 1024, (100003, 3)).astype(np.int32)
    idx = npy(rm.morton3D(T(coords)))
    assert np.array_equal(idx, oracle.morton3D(coords))
    assert np.array_equal(npy(rm.morton3D_invert(T(idx))), oracle.morton3D_invert(idx))
    # known answers
    ka = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1], [2, 0, 0], [3, 5, 7], [64, 64, 64], [127, 0, 0], [0, 127, 0],
                   [0, 0, 127], [100, 37, 90], [127, 127, 127]], np.int32)
    assert npy(rm.morton3D(T(ka))).tolist() == [0, 1, 2, 4, 7, 8, 431, 1835008, 299593, 599186, 1198372, 1427682, 2097151]
    g = scene.density_grid()
    bits = npy(rm.packbits(T(g), 10.0))
    assert np.array_equal(bits, scene.bitfield_from_grid(g)) and np.array_equal(bits, oracle.packbits(g, 10.0))
    assert npy(rm.packbits(T(np.array([[0, 11, 10, 10.0001, 9, 50, -1, 10]], np.float32)), 10.0)).tolist() == [42]
    odd = (rng.random((1, 8 * 37)) * 20).astype(np.float32)            # byte count not a multiple of 4: tail path
    assert np.array_equal(npy(rm.packbits(T(odd), 10.0)), oracle.packbits(odd, 10.0))
    into = torch.zeros(37, dtype=torch.uint8, device="cuda")
    assert rm.packbits(T(odd), 10.0, into).data_ptr() == into.data_ptr() and np.array_equal(npy(into), oracle.packbits(odd, 10.0))
    g2 = rng.standard_normal((2, 32 ** 3)).astype(np.float32)
    assert np.array_equal(npy(rm.morton3D_dilation(T(g2))), oracle.morton3D_dilation(g2, 2, 32))


def test_near_far_exact(rm):
    c = cases.march_case("head16")
    nears, fars = rm.near_far_from_aabb(T(c["rays_o"]), T(c["rays_d"]), T(c["aabb"]), c["min_near"])
    on, of = oracle.near_far_from_aabb(c["rays_o"], c["rays_d"], c["aabb"], c["min_near"])
    assert np.array_equal(npy(nears), on) and np.array_equal(npy(fars), of)
    assert npy(nears)[2] == np.finfo(np.float32).max          # the ray that misses the box
    o, d = scene.frame_rays(0)                                 # full 512x512 frame
    nears, fars = rm.near_far_from_aabb(T(o), T(d), T(scene.AABB), scene.MIN_NEAR)
    on, of = oracle.near_far_from_aabb(o, d, scene.AABB, scene.MIN_NEAR)
    assert np.array_equal(npy(nears), on) and np.array_equal(npy(fars), of)


def test_sph_from_ray(rm):
    rng = np.random.default_rng(1)
    o = (rng.standard_normal((4096, 3)) * 0.3).astype(np.float32); d = cases.dirs_case(4096, 2)
    got = npy(rm.sph_from_ray(T(o), T(d), 2.0))
    np.testing.assert_allclose(got, oracle.sph_from_ray(o, d, 2.0), rtol=0, atol=2e-6)     # atan2f/sqrtf: device vs libm


# ---------------------------------------------------------------------------------------------------------------
# marching
# ---------------------------------------------------------------------------------------------------------------
def _march_ours(rm, c, nears, fars, M):
    from raymarching.backend import _backend
    o, d = T(c["rays_o"]), T(c["rays_d"]); n = o.shape[0]
    xyzs, dirs, deltas = (torch.zeros(M, k, device="cuda") for k in (3, 3, 2))
    rays = torch.empty(n, 3, dtype=torch.int32, device="cuda"); counter = torch.zeros(2, dtype=torch.int32, device="cuda")
    _backend.march_rays_train(o, d, T(c["bitfield"]), c["bound"], c["dt_gamma"], c["max_steps"], n, c["C"], c["H"], M, nears, fars,
                              xyzs, dirs, deltas, rays, counter, T(c["noises"]))
    return xyzs, dirs, deltas, rays, counter


@pytest.mark.parametrize("name", ["head16", "cascade2", "nogamma"])
def test_march_rays_train_bit_exact(rm, ref, name):
    c = cases.march_case(name)
    n = c["rays_o"].shape[0]; M = n * min(c["max_steps"], 256)
    nears, fars = rm.near_far_from_aabb(T(c["rays_o"]), T(c["rays_d"]), T(c["aabb"]), c["min_near"])
    xyzs, dirs, deltas, rays, counter = _march_ours(rm, c, nears, fars, M)
    ox, od, ol, orays, ocnt = oracle.march_rays_train(c["rays_o"], c["rays_d"], c["bitfield"], c["bound"], c["dt_gamma"], c["max_steps"], c["C"],
                                                      c["H"], M, npy(nears), npy(fars), c["noises"])
    assert np.array_equal(npy(rays), orays), "rays (id, offset, count) must be bit-exact and in ray order"
    assert np.array_equal(npy(counter), ocnt)
    assert np.array_equal(npy(xyzs), ox) and np.array_equal(npy(dirs), od) and np.array_equal(npy(deltas), ol)
    assert orays[:, 2].sum() > 0
    if ref is not None:        # the reference's own kernel on the same GPU, canonicalised to ray order
        rx, rd, rl, rrays, rcnt = ref.march_train(c, nears, fars, M)
        counts, cx, cd, cl = ref.canonical(rx, rd, rl, rrays)
        tot = int(counts.sum())
        assert np.array_equal(counts, npy(rays)[:, 2]) and int(npy(rcnt)[0]) == tot
        assert np.array_equal(cx, npy(xyzs)[:tot]) and np.array_equal(cd, npy(dirs)[:tot]) and np.array_equal(cl, npy(deltas)[:tot])


def test_march_rays_train_full_batch_and_cap(rm):
    """BASELINE size (65 536 rays): bit-exact vs the oracle; gap-free partition; M cap drops whole rays (raymarching.cu:457)."""
    o, d = scene.train_rays(0, 65536)
    bf = scene.bitfield_from_grid(scene.density_grid())
    c = dict(rays_o=o, rays_d=d, bitfield=bf, bound=1.0, dt_gamma=scene.DT_GAMMA, max_steps=16, C=1, H=128,
             noises=np.random.default_rng(3).random(65536).astype(np.float32))
    nears, fars = rm.near_far_from_aabb(T(o), T(d), T(scene.AABB), scene.MIN_NEAR)
    for M in (65536 * 16, 200000):
        xyzs, dirs, deltas, rays, counter = _march_ours(rm, c, nears, fars, M)
        ox, od, ol, orays, ocnt = oracle.march_rays_train(o, d, bf, 1.0, scene.DT_GAMMA, 16, 1, 128, M, npy(nears), npy(fars), c["noises"])
        r = npy(rays)
        assert np.array_equal(r, orays) and np.array_equal(npy(counter), ocnt)
        assert np.array_equal(r[:, 1], np.concatenate([[0], np.cumsum(r[:, 2])[:-1]]))          # gap-free partition of [0, counter[0])
        assert np.array_equal(npy(xyzs), ox) and np.array_equal(npy(deltas), ol) and np.array_equal(npy(dirs), od)
    assert ocnt[0] > 200000            # the capped run really dropped rays
    # python-level wrapper: mean_count path pads M to +128 and keeps the buffers
    x2, d2, l2, r2 = rm.march_rays_train(T(o), T(d), 1.0, T(bf), 1, 128, nears, fars, None, int(ocnt[0]), False, 128, False, scene.DT_GAMMA, 16)
    _, _, _, zrays, _ = oracle.march_rays_train(o, d, bf, 1.0, scene.DT_GAMMA, 16, 1, 128, 65536 * 16, npy(nears), npy(fars), np.zeros(65536, np.float32))
    assert x2.shape[0] == int(ocnt[0]) + (128 - int(ocnt[0]) % 128)
    keep = zrays[:, 1] + zrays[:, 2] <= x2.shape[0]            # perturb=False -> zero noise; rays past the estimated buffer are dropped but still counted
    assert np.array_equal(npy(r2), zrays) and keep.sum() > 60000
    x3, _, _, _ = rm.march_rays_train(T(o), T(d), 1.0, T(bf), 1, 128, nears, fars, None, -1, False, 128, False, scene.DT_GAMMA, 16)
    ztot = int(zrays[:, 2].sum())
    assert x3.shape[0] == ztot + (128 - ztot % 128)          # warm-up path trims to the produced samples, 128-aligned


@pytest.mark.parametrize("name", ["head16", "cascade2", "nogamma"])
def test_march_rays_inference_bit_exact(rm, ref, name):
    c = cases.march_case(name)
    o, d = T(c["rays_o"]), T(c["rays_d"]); N = o.shape[0]
    nears, fars = rm.near_far_from_aabb(o, d, T(c["aabb"]), c["min_near"])
    alive = torch.arange(N - 1, -1, -3, dtype=torch.int32, device="cuda")       # a non-trivial compacted id list
    n_alive = alive.shape[0]
    for n_step in (1, 3, 8):
        from raymarching.backend import _backend
        M = n_alive * n_step + 128 - (n_alive * n_step) % 128
        xyzs, dirs, deltas = (torch.zeros(M, k, device="cuda") for k in (3, 3, 2))
        noises = T(c["noises"][:n_alive])
        _backend.march_rays(n_alive, n_step, alive, nears, o, d, c["bound"], c["dt_gamma"], c["max_steps"], c["C"], c["H"], T(c["bitfield"]),
                            nears, fars, xyzs, dirs, deltas, noises)
        ox, od, ol = oracle.march_rays(n_alive, n_step, npy(alive), npy(nears), c["rays_o"], c["rays_d"], c["bound"], c["dt_gamma"], c["max_steps"],
                                       c["C"], c["H"], c["bitfield"], npy(nears), npy(fars), c["noises"][:n_alive], align=128)
        assert np.array_equal(npy(xyzs), ox) and np.array_equal(npy(dirs), od) and np.array_equal(npy(deltas), ol)
        if ref is not None:
            rxyz, rdir, rdel = (torch.zeros(M, k, device="cuda") for k in (3, 3, 2))
            ref.mod("_ref_raymarching_face").march_rays(n_alive, n_step, alive, nears, o, d, c["bound"], c["dt_gamma"], c["max_steps"], c["C"], c["H"],
                                                        T(c["bitfield"]), nears, fars, rxyz, rdir, rdel, noises)
            assert torch.equal(rxyz, xyzs) and torch.equal(rdir, dirs) and torch.equal(rdel, deltas)
    # wrapper: zero-filled, padded past the next multiple of 128 even when aligned (raymarching.py:381-382)
    x, _, dl = rm.march_rays(64, 2, alive, nears, o, d, c["bound"], T(c["bitfield"]), c["C"], c["H"], nears, fars, 128, False, c["dt_gamma"], c["max_steps"])
    assert x.shape[0] == 256 and float(dl[128:].abs().sum()) == 0.0


def test_march_rays_train_backward(rm):
    c = cases.march_case("head16")
    nears, fars = rm.near_far_from_aabb(T(c["rays_o"]), T(c["rays_d"]), T(c["aabb"]), c["min_near"])
    M = 1024 * 16
    xyzs, dirs, deltas, rays, _ = _march_ours(rm, c, nears, fars, M)
    rng = np.random.default_rng(5)
    gx, gd = rng.standard_normal((M, 3)).astype(np.float32), rng.standard_normal((M, 3)).astype(np.float32)
    from raymarching.backend import _backend
    go, gdd = torch.zeros(1024, 3, device="cuda"), torch.zeros(1024, 3, device="cuda")
    _backend.march_rays_train_backward(T(gx), T(gd), rays, deltas, 1024, M, go, gdd)
    ogo, ogd = oracle.march_rays_train_backward(gx, gd, npy(rays), npy(deltas))
    assert np.array_equal(npy(go), ogo) and np.array_equal(npy(gdd), ogd)


# ---------------------------------------------------------------------------------------------------------------
# composites
# ---------------------------------------------------------------------------------------------------------------
VARIANTS = {  # name -> (train fn name, infer fn name, ambient keys, has unc)
    "plain": ("composite_rays_train", "composite_rays_ambient", ["amb_aud"], False),
    "sigma": ("composite_rays_train_sigma", "composite_rays_ambient_sigma", ["amb_aud"], False),
    "uncertainty": ("composite_rays_train_uncertainty", "composite_rays_uncertainty", ["amb_aud"], True),
    "triplane": ("composite_rays_train_triplane", "composite_rays_triplane", ["amb_aud", "amb_eye"], True),
}
REF_TRAIN = {"plain": "composite_rays_train", "sigma": "composite_rays_train_sigma", "uncertainty": "composite_rays_train_uncertainty",
             "triplane": "composite_rays_train_triplane"}


def _segments(counts):
    n = len(counts)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    return np.stack([np.arange(n, dtype=np.int32), offs, counts.astype(np.int32)], 1)


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("layout", ["tiled", "shuffled", "capped"])
@pytest.mark.parametrize("T_thresh", [1e-4, 1e-1])
def test_composite_train_fwd_bwd(rm, ref, variant, layout, T_thresh):
    """tiled = deterministic allocation (shared-memory staged path); shuffled = rows permuted like the reference's atomic
    order (direct path); capped = trailing rays exceed M and must be skipped + zeroed (raymarching.cu:626-634)."""
    rng = np.random.default_rng(42)
    n = 3000
    counts = rng.integers(0, 17, n); counts[rng.random(n) < 0.3] = 0
    rays = _segments(counts); m = int(counts.sum()); M = m
    if layout == "shuffled":
        rays = rays[rng.permutation(n)]
    if layout == "capped":
        M = m - 1000
    f = cases.sample_fields(m, 21); g = cases.ray_grads(n, 22)
    dl = np.stack([np.full(m, 0.027063293, np.float32), 2.0 + 3.0 * rng.random(m).astype(np.float32)], 1).astype(np.float32)
    fn_name, _, amb_keys, has_unc = VARIANTS[variant]
    per = [f[k] for k in amb_keys] + ([f["unc"]] if has_unc else [])
    ins = [T(f["sigmas"])[:M].clone().requires_grad_(), T(f["rgbs"])[:M].clone().requires_grad_()] + [T(p)[:M].clone().requires_grad_() for p in per]
    outs = getattr(rm, fn_name)(*ins, T(dl)[:M], T(rays), T_thresh)
    ofw = oracle.composite_rays_train_forward(variant, f["sigmas"][:M], f["rgbs"][:M], [f[k][:M] for k in amb_keys], f["unc"][:M], dl[:M], rays, T_thresh)
    k = len(amb_keys) + int(has_unc)
    ws, sums, depth, image = outs[0], outs[1:1 + k], outs[1 + k], outs[2 + k]
    # fp32 composite: 1e-5 relative (north star).  The only divergent primitive is the GPU's ex2.approx inside __expf vs libm: alpha = 1 - exp(-s*d)
    # carries an ABSOLUTE error of ~1.2e-7 per sample, hence atol = 16 samples x 1.2e-7 x max|per-sample value| (values <= 1, depth t <= 5).
    tol = dict(rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(npy(ws), ofw["weights_sum"], **tol)
    np.testing.assert_allclose(npy(depth), ofw["depth"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(npy(image), ofw["image"], **tol)
    for i in range(len(amb_keys)):
        np.testing.assert_allclose(npy(sums[i]), ofw["amb_sums"][i], **tol)
    if has_unc:
        np.testing.assert_allclose(npy(sums[-1]), ofw["unc_sum"], **tol)
    # backward (grad_depth is ignored by the reference: pass a non-zero one to prove it is ignored here too)
    g_sums = [g["g_aud"], g["g_eye"]][:len(amb_keys)] + ([g["g_unc"]] if has_unc else [])
    torch.autograd.backward([ws, *sums, depth, image], [T(g["g_ws"]), *[T(x) for x in g_sums], torch.ones_like(depth), T(g["g_img"])])
    obw = oracle.composite_rays_train_backward(variant, g["g_ws"], [g["g_aud"], g["g_eye"]][:len(amb_keys)], g["g_unc"], g["g_img"], f["sigmas"][:M],
                                               f["rgbs"][:M], [f[k][:M] for k in amb_keys], f["unc"][:M], dl[:M], rays, ofw, T_thresh)
    gtol = dict(rtol=2e-4, atol=2e-5)        # differences of O(1) partial sums: abs error ~1e-6 * |grad_image|, documented tolerance
    np.testing.assert_allclose(npy(ins[0].grad), obw["grad_sigmas"], **gtol)
    np.testing.assert_allclose(npy(ins[1].grad), obw["grad_rgbs"], **tol)
    for i in range(len(amb_keys)):
        np.testing.assert_allclose(npy(ins[2 + i].grad), obw["grad_ambs"][i], **tol)
    if has_unc:
        np.testing.assert_allclose(npy(ins[-1].grad), obw["grad_unc"], **tol)
    if ref is None:
        return
    # bit-exact against the reference's kernels on the same GPU (same MUFU.EX2, same op order)
    rmod = ref.mod("_ref_raymarching_face")
    E = lambda *s: torch.empty(*s, device="cuda"); Z = lambda *s: torch.zeros(*s, device="cuda")
    r_ws, r_dep, r_img = E(n), E(n), E(n, 3); r_sums = [E(n) for _ in range(k)]
    d_ins = [t.detach() for t in ins]
    getattr(rmod, REF_TRAIN[variant] + "_forward")(*d_ins, T(dl)[:M], T(rays), M, n, T_thresh, r_ws, *r_sums, r_dep, r_img)
    assert torch.equal(r_ws, ws) and torch.equal(r_dep, depth) and torch.equal(r_img, image)
    for a, b in zip(r_sums, sums):
        assert torch.equal(a, b)
    r_g = [Z(M), Z(M, 3)] + [Z(M) for _ in range(k)]
    getattr(rmod, REF_TRAIN[variant] + "_backward")(T(g["g_ws"]), *[T(x) for x in g_sums], T(g["g_img"]), *d_ins, T(dl)[:M], T(rays), r_ws, *r_sums, r_img,
                                                    M, n, T_thresh, *r_g)
    for a, b in zip(r_g, ins):
        assert torch.equal(a, b.grad)


@pytest.mark.parametrize("cap", ["64", "8"])
def test_composite_train_small_staging_capacity(cap):
    """The training composites stage `cap` samples per pass (default 1024) and take several passes over groups of consecutive rows when a CTA's 128 rays hold
    more; a ray longer than the capacity is walked straight out of global memory.  cap = 64: every CTA of test_composite_train_fwd_bwd takes ~10 passes;
    cap = 8: rays of 9..16 samples take the long-row path.  The capacity is read once per process (B2N_COMP_CAP), so the cases run in a child process."""
    import os, subprocess, sys
    if os.environ.get("B2N_COMP_CAP"):
        pytest.skip("already the child process")
    env = dict(os.environ, B2N_COMP_CAP=cap)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-k", "test_composite_train_fwd_bwd", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=900, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout


@pytest.mark.parametrize("variant", ["rgb", "plain", "sigma", "uncertainty", "triplane"])
def test_composite_inference_loop(rm, ref, variant):
    """Three chained (march, composite) iterations with growing n_step, compaction between them — the renderer.py:503-545 loop."""
    c = cases.march_case("head16")
    o, d = T(c["rays_o"]), T(c["rays_d"]); N = o.shape[0]
    nears, fars = rm.near_far_from_aabb(o, d, T(c["aabb"]), c["min_near"])
    fn = rm.composite_rays if variant == "rgb" else getattr(rm, VARIANTS[variant][1])
    amb_keys = [] if variant == "rgb" else VARIANTS[variant][2]; has_unc = variant != "rgb" and VARIANTS[variant][3]
    Z = lambda *s: torch.zeros(*s, device="cuda")
    alive = torch.arange(N, dtype=torch.int32, device="cuda"); rays_t = nears.clone()
    ws, dep, img = Z(N), Z(N), Z(N, 3); sums = [Z(N) for _ in range(len(amb_keys) + int(has_unc))]
    o_alive = np.arange(N, dtype=np.int32); o_t = npy(nears).copy()
    o_ws, o_dep, o_img = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    o_sums = [np.zeros(N, np.float32) for _ in range(3)]
    r_state = None
    if ref is not None:
        r_alive, r_t = alive.clone(), rays_t.clone(); r_acc = [Z(N), Z(N), Z(N, 3)] + [Z(N) for _ in sums]
        rfn = getattr(ref.mod("_ref_raymarching_face"), "composite_rays" if variant == "rgb" else VARIANTS[variant][1])
    for it, n_step in enumerate((1, 2, 4, 8)):
        n_alive = alive.shape[0]
        assert n_alive == o_alive.shape[0]
        if n_alive == 0:          # renderer.py:509-510 leaves the loop; the reference kernels cannot be launched on an empty grid
            break
        xyzs, dirs, deltas = rm.march_rays(n_alive, n_step, alive, rays_t, o, d, c["bound"], T(c["bitfield"]), c["C"], c["H"], nears, fars, 128, False,
                                           c["dt_gamma"], c["max_steps"])
        M = xyzs.shape[0]
        f = cases.sample_fields(M, 300 + it, scale_sigma=30.0)
        per = [f[k] for k in amb_keys] + ([f["unc"]] if has_unc else [])
        fn(n_alive, n_step, alive, rays_t, T(f["sigmas"]), T(f["rgbs"]), deltas, *[T(p) for p in per], ws, dep, img, *sums, 1e-2)
        ox, od, ol = oracle.march_rays(n_alive, n_step, o_alive, o_t, c["rays_o"], c["rays_d"], c["bound"], c["dt_gamma"], c["max_steps"], c["C"], c["H"],
                                       c["bitfield"], npy(nears), npy(fars), np.zeros(n_alive, np.float32), align=128)
        assert np.array_equal(ol, npy(deltas))
        oracle.composite_rays(variant, n_alive, n_step, 1e-2, o_alive, o_t, f["sigmas"], f["rgbs"], ol, [f[k] for k in amb_keys], f["unc"],
                              o_ws, o_dep, o_img, o_sums[:len(amb_keys)], o_sums[2])
        assert np.array_equal(npy(alive), o_alive), "alive / terminated flags must agree exactly"
        assert np.array_equal(npy(rays_t), o_t)
        if ref is not None:
            rfn(n_alive, n_step, 1e-2, r_alive, r_t, T(f["sigmas"]), T(f["rgbs"]), deltas, *[T(p) for p in per], *r_acc)
            assert torch.equal(r_alive, alive) and torch.equal(r_t, rays_t)
            for a, b in zip(r_acc, [ws, dep, img, *sums]):
                assert torch.equal(a, b)
            r_alive = r_alive[r_alive >= 0]
        alive = alive[alive >= 0]; o_alive = o_alive[o_alive >= 0]
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(npy(ws), o_ws, **tol); np.testing.assert_allclose(npy(dep), o_dep, **tol); np.testing.assert_allclose(npy(img), o_img, **tol)
    for i in range(len(amb_keys)):
        np.testing.assert_allclose(npy(sums[i]), o_sums[i], **tol)
    if has_unc:
        np.testing.assert_allclose(npy(sums[-1]), o_sums[2], **tol)
    assert (o_ws > 0).sum() > 100


# ---------------------------------------------------------------------------------------------------------------
# encoders
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.GRID_CASES))
def test_grid_encode(ref, name):
    from gridencoder.backend import _backend
    c = cases.grid_case(name)
    B, D, L, C = c["inputs"].shape[0], c["D"], c["L"], c["C"]
    dt = torch.float16 if c["half"] else torch.float32
    emb, x, offs = T(c["embeddings"]), T(c["inputs"]), T(c["offsets"])
    out = torch.empty(L, B, C, device="cuda", dtype=dt); dy = torch.empty(B, L * D * C, device="cuda", dtype=dt)
    _backend.grid_encode_forward(x, emb, offs, out, B, D, C, L, c["S"], c["H"], dy, c["gridtype"], c["align_corners"])
    # the oracle is fed the per-level scales the GPU computes (ex2.approx), so the fp32 encoding must match bit for bit
    from gridencoder.backend import grid_level_scales
    oracle.set_level_scales(npy(grid_level_scales(c["S"], c["H"], L)))
    try:
        o_out, o_dy = oracle.grid_encode_forward(c["inputs"], c["embeddings"], c["offsets"], c["S"], c["H"], c["gridtype"], c["align_corners"], True, c["half"])
    finally:
        oracle.set_level_scales(None)
    scale_top = 2.0 ** (c["S"] * (L - 1)) * c["H"]
    if c["half"]:      # fp16 tables: half rounding after every corner; tolerance 1e-3-class (north star), derivative scaled by the top-level resolution
        np.testing.assert_allclose(npy(out).astype(np.float32), o_out.astype(np.float32), rtol=0, atol=2e-3)
        np.testing.assert_allclose(npy(dy).astype(np.float32), o_dy.astype(np.float32), rtol=2e-2, atol=8.0 * scale_top / 64)
    else:
        assert np.array_equal(npy(out), o_out) and np.array_equal(npy(dy), o_dy)
    assert float(out[:, 2].abs().sum()) == 0 and float(out[:, 3].abs().sum()) == 0     # out-of-range rows -> zeros
    gemb = torch.zeros_like(emb); gin = torch.zeros(B, D, device="cuda", dtype=dt)
    _backend.grid_encode_backward(T(c["grad"]), x, emb, offs, gemb, B, D, C, L, c["S"], c["H"], dy, gin, c["gridtype"], c["align_corners"])
    if not c["half"]:
        o_ge, o_gi = oracle.grid_encode_backward(c["grad"], c["inputs"], c["offsets"], C, c["S"], c["H"], c["gridtype"], c["align_corners"], npy(dy))
        # fp32 atomics are order-nondeterministic in the reference too: tolerance vs the double-accumulated oracle
        np.testing.assert_allclose(npy(gemb), o_ge, rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(npy(gin), o_gi, rtol=1e-4, atol=1e-3 * scale_top / 64)
    if ref is None:
        return
    ge = ref.mod("_ref_grid_encoder")
    r_out, r_dy = torch.empty_like(out), torch.empty_like(dy)
    ge.grid_encode_forward(x, emb, offs, r_out, B, D, C, L, c["S"], c["H"], r_dy, c["gridtype"], c["align_corners"])
    assert torch.equal(r_out, out), "forward must be bit-identical to the reference kernel (same ex2.approx scale, same fma order)"
    assert torch.equal(r_dy, dy)
    r_gemb = torch.zeros_like(emb); r_gin = torch.zeros_like(gin)
    ge.grid_encode_backward(T(c["grad"]), x, emb, offs, r_gemb, B, D, C, L, c["S"], c["H"], r_dy, r_gin, c["gridtype"], c["align_corners"])
    np.testing.assert_allclose(npy(gemb).astype(np.float32), npy(r_gemb).astype(np.float32), rtol=1e-4 if not c["half"] else 2e-2, atol=1e-4 if not c["half"] else 5e-2)
    assert torch.equal(r_gin, gin)


@pytest.mark.parametrize("name,B", [("triplane", 40000), ("hash3d", 30000), ("tiled_ac", 20000)])
def test_grid_backward_privatised_path(name, B, monkeypatch):
    """Batches >= 16 384 points take the shared-memory privatised table backward (gridenc.cu:k_grid_bwd_priv): same sums as the direct
    red.global scatter and as the double-accumulated oracle, up to fp32 reassociation; samples are clustered like a training batch
    (many points per coarse cell) and include out-of-range rows."""
    from gridencoder.backend import _backend
    from gridencoder.grid import level_table
    D, L, C, H, log2T, desired, gridtype, ac, _, half = cases.GRID_CASES[name]
    assert not half
    pls = np.exp2(np.log2(desired / H) / (L - 1))
    offsets = np.array(level_table(D, L, pls, H, log2T, ac), np.int32)
    rng = np.random.default_rng(99)
    x = (0.5 + 0.12 * rng.standard_normal((B, D))).astype(np.float32)          # concentrated around the centre; a few rows fall outside [0,1]
    x[:8] = rng.random((8, D)) * 3 - 1
    grad = rng.standard_normal((L, B, C)).astype(np.float32)
    S = float(np.log2(pls))
    xs, gs, offs = T(x), T(grad), T(offsets)
    emb = torch.zeros(int(offsets[-1]), C, device="cuda")
    g_priv, g_direct = torch.zeros_like(emb), torch.zeros_like(emb)
    _backend.grid_encode_backward(gs, xs, emb, offs, g_priv, B, D, C, L, S, H, None, None, gridtype, ac)
    monkeypatch.setenv("B2N_GRID_BWD_DIRECT", "1")
    _backend.grid_encode_backward(gs, xs, emb, offs, g_direct, B, D, C, L, S, H, None, None, gridtype, ac)
    monkeypatch.delenv("B2N_GRID_BWD_DIRECT")
    scale = float(g_direct.abs().max())
    assert scale > 1.0
    np.testing.assert_allclose(npy(g_priv), npy(g_direct), rtol=0, atol=2e-5 * scale)
    from gridencoder.backend import grid_level_scales
    oracle.set_level_scales(npy(grid_level_scales(S, H, L)))
    try:
        o_ge, _ = oracle.grid_encode_backward(grad, x, offsets, C, S, H, gridtype, ac, None)
    finally:
        oracle.set_level_scales(None)
    np.testing.assert_allclose(npy(g_priv), o_ge, rtol=0, atol=2e-5 * scale)
    # accumulation semantics: a second call adds on top (the reference's kernel does atomicAdd into the caller's buffer)
    _backend.grid_encode_backward(gs, xs, emb, offs, g_priv, B, D, C, L, S, H, None, None, gridtype, ac)
    np.testing.assert_allclose(npy(g_priv), 2 * o_ge, rtol=0, atol=6e-5 * scale)


def test_grid_encoder_module_triplane_autograd():
    """The nn.Module surface on the tri-plane config (network.py:129-133): shapes, state_dict names, grads flow to embeddings."""
    from gridencoder import GridEncoder
    enc = GridEncoder(input_dim=2, num_levels=12, level_dim=1, base_resolution=64, log2_hashmap_size=14, desired_resolution=512).cuda()
    assert list(enc.state_dict().keys()) == ["embeddings", "offsets"] and tuple(enc.embeddings.shape) == (163584, 1)
    assert enc.offsets.tolist() == [0, 4232, 10480, 19512, 32512, 48896, 65280, 81664, 98048, 114432, 130816, 147200, 163584]
    enc.embeddings.data.uniform_(-1, 1)
    x = torch.rand(5000, 2, device="cuda") * 2 - 1
    with torch.autocast("cuda", dtype=torch.float16):          # C=1 -> tables stay fp32 under autocast (grid.py:38)
        y = enc(x, bound=1)
    assert y.shape == (5000, 12) and y.dtype == torch.float32
    w = torch.randn_like(y)
    (y * w).sum().backward()
    S = float(np.log2(enc.per_level_scale))
    from gridencoder.backend import grid_level_scales
    oracle.set_level_scales(npy(grid_level_scales(S, 64, 12)))
    try:
        o_out, _ = oracle.grid_encode_forward(npy((x + 1) / 2), npy(enc.embeddings), npy(enc.offsets), S, 64)
    finally:
        oracle.set_level_scales(None)
    assert np.array_equal(npy(y), o_out.transpose(1, 0, 2).reshape(5000, 12))
    o_ge, _ = oracle.grid_encode_backward(npy(w).reshape(5000, 12, 1).transpose(1, 0, 2), npy((x + 1) / 2), npy(enc.offsets), 1, S, 64)
    np.testing.assert_allclose(npy(enc.embeddings.grad), o_ge, rtol=1e-4, atol=1e-4)


def test_sh_and_freq(ref):
    from shencoder import SHEncoder
    from freqencoder import FreqEncoder
    from shencoder.backend import _backend as shb
    from freqencoder.backend import _backend as fqb
    dirs = cases.dirs_case(1000, 3)
    raw = (np.random.default_rng(4).standard_normal((333, 3)) * 0.7).astype(np.float32)
    for deg in range(1, 9):
        for v in (dirs, raw):
            B = v.shape[0]
            o = torch.empty(B, deg * deg, device="cuda"); dy = torch.empty(B, 3 * deg * deg, device="cuda")
            shb.sh_encode_forward(T(v), o, B, 3, deg, dy)
            oo, ody = oracle.sh_encode_forward(v, deg, True)
            mag = max(1.0, float(np.abs(oo).max())); dmag = max(1.0, float(np.abs(ody).max()))
            np.testing.assert_allclose(npy(o), oo, rtol=0, atol=2e-6 * mag)          # fp32 recurrence vs double oracle
            np.testing.assert_allclose(npy(dy), ody, rtol=0, atol=4e-6 * dmag)
            o2 = torch.empty_like(o); shb.sh_encode_forward(T(v), o2, B, 3, deg, None)
            assert torch.equal(o, o2)
            if ref is not None:
                ro = torch.empty_like(o); rdy = torch.empty_like(dy)
                ref.mod("_ref_sh_encoder").sh_encode_forward(T(v), ro, B, 3, deg, rdy)
                np.testing.assert_allclose(npy(o), npy(ro), rtol=0, atol=4e-6 * mag)   # reference's expanded polynomials, fp32
                np.testing.assert_allclose(npy(dy), npy(rdy), rtol=0, atol=8e-6 * dmag)
    enc = SHEncoder(degree=4)
    x = T(dirs).requires_grad_()
    y = enc(x); assert y.shape == (1000, 16)
    gw = torch.randn_like(y); (y * gw).sum().backward()
    _, ody = oracle.sh_encode_forward(dirs, 4, True)
    np.testing.assert_allclose(npy(x.grad), oracle.sh_encode_backward(npy(gw), dirs, 4, ody), rtol=1e-4, atol=1e-5)
    fx = (np.random.default_rng(5).random((777, 6)) * 2 - 1).astype(np.float32)
    for D, deg in ((2, 8), (6, 3), (3, 10)):
        enc = FreqEncoder(input_dim=D, degree=deg)
        x = T(np.ascontiguousarray(fx[:, :D])).requires_grad_()
        y = enc(x); assert y.shape == (777, D * (1 + 2 * deg))
        oo = oracle.freq_encode_forward(fx[:, :D], deg)
        # __sinf (MUFU.SIN) absolute error grows with |argument| = 2^f |x|: 2^-21.4 * |arg| bound, documented by CUDA
        np.testing.assert_allclose(npy(y), oo, rtol=0, atol=max(5e-6, 2.0 ** (deg - 1) * 1e-6))
        gw = torch.randn_like(y); (y * gw).sum().backward()
        np.testing.assert_allclose(npy(x.grad), oracle.freq_encode_backward(npy(gw), npy(y), D, deg), rtol=1e-5, atol=1e-4)
        if ref is not None:
            ro = torch.empty_like(y)
            ref.mod("_ref_freqencoder").freq_encode_forward(x.detach(), 777, D, deg, y.shape[1], ro)
            assert torch.equal(ro, y.detach())


def test_errors_are_loud():
    """No silent fallback: wrong device / dtype / shape raises RuntimeError, like the reference's TORCH_CHECKs."""
    from gridencoder.backend import _backend as gb
    from b2nerf import B2NError
    x = torch.rand(8, 2, device="cuda"); emb = torch.rand(64, 1, device="cuda"); offs = torch.tensor([0, 64], dtype=torch.int32, device="cuda")
    out = torch.empty(1, 8, 1, device="cuda")
    with pytest.raises(RuntimeError):
        gb.grid_encode_forward(x.cpu(), emb, offs, out, 8, 2, 1, 1, 0.0, 4, None, 0, False)
    with pytest.raises(RuntimeError):
        gb.grid_encode_forward(x, emb.double(), offs, out, 8, 2, 1, 1, 0.0, 4, None, 0, False)
    with pytest.raises(RuntimeError, match="C must be"):
        gb.grid_encode_forward(x, torch.rand(64, 3, device="cuda"), offs, torch.empty(1, 8, 3, device="cuda"), 8, 2, 3, 1, 0.0, 4, None, 0, False)
    with pytest.raises(RuntimeError, match="D must be"):
        gb.grid_encode_forward(torch.rand(8, 6, device="cuda"), emb, offs, out, 8, 6, 1, 1, 0.0, 4, None, 0, False)
    assert issubclass(B2NError, RuntimeError)

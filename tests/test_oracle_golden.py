"""CPU: the oracle (oracle/oracle.c) replayed against golden vectors that were produced by RUNNING THE REFERENCE'S OWN CUDA
EXTENSIONS on a B200 (tests/golden/*.npz, generator: tests/golden/make_golden_from_ref_ext.py).  This is the parity pin of the
oracle: integer results, sample counts and sample positions must be bit-exact; float results use the tolerance written at the assert
(the only non-reproducible primitives are the GPU's ex2.approx / sin.approx special-function units)."""
import os

import numpy as np
import pytest

import cases
import oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    path = os.path.join(G, name + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"{name}.npz not generated yet")
    return np.load(path)


def gpu_scales(S, H, L):
    """Per-level scales as the B200 computes them (ex2.approx), stored by the generator; fall back to libm when absent."""
    path = os.path.join(G, "level_scales.npz")
    if os.path.exists(path):
        z = np.load(path)
        key = f"S{np.float32(S).view(np.uint32):08x}_H{H}_L{L}"
        if key in z:
            return z[key]
    return None


def test_utils_exact():
    z = load("utils")
    assert np.array_equal(oracle.morton3D(z["morton_in"]), z["morton"])
    assert np.array_equal(oracle.morton3D_invert(z["morton"]), z["morton_back"]) and np.array_equal(z["morton_back"], z["morton_in"])
    assert np.array_equal(oracle.packbits(z["pack_in"], 10.0), z["pack"])
    assert np.array_equal(oracle.morton3D_dilation(z["dil_in"], 2, 16), z["dil"])
    np.testing.assert_allclose(oracle.sph_from_ray(z["sph_o"], z["sph_d"], 2.0), z["sph"], rtol=0, atol=2e-6)      # atan2f / sqrtf: libm vs device


@pytest.mark.parametrize("name", ["head16", "cascade2", "nogamma"])
def test_march_rays_train_bit_exact(name):
    z = load("march_" + name)
    c = cases.march_case(name)
    nears, fars = oracle.near_far_from_aabb(c["rays_o"], c["rays_d"], c["aabb"], c["min_near"])
    assert np.array_equal(nears, z["nears"]) and np.array_equal(fars, z["fars"])
    n = len(nears); M = n * c["max_steps"]
    xyzs, dirs, deltas, rays, counter = oracle.march_rays_train(c["rays_o"], c["rays_d"], c["bitfield"], c["bound"], c["dt_gamma"], c["max_steps"], c["C"], c["H"],
                                                                M, nears, fars, c["noises"])
    tot = int(counter[0])
    assert np.array_equal(rays[:, 2], z["counts"]), "per-ray sample counts must be bit-exact"
    assert tot == int(z["counter"][0]) == z["xyzs"].shape[0] and int(counter[1]) == int(z["counter"][1]) == n
    assert np.array_equal(xyzs[:tot], z["xyzs"]) and np.array_equal(dirs[:tot], z["dirs"]) and np.array_equal(deltas[:tot], z["deltas"])


@pytest.mark.parametrize("name", ["head16", "cascade2", "nogamma"])
def test_march_rays_inference_bit_exact(name):
    z = load("march_infer_" + name)
    c = cases.march_case(name)
    nears, fars = oracle.near_far_from_aabb(c["rays_o"], c["rays_d"], c["aabb"], c["min_near"])
    n_alive = min(256, len(nears))
    xyzs, dirs, deltas = oracle.march_rays(n_alive, 4, np.arange(n_alive, dtype=np.int32), nears, c["rays_o"], c["rays_d"], c["bound"], c["dt_gamma"],
                                           c["max_steps"], c["C"], c["H"], c["bitfield"], nears, fars, c["noises"][:n_alive], align=128)
    assert np.array_equal(xyzs, z["xyzs"]) and np.array_equal(dirs, z["dirs"]) and np.array_equal(deltas, z["deltas"])


def _head16_segments():
    z = load("march_head16")
    counts = z["counts"]; n = len(counts)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
    return np.stack([np.arange(n, dtype=np.int32), offs, counts.astype(np.int32)], 1), z["deltas"], int(counts.sum()), n


@pytest.mark.parametrize("variant", ["plain", "sigma", "uncertainty", "triplane"])
@pytest.mark.parametrize("tag,T_thresh", [("", 1e-4), ("_T1e-1", 1e-1)])
def test_composite_train(variant, tag, T_thresh):
    z = load(f"composite_{variant}{tag}")
    rays, deltas, m, n = _head16_segments()
    f = cases.sample_fields(m, 21); g = cases.ray_grads(n, 22)
    ambs = [f["amb_aud"], f["amb_eye"]][:oracle.VARIANTS[variant][1]]
    fw = oracle.composite_rays_train_forward(variant, f["sigmas"], f["rgbs"], ambs, f["unc"], deltas, rays, T_thresh)
    # fp32 composite: 1e-5 relative (BASELINE north star); the GPU's __expf is ex2.approx (2 ulp), the oracle's is libm
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(fw["weights_sum"], z["ws"], **tol); np.testing.assert_allclose(fw["depth"], z["depth"], **tol)
    np.testing.assert_allclose(fw["image"], z["image"], **tol); np.testing.assert_allclose(fw["amb_sums"][0], z["a0"], **tol)
    if variant == "triplane":
        np.testing.assert_allclose(fw["amb_sums"][1], z["a1"], **tol)
    if oracle.VARIANTS[variant][2]:
        np.testing.assert_allclose(fw["unc_sum"], z["us"], **tol)
    # backward is evaluated with the GOLDEN forward results (what the reference's autograd saves), so only bwd arithmetic is compared
    saved = dict(weights_sum=z["ws"], amb_sums=[z["a0"]] + ([z["a1"]] if variant == "triplane" else []), unc_sum=z["us"] if "us" in z else None, image=z["image"])
    bw = oracle.composite_rays_train_backward(variant, g["g_ws"], [g["g_aud"], g["g_eye"]][:len(ambs)], g["g_unc"], g["g_img"], f["sigmas"], f["rgbs"], ambs, f["unc"],
                                              deltas, rays, saved, T_thresh)
    np.testing.assert_allclose(bw["grad_rgbs"], z["grgb"], **tol)
    np.testing.assert_allclose(bw["grad_sigmas"], z["gs"], rtol=2e-4, atol=2e-5)      # differences of O(1) running sums scaled by delta
    np.testing.assert_allclose(bw["grad_ambs"][0], z["ga0"], **tol)
    if variant == "triplane":
        np.testing.assert_allclose(bw["grad_ambs"][1], z["ga1"], **tol)
    if oracle.VARIANTS[variant][2]:
        np.testing.assert_allclose(bw["grad_unc"], z["gu"], **tol)
    # early termination happened for a visible fraction of rays at the loose threshold, and zero grads follow the stop
    assert (z["gs"] == 0).sum() >= (bw["grad_sigmas"] == 0).sum() * 0.99


@pytest.mark.parametrize("variant", ["rgb", "plain", "sigma", "uncertainty", "triplane"])
def test_composite_inference_loop(variant):
    z = load("composite_infer_" + variant)
    c = cases.march_case("head16")
    nears, fars = oracle.near_far_from_aabb(c["rays_o"], c["rays_d"], c["aabb"], c["min_near"])
    mode, n_amb, has_unc = oracle.VARIANTS[variant]
    N0 = 512
    alive = np.arange(N0, dtype=np.int32); rays_t = nears.copy()
    ws, dep, img = np.zeros(2 * N0, np.float32), np.zeros(2 * N0, np.float32), np.zeros((2 * N0, 3), np.float32)
    sums = [np.zeros(2 * N0, np.float32) for _ in range(3)]
    for it, n_step in enumerate((1, 2, 4)):
        n_alive = len(alive)
        _, _, ol = oracle.march_rays(n_alive, n_step, alive, rays_t, c["rays_o"], c["rays_d"], c["bound"], c["dt_gamma"], c["max_steps"], c["C"], c["H"], c["bitfield"],
                                     nears, fars, np.zeros(n_alive, np.float32), align=128)
        f = cases.sample_fields(ol.shape[0], 300 + it, scale_sigma=30.0)
        oracle.composite_rays(variant, n_alive, n_step, 1e-2, alive, rays_t, f["sigmas"], f["rgbs"], ol, [f["amb_aud"], f["amb_eye"]][:n_amb], f["unc"],
                              ws, dep, img, sums[:n_amb], sums[2])
        assert np.array_equal(alive, z[f"alive{it}"]), "terminated-ray flags must be exact"
        assert np.array_equal(rays_t, z[f"rays_t{it}"])
        alive = alive[alive >= 0]
    tol = dict(rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ws, z["ws"], **tol); np.testing.assert_allclose(dep, z["depth"], **tol); np.testing.assert_allclose(img, z["image"], **tol)
    if n_amb >= 1:
        np.testing.assert_allclose(sums[0], z["s0"], **tol)
    if n_amb >= 2:
        np.testing.assert_allclose(sums[1], z["s1"], **tol)
    if has_unc:
        np.testing.assert_allclose(sums[2], z["su"], **tol)


@pytest.mark.parametrize("name", list(cases.GRID_CASES))
def test_grid_encode(name):
    z = load("grid_" + name)
    c = cases.grid_case(name)
    sc = gpu_scales(c["S"], c["H"], c["L"])
    oracle.set_level_scales(sc)
    try:
        out, dy = oracle.grid_encode_forward(c["inputs"], c["embeddings"], c["offsets"], c["S"], c["H"], c["gridtype"], c["align_corners"], True, c["half"])
        if not c["half"]:
            ge, gi = oracle.grid_encode_backward(c["grad"], c["inputs"], c["offsets"], c["C"], c["S"], c["H"], c["gridtype"], c["align_corners"], z["dy_dx"])
    finally:
        oracle.set_level_scales(None)
    if sc is not None and not c["half"]:
        assert np.array_equal(out, z["outputs"]), "with the GPU's level scales the fp32 encoding is bit-exact"
        assert np.array_equal(dy, z["dy_dx"])
    else:
        top = 2.0 ** (c["S"] * (c["L"] - 1)) * c["H"]
        np.testing.assert_allclose(out.astype(np.float32), z["outputs"].astype(np.float32), rtol=0, atol=4e-3 if c["half"] else 1e-4)
        # libm exp2f vs ex2.approx moves the level scale by an ulp => the cell-relative position by ~scale*6e-8 => d(out)/d(in) ~ scale*|dtable|
        np.testing.assert_allclose(dy.astype(np.float32), z["dy_dx"].astype(np.float32), rtol=2e-3 if not c["half"] else 2e-2, atol=(16.0 if c["half"] else 2e-2) * top / 64)
    if not c["half"]:
        np.testing.assert_allclose(ge, z["grad_embeddings"], rtol=1e-4, atol=1e-4)       # reference: float atomics in arbitrary order
        np.testing.assert_allclose(gi, z["grad_inputs"], rtol=1e-5, atol=1e-4)


def test_sh_all_degrees():
    z = load("sh")
    dirs = cases.dirs_case(256, 3)
    raw = (np.random.default_rng(4).standard_normal((64, 3)) * 0.7).astype(np.float32)
    for deg in range(1, 9):
        for tag, v in (("unit", dirs), ("raw", raw)):
            out, dy = oracle.sh_encode_forward(v, deg, True)
            mag = max(1.0, float(np.abs(z[f"{tag}_out{deg}"]).max())); dmag = max(1.0, float(np.abs(z[f"{tag}_dy{deg}"]).max()))
            # the reference evaluates the expanded polynomials in fp32; the oracle the same functions by recurrence in double
            np.testing.assert_allclose(out, z[f"{tag}_out{deg}"], rtol=0, atol=1e-5 * mag)
            np.testing.assert_allclose(dy, z[f"{tag}_dy{deg}"], rtol=0, atol=1e-5 * dmag)
            if deg in (4, 8):
                gr = np.random.default_rng(deg).standard_normal((v.shape[0], deg * deg)).astype(np.float32)
                gi = oracle.sh_encode_backward(gr, v, deg, z[f"{tag}_dy{deg}"])
                np.testing.assert_allclose(gi, z[f"{tag}_gin{deg}"], rtol=1e-5, atol=1e-5 * dmag)


def test_freq():
    z = load("freq")
    fx = (np.random.default_rng(5).random((256, 6)) * 2 - 1).astype(np.float32)
    for D, deg in ((2, 8), (6, 3), (3, 10)):
        v = np.ascontiguousarray(fx[:, :D])
        out = oracle.freq_encode_forward(v, deg)
        # reference: __sinf (sin.approx) with |argument| up to 2^(deg-1): absolute error ~ 2^-21.4 * |arg|
        np.testing.assert_allclose(out, z[f"out_{D}_{deg}"], rtol=0, atol=max(5e-6, 2.0 ** (deg - 1) * 1e-6))
        gr = np.random.default_rng(D).standard_normal(out.shape).astype(np.float32)
        np.testing.assert_allclose(oracle.freq_encode_backward(gr, z[f"out_{D}_{deg}"], D, deg), z[f"gin_{D}_{deg}"], rtol=1e-5, atol=1e-4)

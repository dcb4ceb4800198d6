"""CPU: the oracle against the integer known-answer vectors derived from the reference formulas and its own
invariants.  These need neither a GPU nor /root/reference."""
import numpy as np

import cases
import oracle
from b2nerf import scene


def test_morton_known_answers():
    ka = [[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1], [2, 0, 0], [3, 5, 7], [64, 64, 64], [127, 0, 0], [0, 127, 0], [0, 0, 127],
          [100, 37, 90], [127, 127, 127]]
    want = [0, 1, 2, 4, 7, 8, 431, 1835008, 299593, 599186, 1198372, 1427682, 2097151]
    assert oracle.morton3D(ka).tolist() == want                                  # raymarching.cu:56-71
    assert oracle.morton3D_invert(want).tolist() == ka                            # raymarching.cu:73-81
    rng = np.random.default_rng(0)
    c = rng.integers(0, 1024, (50000, 3)).astype(np.int32)
    assert np.array_equal(oracle.morton3D_invert(oracle.morton3D(c)), c)
    assert np.array_equal(oracle.morton3D(c), scene.morton3d(c[:, 0], c[:, 1], c[:, 2]).astype(np.int32))


def test_packbits_known_answer_and_scene():
    assert oracle.packbits([0, 11, 10, 10.0001, 9, 50, -1, 10], 10.0).tolist() == [42]      # raymarching.cu:283-288, strict >
    g = scene.density_grid()
    bf = oracle.packbits(g, scene.DENSITY_THRESH)
    assert bf.shape == (262144,) and np.array_equal(bf, scene.bitfield_from_grid(g))
    occ = np.unpackbits(bf).mean()
    assert 0.034 < occ < 0.039          # solid ball r ~ 0.41 => ~3.6 % of voxels


def test_dilation_is_6_neighbour_max():
    rng = np.random.default_rng(1)
    H = 8
    g = rng.standard_normal((1, H ** 3)).astype(np.float32)
    out = oracle.morton3D_dilation(g, 1, H)
    x, y, z = np.meshgrid(np.arange(H), np.arange(H), np.arange(H), indexing="ij")
    dense = np.zeros((H, H, H), np.float32); dense[x, y, z] = g[0, scene.morton3d(x, y, z)]
    pad = np.pad(dense, 1, constant_values=-np.inf)
    want = np.maximum.reduce([pad[1:-1, 1:-1, 1:-1], pad[2:, 1:-1, 1:-1], pad[:-2, 1:-1, 1:-1], pad[1:-1, 2:, 1:-1], pad[1:-1, :-2, 1:-1],
                              pad[1:-1, 1:-1, 2:], pad[1:-1, 1:-1, :-2]])
    assert np.array_equal(out[0, scene.morton3d(x, y, z)], want)


def test_grid_index_and_scale_known_answers():
    from gridencoder.grid import level_table
    pls = np.exp2(np.log2(512 / 64) / 11)
    offs = level_table(2, 12, pls, 64, 14, False)
    assert offs == [0, 4232, 10480, 19512, 32512, 48896, 65280, 81664, 98048, 114432, 130816, 147200, 163584]      # grid.py:111-123
    S = np.float32(np.log2(pls))
    assert S.view(np.uint32) == 0x3E8BA2E9
    sc = oracle.level_scales(S, 64, 12)
    np.testing.assert_allclose(sc, [63.0, 76.3177, 92.4067, 111.8437, 135.3253, 163.6931, 197.9640, 239.3663, 289.3840, 349.8099, 422.8097, 511.0], rtol=2e-6)
    assert sc[0] == 63.0 and sc[11] == 511.0
    # the index function itself (gridencoder.cu:54-72): dense level 0 (resolution 64, 4232 slots, stride 65), hashed level 11
    assert oracle.grid_index([3, 5], 4232, 64) == 328 and oracle.grid_index([64, 64], 4232, 64) == 4224
    for (gx, gy), want in {(0, 0): 0, (1, 0): 1, (0, 1): 14769, (10, 20): 478, (11, 20): 479, (136, 137): 7985, (511, 512): 9215, (512, 512): 8192}.items():
        assert oracle.grid_index([gx, gy], 16384, 512) == want, (gx, gy)
    assert oracle.grid_index([3, 5], 4232, 64, gridtype=1) == 328 and oracle.grid_index([200, 300], 16384, 512, gridtype=1) == (200 + 300 * 513) % 16384


def test_march_constants_and_invariants():
    c = cases.march_case("head16")
    nears, fars = oracle.near_far_from_aabb(c["rays_o"], c["rays_d"], c["aabb"], c["min_near"])
    n = len(nears); M = n * 16
    xyzs, dirs, deltas, rays, counter = oracle.march_rays_train(c["rays_o"], c["rays_d"], c["bitfield"], 1.0, c["dt_gamma"], 16, 1, 128, M, nears, fars, c["noises"])
    tot = int(counter[0])
    assert counter[1] == n and tot == rays[:, 2].sum() and rays[:, 2].max() <= 16
    assert np.array_equal(rays[:, 1], np.concatenate([[0], np.cumsum(rays[:, 2])[:-1]]))
    dt = np.float32(2 * np.sqrt(np.float32(3)) / 128)
    assert np.all(deltas[:tot, 0].view(np.uint32) == 0x3CDDB3D7)                 # dt_min = dt_max = 0.027063293
    assert np.all(deltas[tot:] == 0) and np.all(xyzs[tot:] == 0)
    # every sample sits in an occupied voxel
    v = np.clip((0.5 * (xyzs[:tot].astype(np.float64) + 1) * 128), 0, 127).astype(np.int64)
    m = scene.morton3d(v[:, 0], v[:, 1], v[:, 2])
    assert np.all((c["bitfield"][m // 8] >> (m % 8)) & 1)
    assert rays[2, 2] == 0                                                         # the ray that misses the box
    # inference march of all rays for 16 steps reproduces the training samples ray by ray (same DDA)
    ix, idr, idl = oracle.march_rays(n, 16, np.arange(n, dtype=np.int32), nears, c["rays_o"], c["rays_d"], 1.0, c["dt_gamma"], 16, 1, 128, c["bitfield"],
                                     nears, fars, c["noises"])
    for r in (0, 1, 5, 100, n - 1):
        o, k = rays[r, 1], rays[r, 2]
        assert np.array_equal(ix[r * 16:r * 16 + k], xyzs[o:o + k]) and np.all(idl[r * 16 + k:(r + 1) * 16] == 0)


def test_composite_matches_closed_form_and_gradients():
    """Oracle forward against a float64 numpy restatement; backward against central finite differences in float64."""
    rng = np.random.default_rng(3)
    n = 40; counts = rng.integers(1, 17, n)
    offs = np.concatenate([[0], np.cumsum(counts)[:-1]])
    rays = np.stack([np.arange(n), offs, counts], 1).astype(np.int32); m = int(counts.sum())
    f = cases.sample_fields(m, 1, scale_sigma=1.0)
    dl = np.stack([np.full(m, 0.027, np.float32), rng.random(m).astype(np.float32) + 2], 1)

    def fwd64(sig, rgb, unc):
        out = np.zeros((n, 6))
        for i in range(n):
            T = 1.0
            for k in range(offs[i], offs[i] + counts[i]):
                a = 1 - np.exp(-sig[k] * dl[k, 0]); w = a * T
                out[i, 0] += w; out[i, 1:4] += w * rgb[k]; out[i, 4] += w * dl[k, 1]; out[i, 5] += w * unc[k]
                T *= 1 - a
                if T < 1e-4:
                    break
        return out
    o = oracle.composite_rays_train_forward("triplane", f["sigmas"], f["rgbs"], [f["amb_aud"], f["amb_eye"]], f["unc"], dl, rays)
    ref = fwd64(f["sigmas"].astype(np.float64), f["rgbs"].astype(np.float64), f["unc"].astype(np.float64))
    np.testing.assert_allclose(o["weights_sum"], ref[:, 0], rtol=2e-6); np.testing.assert_allclose(o["image"], ref[:, 1:4], rtol=2e-6)
    np.testing.assert_allclose(o["depth"], ref[:, 4], rtol=2e-6); np.testing.assert_allclose(o["unc_sum"], ref[:, 5], rtol=2e-6)
    np.testing.assert_allclose(o["amb_sums"][0], [f["amb_aud"][a:a + c].sum() for a, c in zip(offs, counts)], rtol=2e-6)
    g = cases.ray_grads(n, 2)
    bw = oracle.composite_rays_train_backward("triplane", g["g_ws"], [g["g_aud"], g["g_eye"]], g["g_unc"], g["g_img"], f["sigmas"], f["rgbs"],
                                              [f["amb_aud"], f["amb_eye"]], f["unc"], dl, rays, o)

    def loss(sig):
        r = fwd64(sig, f["rgbs"].astype(np.float64), f["unc"].astype(np.float64))
        return (r[:, 0] * g["g_ws"]).sum() + (r[:, 1:4] * g["g_img"]).sum() + (r[:, 5] * g["g_unc"]).sum()
    s0 = f["sigmas"].astype(np.float64)
    for k in rng.integers(0, m, 25):
        e = np.zeros(m); e[k] = 1e-5 * max(1.0, s0[k])
        fd = (loss(s0 + e) - loss(s0 - e)) / (2 * e[k])
        assert abs(fd - bw["grad_sigmas"][k]) <= 2e-3 * max(1.0, abs(fd)), (k, fd, bw["grad_sigmas"][k])
    np.testing.assert_allclose(bw["grad_ambs"][0], np.repeat(g["g_aud"], counts), rtol=0, atol=0)


def test_sh_matches_explicit_low_orders_and_unit_norm_sum():
    d = cases.dirs_case(64)
    out, dy = oracle.sh_encode_forward(d, 4, True)
    x, y, z = d[:, 0].astype(np.float64), d[:, 1].astype(np.float64), d[:, 2].astype(np.float64)
    want = np.stack([0.28209479177387814 + 0 * x, -0.48860251190291987 * y, 0.48860251190291987 * z, -0.48860251190291987 * x,
                     1.0925484305920792 * x * y, -1.0925484305920792 * y * z, 0.94617469575755997 * z * z - 0.31539156525251999,
                     -1.0925484305920792 * x * z, 0.54627421529603959 * (x * x - y * y),
                     0.59004358992664352 * y * (-3 * x * x + y * y), 2.8906114426405538 * x * y * z, 0.45704579946446572 * y * (1 - 5 * z * z),
                     0.3731763325901154 * z * (5 * z * z - 3), 0.45704579946446572 * x * (1 - 5 * z * z), 1.4453057213202769 * z * (x * x - y * y),
                     0.59004358992664352 * x * (-x * x + 3 * y * y)], 1)          # shencoder.cu:44-67 (standard real SH, deg <= 4)
    np.testing.assert_allclose(out, want, rtol=0, atol=3e-7)
    o8, _ = oracle.sh_encode_forward(d, 8, False)
    for l in range(8):                                                             # addition theorem on the unit sphere
        np.testing.assert_allclose((o8[:, l * l:(l + 1) ** 2] ** 2).sum(1), (2 * l + 1) / (4 * np.pi), rtol=2e-5)
    eps = 1e-3                                                                     # partials vs central differences
    for k in range(3):
        e = np.zeros(3, np.float32); e[k] = eps
        fd = (oracle.sh_encode_forward(d + e, 8)[0].astype(np.float64) - oracle.sh_encode_forward(d - e, 8)[0]) / (2 * eps)
        dy8 = oracle.sh_encode_forward(d, 8, True)[1].reshape(64, 3, 64)[:, k]
        np.testing.assert_allclose(dy8, fd, rtol=0, atol=2e-3 * max(1.0, np.abs(fd).max()))


def test_freq_layout_and_backward():
    x = (np.random.default_rng(0).random((9, 2)) * 2 - 1).astype(np.float32)
    out = oracle.freq_encode_forward(x, 3)
    assert out.shape == (9, 14)
    np.testing.assert_allclose(out[:, :2], x)
    for f in range(3):                                                             # [x, sin(2^f x), cos(2^f x)] blocks (freqencoder.cu:48-57)
        np.testing.assert_allclose(out[:, 2 + 4 * f:4 + 4 * f], np.sin(2.0 ** f * x), atol=2e-6)
        np.testing.assert_allclose(out[:, 4 + 4 * f:6 + 4 * f], np.cos(2.0 ** f * x), atol=2e-6)
    g = np.random.default_rng(1).standard_normal(out.shape).astype(np.float32)
    want = g[:, :2] + sum(2.0 ** f * (g[:, 2 + 4 * f:4 + 4 * f] * np.cos(2.0 ** f * x) - g[:, 4 + 4 * f:6 + 4 * f] * np.sin(2.0 ** f * x)) for f in range(3))
    np.testing.assert_allclose(oracle.freq_encode_backward(g, out, 2, 3), want, rtol=1e-5, atol=1e-5)

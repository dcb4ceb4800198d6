#!/usr/bin/env python
"""Per-kernel roofline measurements of the per-op (drop-in) kernels at sizes LARGER THAN L2, so HBM-bound kernels are measured
against HBM and not against the 126 MB L2.  Run on a B200:   python profiles/kernel_rooflines.py > gpurun_out/kernel_rooflines.json

achieved = algorithmic bytes per launch (SURVEY §8d per-unit figures x units) / average launch duration (CUDA events on the
launching stream, 3 warm-ups, 10 timed launches back to back);  frac = achieved / MEASURED_PEAKS.json:hbm_gbs.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lzzx-nerf_b200"), ROOT):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import raymarching  # noqa: E402
from raymarching.backend import _backend as rb  # noqa: E402
from gridencoder.backend import _backend as gb  # noqa: E402
from b2nerf import scene  # noqa: E402

dev = torch.device("cuda")
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    SRC = "measured"
except Exception:
    HBM, SRC = 6650.0, "fallback"


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def measure():
    """Runs every measurement; returns {hbm_peak_gbs, peak_source, kernels: [...]}."""
    rows = []


    def report(name, alg_bytes, sec, note=""):
        gbs = alg_bytes / sec / 1e9
        rows.append({"kernel": name, "alg_bytes": int(alg_bytes), "us": sec * 1e6, "achieved_gbs": gbs, "frac_of_hbm": gbs / HBM, "note": note})


    torch.manual_seed(0)
    # ---- composites: 4 M rays, 0..16 samples each (mean 8) -> 33 M samples, 1.2 GB of per-sample inputs -----------------------------
    N = 4 * 1024 * 1024
    counts = torch.randint(0, 17, (N,), device=dev, dtype=torch.int32)
    offs = torch.cumsum(counts, 0, dtype=torch.int32) - counts
    rays = torch.stack([torch.arange(N, device=dev, dtype=torch.int32), offs, counts], 1).contiguous()
    M = int(counts.sum())
    sig = torch.rand(M, device=dev) * 5; rgb = torch.rand(M, 3, device=dev); aud = torch.rand(M, device=dev); eye = torch.rand(M, device=dev)
    unc = torch.rand(M, device=dev); dl = torch.rand(M, 2, device=dev) * 0.03 + 0.01
    ws, a0, a1, us, dep = (torch.empty(N, device=dev) for _ in range(5)); img = torch.empty(N, 3, device=dev)
    f = lambda: rb.composite_rays_train_triplane_forward(sig, rgb, aud, eye, unc, dl, rays, M, N, 1e-4, ws, a0, a1, us, dep, img)
    report("composite_rays_train_triplane_forward", 36 * M + 44 * N, timeit(f), f"N={N} rays, M={M} samples (tiled layout -> smem-staged path)")
    g_ws, g_a0, g_a1, g_u = (torch.randn(N, device=dev) for _ in range(4)); g_img = torch.randn(N, 3, device=dev)
    gs, ga0, ga1, gu = (torch.zeros(M, device=dev) for _ in range(4)); grgb = torch.zeros(M, 3, device=dev)
    f = lambda: rb.composite_rays_train_triplane_backward(g_ws, g_a0, g_a1, g_u, g_img, sig, rgb, aud, eye, unc, dl, rays, ws, a0, a1, us, img, M, N, 1e-4, gs, grgb, ga0, ga1, gu)
    report("composite_rays_train_triplane_backward", (36 + 28) * M + (12 + 32 + 28) * N, timeit(f), "same rays")
    # foreign row orders.  The reference allocates segments with an atomicAdd per ray (raymarching.cu:446-454), so its rows come out in the order in which warps
    # reach the atomic: scrambled within a neighbourhood of a few thousand rays, not over the whole batch — "window 4096" models that; the global permutation
    # is the adversarial case (every segment read and every per-ray output write is a random DRAM access: not an HBM-streaming workload any more)
    win = (torch.arange(N, device=dev) // 4096) * 4096
    local = (win.float() + torch.rand(N, device=dev) * 4096).argsort()
    rays_loc = rays[local].contiguous()
    f = lambda: rb.composite_rays_train_triplane_forward(sig, rgb, aud, eye, unc, dl, rays_loc, M, N, 1e-4, ws, a0, a1, us, dep, img)
    report("composite_rays_train_triplane_forward[rows scrambled within windows of 4096: the reference's atomic order]", 36 * M + 44 * N, timeit(f),
           "per-warp gathered staging (segments copied ray by ray into shared memory)")
    perm = torch.randperm(N, device=dev)
    rays_sh = rays[perm].contiguous()
    f = lambda: rb.composite_rays_train_triplane_forward(sig, rgb, aud, eye, unc, dl, rays_sh, M, N, 1e-4, ws, a0, a1, us, dep, img)
    report("composite_rays_train_triplane_forward[rows globally permuted: adversarial]", 36 * M + 44 * N, timeit(f), "random DRAM accesses per segment and per output row")
    del rays_loc, local, win
    del gs, ga0, ga1, gu, grgb
    # inference composite: 8 M alive rays x 4 steps
    na, ns = 8 * 1024 * 1024, 4
    Mi = na * ns
    alive = torch.arange(na, device=dev, dtype=torch.int32); rays_t = torch.rand(na, device=dev)
    acc = [torch.zeros(na, device=dev) for _ in range(5)]; imgi = torch.zeros(na, 3, device=dev)
    sig_i, rgb_i, dl_i = sig[:Mi], rgb[:Mi], dl[:Mi]
    f = lambda: rb.composite_rays_triplane(na, ns, 1e-4, alive.clone(), rays_t, sig_i, rgb_i, dl_i, aud[:Mi], eye[:Mi], unc[:Mi], acc[0], acc[1], imgi, acc[2], acc[3], acc[4])
    t_clone = timeit(lambda: alive.clone())
    report("composite_rays_triplane", 36 * Mi + (8 + 64) * na, timeit(f) - t_clone, f"n_alive={na}, n_step={ns}")
    del sig, rgb, aud, eye, unc, dl, rays, rays_sh, perm, acc, imgi
    torch.cuda.empty_cache()

    # ---- utils ---------------------------------------------------------------------------------------------------------------------------
    Nr = 32 * 1024 * 1024
    o = torch.randn(Nr, 3, device=dev) * 0.1 + torch.tensor([0.0, 0.0, 3.0], device=dev); d = torch.nn.functional.normalize(torch.randn(Nr, 3, device=dev) * 0.1 + torch.tensor([0.0, 0.0, -1.0], device=dev), dim=1)
    aabb = torch.from_numpy(scene.AABB).to(dev); nears, fars = torch.empty(Nr, device=dev), torch.empty(Nr, device=dev)
    report("near_far_from_aabb", 32 * Nr, timeit(lambda: rb.near_far_from_aabb(o, d, aabb, Nr, 0.05, nears, fars)), f"N={Nr}")
    coords = torch.randint(0, 128, (Nr, 3), device=dev, dtype=torch.int32); idx = torch.empty(Nr, device=dev, dtype=torch.int32)
    report("morton3D", 16 * Nr, timeit(lambda: rb.morton3D(coords, Nr, idx)), f"N={Nr}")
    grid = torch.rand(8, 256 ** 3, device=dev) * 20; bits = torch.empty(8 * 256 ** 3 // 8, device=dev, dtype=torch.uint8)
    report("packbits", 33 * bits.numel(), timeit(lambda: rb.packbits(grid, bits.numel(), 10.0, bits)), "8 cascades x 256^3 cells (537 MB)")
    gd = torch.empty_like(grid)
    report("morton3D_dilation", 8 * grid.numel(), timeit(lambda: rb.morton3D_dilation(grid, 8, 256, gd)), "read 4 B + write 4 B per cell (neighbours from cache)")
    del grid, gd, bits, coords, idx
    torch.cuda.empty_cache()

    # ---- march_rays_train: 1 M rays of the synthetic head scene ------------------------------------------------------------------------------
    bf = torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).to(dev)
    Nm = 1024 * 1024
    oo, dd = [], []
    for s in range(16):
        a, b = scene.train_rays(s, 65536)
        oo.append(torch.from_numpy(a)); dd.append(torch.from_numpy(b))
    ro, rd = torch.cat(oo).to(dev), torch.cat(dd).to(dev)
    nears, fars = raymarching.near_far_from_aabb(ro, rd, aabb, 0.05)
    Mm = Nm * 6
    xyzs, dirs, dls = torch.zeros(Mm, 3, device=dev), torch.zeros(Mm, 3, device=dev), torch.zeros(Mm, 2, device=dev)
    rr = torch.empty(Nm, 3, device=dev, dtype=torch.int32); counter = torch.zeros(2, device=dev, dtype=torch.int32); noises = torch.rand(Nm, device=dev)


    def march():
        counter.zero_()
        rb.march_rays_train(ro, rd, bf, 1.0, 1 / 256, 16, Nm, 1, 128, Mm, nears, fars, xyzs, dirs, dls, rr, counter, noises)


    sec = timeit(march)
    tot = int(counter[0])
    report("march_rays_train (count + write)", 52 * Nm + 32 * tot, sec, f"N={Nm} rays -> {tot} samples; includes the two DDA passes; {Nm / sec / 1e6:.1f} M rays/s")

    # ---- ceilings for the grid encoder: random L2-resident gathers / reductions (csrc/peaks.cu), table = one tri-plane table rounded up to a power of two ------
    from b2nerf import lib
    L_ = lib()
    st = torch.cuda.current_stream().cuda_stream
    ntab = 262144                                          # 1 MB of floats (the 163 584-entry table is 654 KB): L2-resident, 4x an SM's L1
    probe_tab, sink = torch.rand(ntab, device=dev), torch.zeros(4, device=dev)
    blocks, per_thread = 148 * 8, 512
    n_acc = blocks * 256 * per_thread
    peaks_l2 = {}
    for vec in (1, 2, 4):
        sec = timeit(lambda: L_.call("b2n_probe_gather", probe_tab.data_ptr(), ntab, vec, per_thread, blocks, sink.data_ptr(), st))
        peaks_l2[f"gather_{4 * vec}B_per_s"] = n_acc / sec
    sec = timeit(lambda: L_.call("b2n_probe_red", probe_tab.data_ptr(), ntab, per_thread, blocks, st))
    peaks_l2["red_f32_per_s"] = n_acc / sec
    peaks_l2["how"] = f"{blocks} x 256 threads x {per_thread} independent accesses at pseudo-random positions of a {ntab * 4 // 1024} KB table (csrc/peaks.cu), CUDA events"

    # ---- grid encode (API kernels), tri-plane config, 8 M points per plane --------------------------------------------------------------------
    from gridencoder import GridEncoder  # noqa: E402
    from gridencoder.backend import grid_encode_forward_rows  # noqa: E402
    enc = GridEncoder(input_dim=2, num_levels=12, level_dim=1, base_resolution=64, log2_hashmap_size=14, desired_resolution=512).to(dev)
    enc.embeddings.data.uniform_(-1, 1)
    B = 8 * 1024 * 1024
    x = torch.rand(B, 2, device=dev); out = torch.empty(12, B, 1, device=dev)
    S = float(np.log2(enc.per_level_scale))
    f = lambda: gb.grid_encode_forward(x, enc.embeddings.data, enc.offsets, out, B, 2, 1, 12, S, 64, None, 0, False)
    sec = timeit(f)
    # corner reads per point: 12 levels x 2 rows x (one 8-byte pair when the x-neighbours are adjacent, else two 4-byte reads) -> >= 24 gathers
    report("grid_encode_forward (one plane, D=2 L=12 C=1 fp32, reference layout [L,B,C])", (8 + 48) * B, sec,
           f"B={B} random points; {24 * B / sec / 1e9:.1f} G paired corner gathers/s = {24 * B / sec / peaks_l2['gather_8B_per_s']:.2f} of the measured random 8-byte gather rate "
           f"({peaks_l2['gather_8B_per_s'] / 1e9:.1f} G/s); the caller still pays grid.py:52's transposing copy")
    rows[-1]["frac_of_l2_gather_peak"] = 24 * B / sec / peaks_l2["gather_8B_per_s"]
    out_rows = torch.empty(B, 12, device=dev)
    f = lambda: grid_encode_forward_rows(x, enc.embeddings.data, enc.offsets, out_rows, B, 2, 1, 12, S, 64, 0, False)
    sec_r = timeit(f)
    perm = timeit(lambda: out.permute(1, 0, 2).reshape(B, 12))
    report("grid_encode_forward_rows (same, row-major [B, L*C]: what GridEncoder.forward returns)", (8 + 48) * B, sec_r,
           f"{24 * B / sec_r / 1e9:.1f} G paired corner gathers/s = {24 * B / sec_r / peaks_l2['gather_8B_per_s']:.2f} of the gather ceiling; replaces the level-major kernel "
           f"({sec * 1e6:.0f} us) + the transposing copy ({perm * 1e6:.0f} us)")
    rows[-1]["frac_of_l2_gather_peak"] = 24 * B / sec_r / peaks_l2["gather_8B_per_s"]
    assert torch.equal(out_rows, out.permute(1, 0, 2).reshape(B, 12))
    del out_rows
    grad = torch.randn(12, B, 1, device=dev); ge = torch.zeros_like(enc.embeddings.data)
    f = lambda: gb.grid_encode_backward(grad, x, enc.embeddings.data, enc.offsets, ge, B, 2, 1, 12, S, 64, None, None, 0, False)
    sec = timeit(f)
    report("grid_encode_backward (one plane)", (8 + 48) * B, sec,
           f"B={B}; {48 * B / sec / 1e9:.1f} G table reductions/s (48 per point; privatised in shared memory, flushed with red.v4) vs {peaks_l2['red_f32_per_s'] / 1e9:.1f} G/s "
           f"of direct random red.f32 into an L2-resident table = {48 * B / sec / peaks_l2['red_f32_per_s']:.2f}x")
    rows[-1]["vs_l2_red_peak"] = 48 * B / sec / peaks_l2["red_f32_per_s"]

    return {"hbm_peak_gbs": HBM, "peak_source": SRC, "l2_peaks": peaks_l2, "kernels": rows}


if __name__ == "__main__":
    print(json.dumps(measure(), indent=1))

"""GPU: the fused tcgen05 head kernel against the reference graph (network.py:252-311) evaluated op by op with torch under
autocast(fp16) — which is how the reference runs it — and against the same graph in fp32."""
import numpy as np
import pytest
import torch

from b2nerf import scene

pytestmark = pytest.mark.gpu


def _model(seed=0, table_scale=1.0, testing=True):
    from b2nerf.model import HeadModel
    torch.manual_seed(seed)
    m = HeadModel().cuda()
    for enc in (m.encoder_xy, m.encoder_yz, m.encoder_xz):
        enc.embeddings.data.uniform_(-table_scale, table_scale)
    m.testing = testing
    return m


def _samples(n, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.rand(n, 3, device="cuda", generator=g) * 2 - 1
    d = torch.randn(n, 3, device="cuda", generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    return x, d


def _stats(a, b):
    diff = (a.float() - b.float()).abs()
    return float(diff.max()), float(diff.median())


@pytest.mark.parametrize("testing", [True, False])
@pytest.mark.parametrize("n", [128 * 5, 100003])
def test_fused_head_matches_autocast_reference(testing, n):
    m = _model(0, 1.0, testing)
    x, d = _samples(n)
    x[0] = torch.tensor([1.0, -1.0, 1.0]); x[1] = 0.0; x[2] = torch.tensor([-1.0, 0.3, 0.999])
    enc_a = torch.randn(1, 32, device="cuda") * 0.5
    c = m.individual_codes[0:1].detach()
    e = torch.tensor([[0.37]], device="cuda")
    m.pack()
    sig, rgb, aud, eye, unc = m(x, d, enc_a, c, e)
    with torch.no_grad():
        with torch.autocast("cuda", dtype=torch.float16):
            r_sig, r_rgb, r_aud, r_eye, r_unc = m.forward_unfused(x, d, enc_a, c, e)
        f_sig, f_rgb, f_aud, f_eye, f_unc = m.forward_unfused(x, d, enc_a, c, e)          # fp32 "truth"
    torch.cuda.synchronize()
    assert torch.isfinite(sig).all() and torch.isfinite(rgb).all()
    # rgb in [0,1]: fp16 chain -> 1e-3 (north star: "within 1e-3 (fp16 ...)"); allow a few ulp flips at the tails
    mx, med = _stats(rgb, r_rgb)
    assert med < 3e-4 and mx < 8e-3, (mx, med)
    # log-density: |h0_fused - h0_ref| is a few fp16 ulps
    dl = (sig.log() - r_sig.float().log()).abs()
    assert float(dl.median()) < 1e-3 and float(dl.max()) < 3e-2, (float(dl.max()), float(dl.median()))
    mx, med = _stats(aud, r_aud); assert med < 1e-3 * max(1.0, float(r_aud.float().abs().median())) and mx < 3e-2 * max(1.0, float(r_aud.float().abs().max())), (mx, med)
    mx, med = _stats(eye, r_eye); assert mx < 4e-3, (mx, med)
    if testing:
        assert torch.allclose(unc.view(-1), torch.full((n,), float(np.log(2.0)), device="cuda"))
        assert r_unc.shape == (n, 36, 1)                      # the reference's shape quirk when testing (network.py:245)
    else:
        mx, med = _stats(unc.view(-1), r_unc.view(-1)); assert mx < 2e-2 and med < 1e-3, (mx, med)
    # the fused path is as close to the fp32 graph as the reference's own autocast path is (within 2x)
    err_fused = float((rgb - f_rgb).abs().mean()); err_ref = float((r_rgb.float() - f_rgb).abs().mean())
    assert err_fused < 2.0 * err_ref + 1e-5, (err_fused, err_ref)


def test_fused_head_n_valid_and_reference_init():
    """Reference init (tables +-1e-4): features ~1e-4, outputs near-constant but must still agree; n_valid limits the work."""
    m = _model(1, 1e-4, True)
    x, d = _samples(4096, 5)
    enc_a = torch.randn(1, 32, device="cuda"); c = m.individual_codes[3:4].detach(); e = torch.tensor([[0.8]], device="cuda")
    m.pack()
    nv = torch.tensor([1000], dtype=torch.int32, device="cuda")
    sig = m(x, d, enc_a, c, e)[0].clone()
    out = torch.full((4096,), -7.0, device="cuda")
    from b2nerf import lib
    keep = (enc_a.float().contiguous(), c.float().contiguous().view(-1), e.float().contiguous().view(-1))
    lib().call("b2n_head_forward", m.handle, x.data_ptr(), d.data_ptr(), 4096, keep[0].data_ptr(), keep[1].data_ptr(), keep[2].data_ptr(), nv.data_ptr(),
               out.data_ptr(), None, None, None, None, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(out[:1000], sig[:1000]) and bool((out[1000:] == -7.0).all())
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        r_sig = m.forward_unfused(x, d, enc_a, c, e)[0]
    assert float((sig.log() - r_sig.float().log()).abs().max()) < 5e-3


def test_fused_gather_equals_grid_encoder_bits():
    """The fused kernel's tri-plane gather uses the same arithmetic as the grid_encode kernel: with identity-like weights the
    first layer reproduces fp16(features) exactly.  Checked indirectly: zero tables except one plane => outputs depend only on it."""
    m = _model(2, 1.0, True)
    x, d = _samples(2048, 9)
    enc_a = torch.zeros(1, 32, device="cuda"); c = torch.zeros(1, 4, device="cuda"); e = torch.zeros(1, 1, device="cuda")
    m.pack()
    a = m(x, d, enc_a, c, e)[3].clone()           # eye attention depends on enc_x only
    x2 = x.clone(); x2[:, 0] = -x2[:, 0]           # change x only: planes xy and xz change
    m.encoder_xy.embeddings.data.zero_(); m.encoder_xz.embeddings.data.zero_()
    m.pack()                                       # the tables are part of the packed model (corner-quad image), like the MLP weights
    b1 = m(x, d, enc_a, c, e)[3].clone(); b2 = m(x2, d, enc_a, c, e)[3].clone()
    torch.cuda.synchronize()
    assert torch.equal(b1, b2) and not torch.equal(a, b1)       # with only the yz plane alive, x is irrelevant


@pytest.mark.parametrize("table_scale", [1.0, 1e-4])
def test_quad_gather_equals_table_gather_bits(table_scale, monkeypatch):
    """The corner-quad image (one 16-byte read per cell, hashed levels de-hashed; fused_head.cu:k_pack_quads) holds the same values the
    reference index function (gridencoder.cu:54-72) reaches, so both gather modes give identical outputs, bit for bit — including samples on
    the faces of the cube (last cell of every level) and out-of-range ones."""
    import copy
    m1 = _model(3, table_scale, False)
    monkeypatch.setenv("B2N_HEAD_QUADS", "0")
    m0 = copy.deepcopy(m1); m0._handle = None; m0._packed_weights = None
    m0.pack()                                       # model created with the environment set: gathers from the reference-format tables
    monkeypatch.delenv("B2N_HEAD_QUADS")
    m1.pack()
    x, d = _samples(200003, 11)
    x[:6] = torch.tensor([[1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], [1.0, -1.0, 0.0], [0.0, 1.0, -1.0], [1.5, 0.2, 0.1], [0.999999, 0.999999, 0.999999]], device="cuda")
    enc_a = torch.randn(1, 32, device="cuda") * 0.5; c = m1.individual_codes[2:3].detach(); e = torch.tensor([[0.6]], device="cuda")
    o1 = [t.clone() for t in m1(x, d, enc_a, c, e)]
    o0 = [t.clone() for t in m0(x, d, enc_a, c, e)]
    torch.cuda.synchronize()
    for a, b in zip(o1, o0):
        assert torch.equal(a, b)


def _loop_reference(m, rays_o, rays_d, enc_a, c, e, fused_net, max_steps=16, dt_gamma=1 / 256, T_thresh=1e-4):
    """The reference's host-driven loop (renderer.py:442-561) on the drop-in per-op kernels; network = fused or unfused."""
    import raymarching
    N = rays_o.shape[0]
    nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, m.aabb_infer, 0.05)
    ws, depth, image = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda"), torch.zeros(N, 3, device="cuda")
    s_aud, s_eye, s_unc = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
    alive = torch.arange(N, dtype=torch.int32, device="cuda"); rays_t = nears.clone()
    step, trace = 0, []
    while step < max_steps:
        n_alive = alive.shape[0]
        if n_alive <= 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        trace.append((n_alive, n_step))
        xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, alive, rays_t, rays_o, rays_d, m.bound, m.density_bitfield, m.cascade, m.grid_size,
                                                    nears, fars, 128, False, dt_gamma, max_steps)
        if fused_net:
            sig, rgb, aa, ae, un = m(xyzs, dirs, enc_a, c, e)
        else:
            with torch.autocast("cuda", dtype=torch.float16):
                sig, rgb, aa, ae, un = m.forward_unfused(xyzs, dirs, enc_a, c, e)
        with torch.autocast("cuda", dtype=torch.float16):      # the reference renders inside autocast: custom_fwd casts the fp16 outputs to fp32
            raymarching.composite_rays_triplane(n_alive, n_step, alive, rays_t, sig, rgb, deltas, aa, ae, un, ws, depth, image, s_aud, s_eye, s_unc, T_thresh)
        alive = alive[alive >= 0]
        step += n_step
    image = (image + (1 - ws).unsqueeze(-1) * 1.0).clamp(0, 1)
    return image, ws, depth, trace


@pytest.mark.parametrize("hw", [64, 512])
def test_render_frame_matches_host_loop(hw):
    """b2n_render_frame (device-side loop control, no host sync) == the reference loop on the same kernels, bit for bit;
    and within fp16-chain tolerance of the loop that evaluates the network op by op under autocast."""
    from b2nerf.model import HeadModel
    torch.manual_seed(0)
    m = HeadModel().cuda()
    for enc in (m.encoder_xy, m.encoder_yz, m.encoder_xz):
        enc.embeddings.data.uniform_(-1, 1)
    m.testing = True
    m.sigma_net.net[2].weight.data[0] *= 4.0        # livelier densities: some rays saturate (T < T_thresh) and terminate early
    m.density_bitfield.copy_(torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).cuda())
    m.pack()
    j, i = np.meshgrid(np.arange(hw), np.arange(hw), indexing="ij")
    o, d = scene.rays_for_pixels(scene.camera_pose(1), hw, hw, i.ravel(), j.ravel())
    rays_o, rays_d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    enc_a = torch.randn(1, 32, device="cuda") * 0.5; c = m.individual_codes[0:1].detach(); e = torch.tensor([[0.4]], device="cuda")
    img, ws, depth = m.render_frame(rays_o, rays_d, enc_a, c, e)
    r_img, r_ws, r_depth, trace = _loop_reference(m, rays_o, rays_d, enc_a, c, e, fused_net=True)
    torch.cuda.synchronize()
    assert len(trace) >= 5 and trace[0][1] == 1 and trace[1][1] >= 2, trace
    assert torch.equal(ws, r_ws) and torch.equal(depth, r_depth) and torch.equal(img, r_img)
    a_img, a_ws, _, a_trace = _loop_reference(m, rays_o, rays_d, enc_a, c, e, fused_net=False)
    assert float((ws > 0).float().mean()) > 0.2
    assert float((img - a_img).abs().mean()) < 2e-3 and float((ws - a_ws).abs().mean()) < 2e-3


@pytest.mark.parametrize("dim_in,L", [(1024, 2), (29, 16), (44, 16)])
def test_fused_audio_encoder_matches_torch_autocast(dim_in, L):
    """AudioNet + AudioAttNet as one cluster kernel vs the torch modules under autocast(fp16) (network.py:9-70, 226-240)."""
    from b2nerf.model import HeadModel
    torch.manual_seed(3)
    m = HeadModel(audio_in_dim=dim_in).cuda()
    for p in list(m.audio_net.parameters()) + list(m.audio_att_net.parameters()):     # livelier than the default init so every layer matters
        p.data.mul_(2.0)
    for trial in range(3):
        auds = torch.randn(8, dim_in, L, device="cuda")
        got = m.encode_audio_fused(auds)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            ref16 = m.encode_audio(auds).float()
        with torch.no_grad():
            ref32 = m.encode_audio(auds)
        torch.cuda.synchronize()
        assert got.shape == (1, 32) and torch.isfinite(got).all()
        scale = float(ref32.abs().max())
        # fp16 chain of 11 layers: agreement with the autocast graph at the 1e-3 level of the output scale, and as close to fp32 as autocast is
        assert float((got - ref16).abs().max()) < 4e-3 * scale, (float((got - ref16).abs().max()), scale)
        assert float((got - ref32).abs().max()) < 2.0 * float((ref16 - ref32).abs().max()) + 2e-3 * scale


def test_frame_renderer_while_graph_matches_fixed_sequence():
    """FrameRenderer's device-controlled WHILE-loop graph renders the same image as the eager fixed launch sequence, bit for bit,
    executes only the live iterations, and stays correct when replayed with new inputs."""
    from b2nerf.model import HeadModel
    from b2nerf.render import FrameRenderer
    torch.manual_seed(0)
    m = HeadModel().cuda()
    for enc in (m.encoder_xy, m.encoder_yz, m.encoder_xz):
        enc.embeddings.data.uniform_(-1, 1)
    m.testing = True
    m.density_bitfield.copy_(torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).cuda())
    hw = 256
    r_graph = FrameRenderer(m, hw * hw, use_graph=True)
    r_eager = FrameRenderer(m, hw * hw, use_graph=False)
    assert r_graph.loop_graph is not None, getattr(r_graph, "loop_graph_error", "no loop graph")
    for frame in (0, 1, 2):
        j, i = np.meshgrid(np.arange(hw), np.arange(hw), indexing="ij")
        o, d = scene.rays_for_pixels(scene.camera_pose(frame), hw, hw, i.ravel(), j.ravel())
        rays_o, rays_d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
        auds = torch.from_numpy(scene.audio_window(frame)).cuda()
        a = r_graph.render_device(rays_o, rays_d, auds).clone()
        b = r_eager.render_device(rays_o, rays_d, auds).clone()
        torch.cuda.synchronize()
        assert torch.equal(a, b)
        iters = r_graph.last_iterations()
        assert 3 <= iters <= 16 and iters == r_eager.last_iterations() or r_eager.last_iterations() >= iters
    assert float(a.mean()) < 0.999          # the head is visible against the white background


@pytest.mark.parametrize("n_out,n_in", [(64, 36), (32, 64), (16, 36), (1, 16), (64, 69), (65, 64), (64, 84), (3, 64), (128, 128), (32, 32),
                                         (8, 64), (72, 64), (64, 88), (64, 40), (16, 40), (8, 16)])     # padded pitches of the fused path: cp.async ring
@pytest.mark.parametrize("M", [128, 1000, 40064])
def test_linear_wgrad_matches_fp32_matmul(n_out, n_in, M):
    """b2n_linear_wgrad: dW += dY^T X over M samples (fp16 operands as MN-major tcgen05 operands, fp32 accumulation) vs torch in fp32.
    Products of two halves are exact in fp32, so the only difference is the summation order: 1e-4 of the result scale."""
    from b2nerf._lib import lib
    g = torch.Generator(device="cuda").manual_seed(M * 1000 + n_out * 7 + n_in)
    dy = torch.randn(M, n_out, device="cuda", generator=g).half()
    x = torch.randn(M, n_in, device="cuda", generator=g).half()
    dw = torch.full((n_out, n_in), 0.5, device="cuda")              # accumulates on top of the caller's buffer
    lib().call("b2n_linear_wgrad", dy.data_ptr(), x.data_ptr(), M, n_out, n_in, dw.data_ptr(), torch.cuda.current_stream().cuda_stream)
    ref = dy.float().t() @ x.float() + 0.5
    scale = float(ref.abs().max())
    assert float((dw - ref).abs().max()) <= 1e-4 * scale + 1e-3, float((dw - ref).abs().max())
    # a 2-byte-aligned (odd element offset) view of the operands takes the narrow-access path
    if M == 1000:
        buf = torch.zeros(M * n_in + 1, device="cuda", dtype=torch.float16)
        xo = buf[1:].view(M, n_in); xo.copy_(x)
        dw2 = torch.zeros(n_out, n_in, device="cuda")
        lib().call("b2n_linear_wgrad", dy.data_ptr(), xo.data_ptr(), M, n_out, n_in, dw2.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert float((dw2 - (ref - 0.5)).abs().max()) <= 1e-4 * scale + 1e-3


def test_head_forward_train_saves_the_reference_activations():
    """b2n_head_forward_train: same outputs as the inference kernel, and the saved fp16 activations equal the intermediate tensors of the
    reference graph under autocast (network.py:252-311) up to fp16 rounding of different accumulation orders."""
    m = _model()
    m.testing = False                      # unc_net is evaluated (network.py:276-278)
    m.pack()
    M = 4096 + 77                          # not a multiple of the 128-sample tile
    x, d = _samples(M, 9)
    enc_a = torch.randn(1, 32, device="cuda") * 0.5
    c, e = m.individual_codes[2:3].detach(), torch.tensor([[0.37]], device="cuda")
    sig, rgb, aud, eye_o, unc, sv = m.forward_train_fused(x, d, enc_a, c, e)
    o_sig, o_rgb, o_aud, o_eye, o_unc = m(x, d, enc_a, c, e)
    assert torch.equal(sig, o_sig) and torch.equal(rgb, o_rgb) and torch.equal(aud, o_aud.view(-1)) and torch.equal(eye_o, o_eye.view(-1)) and torch.equal(unc, o_unc.view(-1))
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        enc_x = m.encode_x(x)
        ha = torch.relu(m.aud_ch_att_net.net[0](enc_x)); att = m.aud_ch_att_net.net[1](ha)
        he = torch.relu(m.eye_att_net.net[0](enc_x)); eye_att = torch.sigmoid(m.eye_att_net.net[1](he))
        hu = torch.relu(m.unc_net.net[0](enc_x)); ul = m.unc_net.net[1](hu)
        s_in = torch.cat([enc_x, enc_a.repeat(M, 1) * att, e * eye_att], dim=-1)
        h1 = torch.relu(m.sigma_net.net[0](s_in)); h2 = torch.relu(m.sigma_net.net[1](h1)); o = m.sigma_net.net[2](h2)
        c_in = torch.cat([m.encoder_dir(d), o[..., 1:], c.repeat(M, 1)], dim=-1)
        hc = torch.relu(m.color_net.net[0](c_in)); s3 = torch.sigmoid(m.color_net.net[1](hc))
    def close(name, got, want, atol):
        got, want = got.float(), want.float()
        err = float((got - want).abs().max())
        assert err <= atol, (name, err)
    close("x36", sv["x36"][:, :36], enc_x, 2e-3)
    assert float(sv["x36"][:, 36:].abs().max()) == 0 and float(sv["s_in"][:, 36:40].abs().max()) == 0 and float(sv["s_in"][:, 73:].abs().max()) == 0
    assert float(sv["c_in"][:, 84:].abs().max()) == 0
    close("ha", sv["ha"], ha, 1e-2); close("he", sv["he"], he, 1e-2); close("hu", sv["hu"], hu, 1e-2); close("att", sv["att"], att, 1e-2)
    close("s_in", torch.cat([sv["s_in"][:, :36], sv["s_in"][:, 40:73]], dim=1), s_in, 1e-2)
    close("h1", sv["h1"], h1, 2e-2); close("h2", sv["h2"], h2, 2e-2)
    close("c_in", sv["c_in"][:, :84], c_in, 3e-2); close("hc", sv["hc"], hc, 3e-2)
    assert float((sv["misc"][:, 5] - 1).abs().max()) == 0 and float(sv["misc"][:, 6:].abs().max()) == 0
    close("misc rgb", sv["misc"][:, :3], s3, 1e-2); close("misc eye", sv["misc"][:, 3:4], eye_att, 5e-3); close("misc unc", sv["misc"][:, 4:5], ul, 1e-2)


def test_update_extra_state_fused_matches_reference_graph():
    """HeadModel.update_extra_state (renderer.py:699-768): the fused head kernel on the 128^3 jittered grid vs the reference graph (`density` through
    the per-op encoders + torch MLPs under autocast) with the same random jitter: same density grid up to fp16 noise, bitfields differ only on
    cells whose density sits at the threshold."""
    import copy
    m = _model(seed=4, table_scale=1.0, testing=True)
    with torch.no_grad():
        m.sigma_net.net[2].weight[0] *= 4.0           # spread the densities around the threshold
        for enc in (m.encoder_xy, m.encoder_yz, m.encoder_xz):
            enc.embeddings[int(enc.offsets[1]):] = 0   # only the coarsest level varies: a smooth field, so the dilated grid is not all-occupied
    m2 = copy.deepcopy(m)
    auds = torch.randn(8, m.audio_in_dim, 2, device="cuda")
    eye = torch.tensor([[0.3]], device="cuda")
    res = []
    for model, fused in ((m, True), (m2, False)):
        torch.manual_seed(11)                          # same jitter on both sides
        with torch.autocast("cuda", dtype=torch.float16):
            md = model.update_extra_state(auds, eye, fused=fused, density_thresh=1e9)      # threshold = mean density
        res.append((md, model.density_grid.clone(), model.density_bitfield.clone()))
    (md_f, g_f, b_f), (md_r, g_r, b_r) = res
    assert abs(md_f - md_r) <= 2e-3 * max(md_r, 1e-6), (md_f, md_r)
    rel = float(((g_f - g_r).abs() / (g_r.abs() + 1e-3)).max())
    assert rel < 3e-2, rel
    popc = torch.tensor([bin(i).count("1") for i in range(256)], device="cuda")
    diff_bits = int(popc[(b_f ^ b_r).long()].sum())
    assert diff_bits <= 1e-3 * m.grid_size ** 3, diff_bits
    occ = int(popc[b_f.long()].sum()) / m.grid_size ** 3
    assert 0.05 < occ < 0.95, occ                      # a non-trivial bitfield
    # second call applies the EMA decay path
    with torch.autocast("cuda", dtype=torch.float16):
        md2 = m.update_extra_state(auds, eye, fused=True, density_thresh=1e9)
    assert md2 > 0


def test_get_rays_and_rgb8_match_reference_formulas():
    """b2n_get_rays vs get_rays' all-pixel branch (utils.py:227-312, float64 numpy restatement in scene.rays_for_pixels) and b2n_image_to_rgb8 vs
    (image * 255).astype(uint8) (TrainerUtil.py:668, bit-exact)."""
    from b2nerf._lib import lib
    from b2nerf import scene
    H, W = 96, 160
    pose = scene.camera_pose(5).astype(np.float32)
    fx, fy, cx, cy = scene.intrinsics(H, W)
    pd = torch.from_numpy(pose).cuda().contiguous()
    ro, rd = torch.empty(H * W, 3, device="cuda"), torch.empty(H * W, 3, device="cuda")
    lib().call("b2n_get_rays", pd.data_ptr(), float(fx), float(fy), float(cx), float(cy), H, W, ro.data_ptr(), rd.data_ptr(), torch.cuda.current_stream().cuda_stream)
    j, i = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    o_ref, d_ref = scene.rays_for_pixels(pose.astype(np.float64), H, W, i.ravel(), j.ravel())
    # rays_for_pixels uses one focal derived from H: same as intrinsics()
    assert np.abs(rd.cpu().numpy() - d_ref).max() < 2e-6 and np.array_equal(ro.cpu().numpy(), o_ref)
    assert abs(float(rd.norm(dim=1).mean()) - 1.0) < 1e-6
    img = torch.rand(H * W, 3, device="cuda"); img[0] = 1.0; img[1] = 0.0
    u8 = torch.empty(H * W, 3, dtype=torch.uint8, device="cuda")
    lib().call("b2n_image_to_rgb8", img.data_ptr(), H * W, u8.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert np.array_equal(u8.cpu().numpy(), (img.cpu().numpy() * 255).astype(np.uint8))


def test_pose_in_rgb8_out_frame_equals_rays_in_frame():
    """FrameRenderer.render_host_pose (pose -> rays on the device -> frame -> RGB24) == the rays-in path fed with b2n_get_rays' rays, bit for bit."""
    from b2nerf.render import FrameRenderer
    from b2nerf._lib import lib
    from b2nerf import scene
    m = _model(seed=2)
    m.density_bitfield.copy_(torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).cuda())
    H = W = 128
    cam = (H, W) + tuple(scene.intrinsics(H, W))
    r = FrameRenderer(m, H * W, camera=cam)
    pose = torch.from_numpy(scene.camera_pose(3).astype(np.float32)).pin_memory()
    auds = torch.from_numpy(scene.audio_window(3)).pin_memory()
    out_u8 = torch.empty(H * W, 3, dtype=torch.uint8).pin_memory()
    r.render_host_pose(pose, auds, out_u8)
    torch.cuda.synchronize()
    img_pose = r.image.clone()
    ro, rd = torch.empty(H * W, 3, device="cuda"), torch.empty(H * W, 3, device="cuda")
    lib().call("b2n_get_rays", pose.cuda().data_ptr(), *[float(v) for v in cam[2:]], H, W, ro.data_ptr(), rd.data_ptr(), torch.cuda.current_stream().cuda_stream)
    out_f = torch.empty(H * W, 3).pin_memory()
    r.render_host(ro.cpu().pin_memory(), rd.cpu().pin_memory(), auds, out_f)
    torch.cuda.synchronize()
    assert torch.equal(r.image, img_pose)
    assert np.array_equal(out_u8.numpy(), (out_f.numpy() * 255).astype(np.uint8))
    assert 0.0 < float(img_pose.mean()) < 1.0 and int((out_u8 < 250).sum()) > 100      # the head is in the picture


def test_linear_wgrad_batch_matches_single_launches():
    """b2n_linear_wgrad_batch (all products of a backward in one launch, alternating TMEM accumulators) == the per-matrix kernel, in the fused
    path's shapes, incl. a row count that is not a multiple of the 64-sample chunk and more jobs than accumulators."""
    import ctypes
    from b2nerf._lib import lib
    from b2nerf.fused_train import _WgradJobC
    M = 20000 + 24
    shapes = [(8, 64), (64, 88), (72, 64), (64, 64), (64, 80), (32, 64), (64, 40), (8, 16), (16, 40), (32, 32), (8, 8), (8, 32), (32, 40)]
    g = torch.Generator(device="cuda").manual_seed(21)
    pairs = [(torch.randn(M, o, device="cuda", generator=g).half(), torch.randn(M, i, device="cuda", generator=g).half()) for o, i in shapes]
    R = 4
    total = sum(o * i for o, i in shapes)
    buf = torch.zeros(R, total, device="cuda")
    jobs, off = (_WgradJobC * len(pairs))(), 0
    for k, (dy, x) in enumerate(pairs):
        jobs[k] = _WgradJobC(dy.data_ptr(), x.data_ptr(), buf.data_ptr() + 4 * off, dy.shape[1], x.shape[1])
        off += dy.shape[1] * x.shape[1]
    st = torch.cuda.current_stream().cuda_stream
    lib().call("b2n_linear_wgrad_batch", jobs, len(pairs), M, R, total, st)
    flat, off = buf.sum(0), 0
    for (dy, x), (o, i) in zip(pairs, shapes):
        ref = dy.float().t() @ x.float()
        got = flat[off:off + o * i].view(o, i)
        scale = float(ref.abs().max())
        assert float((got - ref).abs().max()) <= 1e-4 * scale + 1e-3, ((o, i), float((got - ref).abs().max()), scale)
        off += o * i


def test_load_state_dict_refreshes_the_packed_model():
    """The packed model (fp16 operand images + corner-quad image) is a snapshot of weights and tables: load_state_dict on a packed HeadModel re-packs it."""
    a, b = _model(7, 1.0, True), _model(8, 1.0, True)
    x, d = _samples(5000, 3)
    enc_a = torch.randn(1, 32, device="cuda") * 0.5; c = a.individual_codes[1:2].detach().clone(); e = torch.tensor([[0.5]], device="cuda")
    a.pack(); b.pack()
    before = a(x, d, enc_a, c, e)[1].clone()
    want = b(x, d, enc_a, c, e)[1].clone()
    a.load_state_dict(b.state_dict())
    got = a(x, d, enc_a, c, e)[1].clone()
    torch.cuda.synchronize()
    assert torch.equal(got, want) and not torch.equal(before, want)


@pytest.mark.parametrize("world,n", [(2, 683509), (4, 40001)])
def test_peer_allreduce_kernel_ranks_as_streams_of_one_process(world, n):
    """csrc/peer_allreduce.cu on ONE GPU: `world` ranks = `world` buffers + streams of this process (b2n_peer_comm_create_local); every rank's kernel spins on
    flags the other ranks' kernels store, all co-resident.  Result = the mean, identical bits on every rank, repeated calls (epochs), inf propagation, no timeout.
    (The IPC / NVLink set-up between processes is covered by tests/test_gpu_multi.py on a 2-GPU box.)"""
    import ctypes
    from b2nerf import lib
    from b2nerf.dist import _DevMem
    L = lib()
    # (every rank's CTAs must be resident at once — each rank owns a whole GPU in a real job; on one GPU four full-size launches of 148 CTAs x 101 registers
    # would queue behind each other and trip the barrier timeout, so the 4-rank case uses a small buffer)
    nbytes = (n * 4 + 15) // 16 * 16
    ptrs, handles = [], []
    for r in range(world):
        p, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
        L.call("b2n_peer_alloc", nbytes, ctypes.byref(p), h)
        ptrs.append(p)
    arr = (ctypes.c_void_p * world)(*[p.value for p in ptrs])
    comms = []
    for r in range(world):
        c = ctypes.c_void_p()
        L.call("b2n_peer_comm_create_local", ctypes.byref(c), r, world, arr, nbytes)
        comms.append(c)
    bufs = [torch.as_tensor(_DevMem(p.value, nbytes // 4), device="cuda")[:n] for p in ptrs]
    streams = [torch.cuda.Stream() for _ in range(world)]
    g = torch.Generator(device="cuda").manual_seed(3)
    try:
        for it in range(6):
            xs = [torch.randn(n, device="cuda", generator=g) for _ in range(world)]
            if it == 5:
                xs[1][7] = float("inf")
            want = xs[0].clone()
            for x in xs[1:]:
                want = want + x
            want = want * (1.0 / world)
            for b, x in zip(bufs, xs):
                b.copy_(x)
            torch.cuda.synchronize()
            for r in range(world):
                L.call("b2n_peer_allreduce_mean", comms[r], n, streams[r].cuda_stream)
            torch.cuda.synchronize()
            for b in bufs:
                assert torch.equal(b, bufs[0])
            if it < 5:
                assert torch.equal(bufs[0], want), float((bufs[0] - want).abs().max())      # same summation order (rank 0, 1, ...), same scale
            else:
                assert bool(torch.isinf(bufs[0][7])) and bool(torch.isfinite(bufs[0][8]))
        err = ctypes.c_int32(-1)
        L.call("b2n_peer_error", comms[0], ctypes.byref(err), torch.cuda.current_stream().cuda_stream)
        assert err.value == 0
    finally:
        del bufs
        for c in comms:
            L.raw("b2n_peer_comm_destroy")(c)
        for p in ptrs:
            L.call("b2n_peer_free", p)


def test_four_warpgroup_tmem_head_kernel_equals_three_warpgroup_kernel(monkeypatch):
    """k_head_infer4 (4 tiles in flight per SM, hidden activations in tensor memory: tcgen05.st + MMA with the A operand in TMEM) runs the same MMAs on the same
    operands with the same epilogue arithmetic as k_head_forward (3 tiles, activations through shared memory): every output bit for bit, on a size with ragged
    last tiles, out-of-range samples and n_valid."""
    m = _model(5, 1.0, True)
    x, d = _samples(150003, 21)
    x[:4] = torch.tensor([[1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], [1.5, 0.2, 0.1], [0.0, 0.0, 0.0]], device="cuda")
    enc_a = torch.randn(1, 32, device="cuda") * 0.5; c = m.individual_codes[2:3].detach(); e = torch.tensor([[0.6]], device="cuda")
    m.pack()
    nv = torch.tensor([140001], dtype=torch.int32, device="cuda")
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("B2N_HEAD_WG4", flag)
        o = [t.clone() for t in m(x, d, enc_a, c, e)]
        o2 = [t.clone() for t in m(x, d, enc_a, None, None, n_valid=nv, out=tuple(torch.full_like(t, -3.0) for t in o))]
        outs.append(o + o2)
    torch.cuda.synchronize()
    for a_, b_, name in zip(outs[0], outs[1], ("sigma", "rgb", "aud", "eye", "unc") * 2):
        assert torch.equal(a_, b_), (name, float((a_ - b_).abs().max()))
    assert bool(torch.isfinite(outs[0][0]).all()) and float(outs[0][1].std()) > 1e-3


@pytest.mark.parametrize("wg4", ["1", "0"])
def test_head_tile_schedules_give_identical_outputs(monkeypatch, wg4):
    """The tile -> CTA mapping of the head kernels (balanced contiguous shares for a frame alone, grid-strided walk on a capped grid for frames in flight, the spread
    variant) only changes which SM evaluates a tile: every output bit for bit, full launches, device-side n_valid (short launches) and a capped grid."""
    m = _model(6, 1.0, True)
    x, d = _samples(200007, 33)
    enc_a = torch.randn(1, 32, device="cuda") * 0.5; c = m.individual_codes[1:2].detach(); e = torch.tensor([[0.3]], device="cuda")
    m.pack()
    monkeypatch.setenv("B2N_HEAD_WG4", wg4)
    ref = None
    for sched, ctas in (("0", None), ("1", None), ("2", None), ("2", "74"), ("0", "37"), ("2", "5")):
        monkeypatch.setenv("B2N_HEAD_SCHED", sched)
        if ctas is None:
            monkeypatch.delenv("B2N_HEAD_CTAS", raising=False)
        else:
            monkeypatch.setenv("B2N_HEAD_CTAS", ctas)
        outs = [t.clone() for t in m(x, d, enc_a, c, e)]
        for nv in (1, 127, 129, 40000, 200007):
            nvt = torch.tensor([nv], dtype=torch.int32, device="cuda")
            o2 = m(x, d, enc_a, c, e, n_valid=nvt, out=tuple(torch.full_like(t, -3.0) for t in outs[:5]))
            outs += [t.clone() for t in o2]
        torch.cuda.synchronize()
        if ref is None:
            ref = outs
            assert bool((ref[5][1:] == -3.0).all()) and bool((ref[5][:1] != -3.0).all())      # n_valid = 1: only the first row is written
        else:
            for k, (a_, b_) in enumerate(zip(ref, outs)):
                assert torch.equal(a_, b_), (sched, ctas, k, float((a_ - b_).abs().max()))


def test_tiled_first_ray_order_renders_the_same_image():
    """b2n_render_cfg.image_width only changes the ORDER in which an image-shaped frame walks its rays (8 x 16 pixel tiles instead of rows): every pixel bit for bit,
    WHILE-graph and eager path; a width that does not tile (not a multiple of 16) is ignored."""
    from b2nerf.render import FrameRenderer
    m = _model(seed=7)
    m.density_bitfield.copy_(torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).cuda())
    H, W = 96, 160
    j, i = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    o, d = scene.rays_for_pixels(scene.camera_pose(2), H, W, i.ravel(), j.ravel())
    rays_o, rays_d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    auds = torch.from_numpy(scene.audio_window(2)).cuda()
    ref = FrameRenderer(m, H * W, use_graph=True, image_width=0).render_device(rays_o, rays_d, auds).clone()
    for use_graph in (True, False):
        for width in (W, W - 8):          # 152 is not a multiple of 16 -> ignored (and H * W % 152 != 0)
            img = FrameRenderer(m, H * W, use_graph=use_graph, image_width=width).render_device(rays_o, rays_d, auds).clone()
            torch.cuda.synchronize()
            assert torch.equal(img, ref), (use_graph, width)
    assert 0.0 < float(ref.mean()) < 1.0

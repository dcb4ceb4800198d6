"""GPU: the fused torso kernel (csrc/fused_torso.cu) against the reference graph run_torso / forward_torso (renderer.py:572-631,
network.py:170-205) evaluated op by op on the drop-in encoders + torch Linear under autocast(fp16) — which is how the reference runs it."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(seed=0, deform_scale=0.02):
    from b2nerf.torso import TorsoModel
    torch.manual_seed(seed)
    m = TorsoModel().cuda().eval()
    m.torso_encoder.embeddings.data.uniform_(-0.5, 0.5)
    # keep the deformation small: dx is an fp16 value that moves the lookup in a 2048-cell grid; with |dx| ~ 1e-2 one fp16 ulp of dx (~1e-5) stays far
    # below the finest cell (1e-3), so the two paths read the same cells
    m.torso_deform_net.net[2].weight.data.mul_(deform_scale)
    G = m.grid_size
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, G), torch.linspace(-1, 1, G), indexing="ij")
    m.density_grid_torso.copy_((0.05 * torch.exp(-((xx * 1.2) ** 2 + ((yy - 0.5) * 1.5) ** 2) / 0.3)).reshape(-1).cuda())      # a torso-sized blob
    m.mean_density_torso = float(m.density_grid_torso.mean())
    return m


def _pose():
    th = math.radians(7.0)
    p = torch.eye(4)
    p[:3, :3] = torch.tensor([[math.cos(th), 0, math.sin(th)], [0, 1, 0], [-math.sin(th), 0, math.cos(th)]])
    p[:3, 3] = torch.tensor([0.1, -0.05, 3.35])
    return p[None].cuda()


@pytest.mark.parametrize("hw", [(64, 80), (512, 512)])
@pytest.mark.parametrize("bg_kind", ["white", "const", "per_ray"])
def test_fused_torso_matches_autocast_reference(hw, bg_kind):
    from b2nerf.torso import get_bg_coords
    m = _model(1)
    H, W = hw
    coords = get_bg_coords(H, W, "cuda")
    N = H * W
    g = torch.Generator(device="cuda").manual_seed(3)
    bg = {"white": None, "const": torch.tensor([0.2, 0.5, 0.9], device="cuda"), "per_ray": torch.rand(N, 3, device="cuda", generator=g)}[bg_kind]
    poses = _pose()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        ref = m.run_torso(coords, poses, 0, bg)
    out = m.run_torso_fused(coords, poses, 0, bg, want_deform=True)
    torch.cuda.synchronize()
    mask = ref["mask"]
    assert 0.05 < float(mask.float().mean()) < 0.9                       # the blob covers part of the image
    # occupancy mask: same bilinear arithmetic as F.grid_sample; pixels exactly at the threshold may flip
    on = out["torso_alpha"].view(-1) != 0
    flips = int((on & ~mask).sum())                                      # fused says torso where the reference did not
    assert flips <= max(2, N // 20000), flips
    a_ref, a_out = ref["torso_alpha"].view(-1), out["torso_alpha"].view(-1)
    da = (a_ref - a_out).abs()
    assert float(da.median()) < 3e-4 and float(da.max()) < 1.5e-2, (float(da.median()), float(da.max()))
    dc = (ref["bg_color"] - out["bg_color"]).abs()
    assert float(dc.median()) < 3e-4 and float(dc.max()) < 1.5e-2, (float(dc.median()), float(dc.max()))
    dxr = torch.zeros(N, 2, device="cuda"); dxr[mask] = ref["deform"].float()
    dd = (dxr - out["deform"]).abs()[mask & on]
    assert float(dd.max()) < 2e-4, float(dd.max())
    # outside the mask the background passes through untouched
    off = ~mask & ~on
    want = torch.ones(N, 3, device="cuda") if bg is None else (bg.expand(N, 3) if bg.numel() == 3 else bg)
    assert torch.equal(out["bg_color"][off], want[off])


def test_fused_torso_no_cpu_fallback_and_empty_mask():
    from b2nerf.torso import get_bg_coords
    m = _model(2)
    coords = get_bg_coords(32, 32, "cuda")
    with pytest.raises(RuntimeError):
        m.run_torso_fused(coords.cpu(), _pose(), 0, None)
    m.density_grid_torso.zero_(); m.mean_density_torso = 0.0             # threshold 0, occupancy 0: nothing is torso (renderer.py:603-606: strict >)
    out = m.run_torso_fused(coords, _pose(), 0, None)
    torch.cuda.synchronize()
    assert float(out["torso_alpha"].abs().max()) == 0 and torch.equal(out["bg_color"], torch.ones(32 * 32, 3, device="cuda"))


def test_frame_renderer_with_torso_background():
    """FrameRenderer(torso=...): every frame runs the fused torso kernel into the per-ray bg_color buffer the head frame reads (renderer.py:572-631 then
    :559-561) — same image, bit for bit, as calling the two stages by hand; the graph path equals the eager path; the torso shows behind the head."""
    from b2nerf import scene
    from b2nerf.model import HeadModel
    from b2nerf.render import FrameRenderer
    from b2nerf.torso import get_bg_coords
    torch.manual_seed(0)
    m = HeadModel().cuda()
    for enc in (m.encoder_xy, m.encoder_yz, m.encoder_xz):
        enc.embeddings.data.uniform_(-1, 1)
    m.testing = True
    m.density_bitfield.copy_(torch.from_numpy(scene.bitfield_from_grid(scene.density_grid())).cuda())
    t = _model(4)
    hw = 128
    coords = get_bg_coords(hw, hw, "cuda")
    pose = _pose()
    r_graph = FrameRenderer(m, hw * hw, use_graph=True, torso=t, bg_coords=coords)
    r_eager = FrameRenderer(m, hw * hw, use_graph=False, torso=t, bg_coords=coords)
    r_plain = FrameRenderer(m, hw * hw, use_graph=True)
    for r in (r_graph, r_eager):
        r.set_torso_pose(pose)
    j, i = np.meshgrid(np.arange(hw), np.arange(hw), indexing="ij")
    o, d = scene.rays_for_pixels(scene.camera_pose(1), hw, hw, i.ravel(), j.ravel())
    rays_o, rays_d, auds = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(scene.audio_window(1)).cuda()
    a = r_graph.render_device(rays_o, rays_d, auds).clone()
    b = r_eager.render_device(rays_o, rays_d, auds).clone()
    p = r_plain.render_device(rays_o, rays_d, auds).clone()
    bg = t.run_torso_fused(coords, pose, 0, None)["bg_color"]
    enc_a = m.encode_audio_fused(auds)
    c, _, _ = m.render_frame(rays_o, rays_d, enc_a, r_graph.ind_code, r_graph.eye, bg_color=bg)
    torch.cuda.synchronize()
    assert torch.equal(a, b) and torch.equal(a, c)
    assert not torch.equal(a, p) and float((a - p).abs().max()) > 0.05       # the torso changed the background pixels


def test_torso_occupancy_refresh_fused_matches_reference_graph():
    """TorsoModel.update_extra_state (renderer.py:768-809): the jittered 128^2 lattice through the fused kernel vs through forward_torso under autocast with
    the same jitter: same density grid up to the fp16 chain (1e-2 of values in [0, 1]), same mean to 1e-3; EMA-max decay applied."""
    import copy
    m = _model(5)
    m.density_grid_torso.fill_(0.02)
    m2 = copy.deepcopy(m)
    g = torch.Generator(device="cuda").manual_seed(7)
    noise = torch.rand(m.grid_size ** 2, 2, device="cuda", generator=g)
    pose = _pose()
    a = m.update_extra_state(pose, index=3, fused=True, noise=noise)
    b = m2.update_extra_state(pose, index=3, fused=False, noise=noise)
    torch.cuda.synchronize()
    d = (m.density_grid_torso - m2.density_grid_torso).abs()
    assert float(d.max()) < 1.5e-2 and float(d.median()) < 5e-4, (float(d.max()), float(d.median()))
    assert abs(a - b) < 1e-3 and a == m.mean_density_torso
    assert float(m.density_grid_torso.min()) >= 0.02 * 0.95 - 1e-7            # EMA-max: never below the decayed previous grid
    assert m.density_thresh_torso == 0.01                                      # the temporary -inf threshold was restored


def test_torso_training_step_through_the_per_op_graph():
    """Torso training (the reference's `--torso` stage, TrainerUtil.py:188-236 + renderer.py:572-631) runs through autograd on the drop-in kernels:
    freq_encode backward, the tiled fp16 grid backward (`grid_encode_backward`, half atomics) and torch Linear — every torso parameter gets a finite,
    non-zero gradient and a few AdamW steps reduce an image loss."""
    from b2nerf.torso import get_bg_coords
    m = _model(6, deform_scale=1.0).train()
    coords = get_bg_coords(96, 96, "cuda")
    pose = _pose()
    g = torch.Generator(device="cuda").manual_seed(5)
    target = torch.rand(96 * 96, 3, device="cuda", generator=g) * 0.5
    opt = torch.optim.AdamW(m.parameters(), lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    scaler = torch.amp.GradScaler("cuda")
    losses = []
    for it in range(6):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            out = m.run_torso(coords, pose, index=7, bg_color=None)
            loss = ((out["bg_color"] - target) ** 2).mean()
        scaler.scale(loss).backward()
        if it == 0:
            for name, p in m.named_parameters():
                if name == "individual_codes_torso":
                    assert p.grad is not None and float(p.grad[7].abs().sum()) > 0 and float(p.grad[:7].abs().sum()) == 0
                else:
                    assert p.grad is not None and bool(torch.isfinite(p.grad).all()) and float(p.grad.abs().sum()) > 0, name
        scaler.step(opt); scaler.update()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses


def test_fused_torso_against_the_cpu_port():
    """The fused torso kernel against the oracle's fp32 CPU restatement of run_torso / forward_torso (oracle/torch_port.py:torso_forward) on a 96x96 image:
    fp16 operands / activations with fp32 accumulation vs plain fp32 -> agreement at the 1e-2 level, identical occupancy mask up to threshold ties."""
    import math
    from oracle import torch_port
    from b2nerf.torso import get_bg_coords
    m = _model(9)
    coords = get_bg_coords(96, 96, "cuda")
    pose = _pose()
    hc = m.frame_constants(pose, 0).detach()
    out = m.run_torso_fused(coords, pose, 0, None, h_const=hc)
    p = {k: v.detach().float().cpu() for k, v in m.state_dict().items() if v.dtype.is_floating_point}
    p["torso_encoder.offsets"] = m.torso_encoder.offsets.cpu()
    p["S"], p["H"] = float(np.float32(math.log2(m.torso_encoder.per_level_scale))), m.torso_encoder.base_resolution
    bg, alpha, mask = torch_port.torso_forward(p, coords.view(-1, 2).cpu(), hc.view(-1).cpu(), m.density_grid_torso.cpu(), m.grid_size, m.density_thresh(),
                                               shrink=m.torso_shrink)
    on = out["torso_alpha"].view(-1).cpu() != 0
    assert int((on != mask).sum()) <= 2
    both = on & mask
    da = (out["torso_alpha"].view(-1).cpu() - alpha).abs()[both]
    db = (out["bg_color"].cpu() - bg).abs()[both]
    assert float(da.mean()) < 3e-3 and float(da.max()) < 3e-2, (float(da.mean()), float(da.max()))
    assert float(db.mean()) < 3e-3 and float(db.max()) < 3e-2, (float(db.mean()), float(db.max()))

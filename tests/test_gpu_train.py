"""GPU: the training step (march_rays_train -> head network -> composite_rays_train_triplane -> backward incl. grid_encode backward ->
AdamW) on the drop-in ops.  Gradients are checked against the pure-PyTorch port on the CPU (fp32 autograd through the same graph)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_port as tp
from b2nerf import scene

pytestmark = pytest.mark.gpu


def _setup(n_rays, seed=0, table_scale=0.5):
    from b2nerf.model import HeadModel
    from b2nerf.train import Trainer
    torch.manual_seed(seed)
    m = HeadModel(audio_in_dim=29).cuda()             # DeepSpeech-sized audio front-end keeps the CPU comparison light
    for enc in (m.encoder_xy, m.encoder_yz, m.encoder_xz):
        enc.embeddings.data.uniform_(-table_scale, table_scale)
    bf = scene.bitfield_from_grid(scene.density_grid())
    m.density_bitfield.copy_(torch.from_numpy(bf).cuda())
    o, d = scene.train_rays(seed, n_rays)
    auds = torch.from_numpy(scene.audio_window(seed, hubert=False))
    gt = torch.from_numpy(np.random.default_rng(seed).random((n_rays, 3)).astype(np.float32))
    return m, Trainer, bf, o, d, auds, gt


def test_train_step_gradients_match_cpu_port_fp32():
    n = 2048
    torch.backends.cudnn.allow_tf32 = False           # the reference disables TF32 (train.py:11-13); cuDNN convs of the audio nets would otherwise run in TF32
    torch.backends.cuda.matmul.allow_tf32 = False
    m, Trainer, bf, o, d, auds, gt = _setup(n)
    tr = Trainer(m, fp16=False)
    tr.global_step = tr.iters // 2                    # step_factor 0.5: the uncertainty / ambient terms of TrainerUtil.py:256-331 are all active
    face = torch.from_numpy(np.random.default_rng(5).random(n) < 0.5)
    rays_o, rays_d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    eye = torch.full((1, 1), 0.4, device="cuda"); bg = torch.ones(1, 3, device="cuda")
    tr.grads.zero_()
    out = tr.render_train(rays_o, rays_d, auds.cuda(), 3, eye, bg, perturb=False)
    loss = tr.loss(out, gt.cuda(), face.cuda())
    loss.backward()
    tr.grads.check_attached()
    # ---- CPU: same graph in pure PyTorch (fp32), samples from the C oracle's marcher ----
    import copy
    mc = copy.deepcopy(m).cpu()
    nears, fars = oracle.near_far_from_aabb(o, d, scene.AABB, 0.05)
    xyzs, dirs, deltas, rays, cnt = oracle.march_rays_train(o, d, bf, 1.0, 1 / 256, 16, 1, 128, n * 16, nears, fars, np.zeros(n, np.float32))
    tot = int(cnt[0])
    P = {k: v for k, v in mc.named_parameters()}
    p = {k: v for k, v in P.items()}
    p["encoder_xy.offsets"] = mc.encoder_xy.offsets; p["S"] = float(np.log2(mc.encoder_xy.per_level_scale)); p["H"] = 64; p["bound"] = 1.0
    enc_a = mc.encode_audio(auds)[0]
    sig, rgb, aa, ae, un = tp.head_forward(p, torch.from_numpy(xyzs[:tot]), torch.from_numpy(dirs[:tot]), enc_a, P["individual_codes"][3], torch.tensor([0.4]), testing=False)
    ws, s_a, s_e, s_u, dep, img = tp.composite_rays_train_triplane_ragged(sig, rgb, aa.abs(), ae.abs(), un, torch.from_numpy(deltas[:tot]), torch.from_numpy(rays.astype(np.int64)))
    img = (img + (1 - ws)[:, None]).clamp(0, 1)
    from b2nerf.fused_train import torch_head_loss      # the reference's loss, op by op (TrainerUtil.py:238-334) — plain torch, runs on the CPU
    cpu_loss = torch_head_loss(img, ws, s_a, s_e, gt, unc_sum=s_u, face_mask=face, step_factor=0.5, lambda_ent=1e-4, lambda_amb=1e-4, max_steps=16)
    cpu_loss.backward()
    assert abs(float(loss) - float(cpu_loss)) < 2e-4 * max(1.0, abs(float(cpu_loss))), (float(loss), float(cpu_loss))
    worst, mean_rel = {}, {}
    for name, q in m.named_parameters():
        g_gpu = q.grad.detach().cpu()
        g_cpu = P[name].grad if P[name].grad is not None else torch.zeros_like(P[name])
        denom = float(g_cpu.abs().max()) + 1e-12
        worst[name] = float((g_gpu - g_cpu).abs().max()) / denom
        mean_rel[name] = float((g_gpu - g_cpu).abs().sum()) / (float(g_cpu.abs().sum()) + 1e-12)
    # fp32 on both sides; differences: fma vs separate ops in the interpolation weights, ex2.approx, atomics order -> 3e-3 of the max gradient.
    # Looser bounds where the gradient is a sum of ~10^4 signed per-sample terms that largely cancel, so the fp32 noise floor of the
    # terms (~1e-8 absolute here) is a larger fraction of the result: the audio nets (through enc_a) and the tables (max |g| ~ 2e-6:
    # a handful of entries sit at 6e-3..1e-2 of the max while the aggregate error stays below 1e-3, checked separately).
    tol = lambda k: 3e-2 if k.startswith("audio") else (2e-2 if k.endswith("embeddings") else 3e-3)
    bad = {k: v for k, v in worst.items() if v > tol(k)}
    assert not bad, bad
    bad_mean = {k: v for k, v in mean_rel.items() if k.endswith("embeddings") and v > 1e-3}
    assert not bad_mean, bad_mean
    assert float(m.encoder_xy.embeddings.grad.abs().sum()) > 0 and float(m.sigma_net.net[0].weight.grad.abs().sum()) > 0


def test_train_steps_reduce_loss_fp16_autocast():
    n = 65536
    m, Trainer, bf, o, d, auds, gt = _setup(n, table_scale=1e-4)          # reference init
    tr = Trainer(m, fp16=True)
    rays_o, rays_d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    target = torch.tensor([0.2, 0.5, 0.7], device="cuda").expand(n, 3).contiguous()
    losses = []
    for step in range(24):
        l, m_buf = tr.train_step(rays_o, rays_d, auds.cuda(), target, index=step % 7)
        losses.append(float(l))
        if step == 15:
            tr.update_mean_count()
            assert tr.mean_count > 100000          # ~0.33 M samples for 65 536 rays (~5 samples/ray)
    assert all(np.isfinite(losses))
    assert np.mean(losses[-4:]) < 0.9 * np.mean(losses[:4]), losses
    assert m_buf == tr.mean_count + (128 - tr.mean_count % 128)           # steady state: mean_count-sized, 128-aligned buffers


def test_graphed_step_matches_eager_step():
    """train_step_graphed (forward + backward + all-reduce replayed from one CUDA graph) takes the same steps as the eager train_step."""
    import copy
    n = 8192
    m, Trainer, bf, o, d, auds, gt = _setup(n, seed=3)
    m2 = copy.deepcopy(m)
    rays_o, rays_d = torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda()
    a, g = auds.cuda(), gt.cuda()
    losses = []
    for model, graphed in ((m, False), (m2, True)):
        tr = Trainer(model, fp16=True)
        for s in range(16):                                  # the reference's warm-up: worst-case buffers, then the mean_count estimate
            tr.train_step(rays_o, rays_d, a, g, index=1, perturb=False)
        tr.update_mean_count()
        assert tr.mean_count > 0
        step = tr.train_step_graphed if graphed else tr.train_step
        ls = []
        for s in range(6):                                   # global_step 16..21: the first of them carries the smoothness regulariser (its own graph)
            loss, m_buf = step(rays_o, rays_d, a, g, index=1, perturb=False)
            ls.append(float(loss))
        losses.append(ls)
        if graphed:
            assert len(tr._graphs) == 2 and m_buf >= -(-tr.mean_count // Trainer.M_BUCKET) * Trainer.M_BUCKET
    eager, graphed = np.array(losses[0]), np.array(losses[1])
    # same arithmetic; the sample buffer is rounded up to M_BUCKET in graph mode (never truncates more rays than the eager step) and the
    # atomic scatter order differs -> tiny drift that grows over the steps
    assert np.allclose(eager, graphed, rtol=2e-2, atol=2e-3), (eager, graphed)
    for (k, p), (_, q) in zip(m.named_parameters(), m2.named_parameters()):
        assert torch.isfinite(q).all(), k


def test_flat_adamw_matches_torch_adamw():
    """FlatAdamW (one kernel over a flat buffer, csrc/optim.cu) vs torch.optim.AdamW with the reference's groups / betas, incl. a GradScaler
    step that is skipped because of an overflow."""
    from b2nerf.optim import FlatAdamW
    torch.manual_seed(0)
    shapes0, shapes1, shapes2 = [(1000, 1), (333, 1)], [(64, 36), (3, 64), (7,)], [(16, 32, 3), (8,)]
    mk = lambda shapes: [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    a0, a1, a2 = mk(shapes0), mk(shapes1), mk(shapes2)
    b0, b1, b2 = ([torch.nn.Parameter(p.detach().clone()) for p in a] for a in (a0, a1, a2))
    # network.py:332-356: tables (AdamW's default decay), networks (wd 0), audio_att_net (5 x lr_net, wd 1e-4); train.py:287 LambdaLR stepped every iteration
    ref = torch.optim.AdamW([{"params": b0, "lr": 1e-2}, {"params": b1, "lr": 1e-3, "weight_decay": 0}, {"params": b2, "lr": 5e-3, "weight_decay": 1e-4}],
                            betas=(0.0, 0.99), eps=1e-8)
    opt = FlatAdamW([{"params": a0, "lr": 1e-2, "weight_decay": 0.01}, {"params": a1, "lr": 1e-3, "weight_decay": 0.0}, {"params": a2, "lr": 5e-3, "weight_decay": 1e-4}],
                    betas=(0.0, 0.99), eps=1e-8, ema_decay=0.95, ema_update_interval=2)
    sched_a = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: 0.5 ** (it / 4))
    sched_b = torch.optim.lr_scheduler.LambdaLR(ref, lambda it: 0.5 ** (it / 4))
    flat_g = torch.zeros(opt.n, device="cuda")
    off = 0
    for p in a0 + a1 + a2:
        p.grad = flat_g[off:off + p.numel()].view_as(p); off += p.numel()
    opt.attach_grads(flat_g)
    # torch_ema.ExponentialMovingAverage restated (TrainerUtil.py:98-99, 1055-1056): shadow -= (1 - min(decay, (1 + n) / (10 + n))) * (shadow - param)
    shadow, n_upd = [q.detach().clone() for q in b0 + b1 + b2], 0
    a0, b0 = a0 + a1 + a2, b0 + b1 + b2
    a1, b1 = [], []
    sa, sb = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    for step in range(6):
        gs = [torch.randn_like(p) for p in a0 + a1]
        if step == 3:
            gs[1][5] = float("inf")                       # overflow: both must skip this step and halve the scale
        for p, q, g in zip(a0 + a1, b0 + b1, gs):
            p.grad.copy_(g * sa.get_scale()); q.grad = g * sb.get_scale()
        # GradScaler bookkeeping needs a scaled loss to have been produced
        sa.scale(torch.zeros((), device="cuda")); sb.scale(torch.zeros((), device="cuda"))
        sa.step(opt); sa.update()
        sb.step(ref); sb.update()
        sched_a.step(); sched_b.step()
        if (step + 1) % 2 == 0:                           # EMA every 2nd optimizer call — also on the overflowed step (ema.update() does not know about the skip)
            n_upd += 1
            d = min(0.95, (1 + n_upd) / (10 + n_upd))
            for sh, q in zip(shadow, b0):
                sh.sub_((1 - d) * (sh - q.detach()))
        assert sa.get_scale() == sb.get_scale()
    assert float(opt.step_count) == 5.0
    for p, q in zip(a0 + a1, b0 + b1):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), float((p - q).abs().max())
    off = 0
    for sh in shadow:
        got = opt.ema[off:off + sh.numel()].view_as(sh); off += sh.numel()
        assert torch.allclose(got, sh, rtol=2e-5, atol=2e-6), float((got - sh).abs().max())
    # checkpoint round trip keeps the moments (they live outside Optimizer.state)
    sd = opt.state_dict()
    opt.exp_avg.zero_(); opt.step_count.zero_()
    opt.load_state_dict(sd)
    assert float(opt.step_count) == 5.0 and float(opt.exp_avg.abs().sum()) > 0


def test_fused_head_gradients_match_autograd_under_autocast():
    """fused_head_train (forward with saved activations + backward-data kernel + weight-gradient kernel + grid backward) vs autograd through
    forward_unfused under autocast(fp16) on the same samples and the same upstream gradients.  Both sides round activations / gradients to fp16
    between layers in slightly different places: tolerance 4e-2 of each gradient's max (2e-2 typical), outputs 1e-2."""
    import copy
    from b2nerf.model import HeadModel, MLP
    from b2nerf.fused_train import fused_head_train, head_parameters
    torch.manual_seed(5)
    m = HeadModel(audio_in_dim=29).cuda()
    for enc in (m.encoder_xy, m.encoder_yz, m.encoder_xz):
        enc.embeddings.data.uniform_(-0.5, 0.5)
    m.testing = False
    m2 = copy.deepcopy(m)
    M = 20000 + 37
    g = torch.Generator(device="cuda").manual_seed(1)
    x = (torch.rand(M, 3, device="cuda", generator=g) * 2 - 1) * torch.tensor([1.0, 0.5, 1.0], device="cuda")
    d = torch.nn.functional.normalize(torch.randn(M, 3, device="cuda", generator=g), dim=1)
    enc_a0 = (torch.randn(1, 32, device="cuda", generator=g) * 0.5)
    eye = torch.tensor([[0.4]], device="cuda")
    ups = [torch.randn(M, device="cuda", generator=g) * 0.1, torch.randn(M, 3, device="cuda", generator=g), torch.randn(M, 1, device="cuda", generator=g) * 0.1,
           torch.randn(M, 1, device="cuda", generator=g) * 0.1, torch.randn(M, 1, 1, device="cuda", generator=g) * 0.1]
    results = []
    for model, fused in ((m, True), (m2, False)):
        enc_a = enc_a0.clone().requires_grad_(True)
        ind = model.individual_codes
        with torch.autocast("cuda", dtype=torch.float16):
            if fused:
                model.pack()
                outs = fused_head_train(model, x, d, enc_a, ind[3], eye)
            else:
                MLP.tall_linear = False
                try:
                    outs = model.forward_unfused(x, d, enc_a, ind[3], eye)
                finally:
                    MLP.tall_linear = True
            loss = sum((o.float() * u).sum() for o, u in zip(outs, ups))
        loss.backward()
        grads = {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}
        grads["enc_a"] = enc_a.grad.detach().float().clone()
        results.append(([o.detach().float() for o in outs], grads))
    (o_f, g_f), (o_r, g_r) = results
    for a, b, name in zip(o_f, o_r, ("sigma", "rgb", "amb_aud", "amb_eye", "unc")):
        assert a.shape == b.shape, (name, a.shape, b.shape)
        rel = float((a - b).abs().max()) / (float(b.abs().max()) + 1e-6)
        assert rel < 1e-2, (name, rel)
    assert set(g_r) <= set(g_f) | {"audio_net", "audio_att_net"}, set(g_r) - set(g_f)
    worst = {}
    for n, gr_ in g_r.items():
        if n not in g_f:
            continue
        denom = float(gr_.abs().max()) + 1e-12
        worst[n] = float((g_f[n] - gr_).abs().max()) / denom
    bad = {k: v for k, v in worst.items() if v > 4e-2}
    assert not bad, (bad, worst)
    assert len(worst) >= 14, worst


@pytest.mark.parametrize("with_unc,step_factor", [(False, 0.0), (True, 0.0), (True, 0.37), (True, 1.0)])
def test_fused_loss_matches_torch_loss_and_gradients(with_unc, step_factor):
    """b2n_head_loss_forward/backward vs the reference's loss (TrainerUtil.py:238-334: uncertainty-weighted MSE, loss_u, static-uncertainty, entropy 1e-4, masked /
    ramped ambient terms) written op by op in torch on the blended image, through autograd (fp32): value 2e-6, gradients 2e-5 of their max."""
    from b2nerf.fused_train import fused_head_loss, torch_head_loss
    g = torch.Generator(device="cuda").manual_seed(8)
    n = 65536 + 13
    image = torch.rand(n, 3, device="cuda", generator=g) * 0.9
    ws = torch.rand(n, device="cuda", generator=g)
    ws[:50] = 0.0; ws[50:100] = 1.0                      # both clamp regions of the entropy term
    image[100:200] = 1.5; image[200:300] = -0.25         # both clamp regions of the colour
    aud, eye = torch.rand(n, device="cuda", generator=g), torch.rand(n, device="cuda", generator=g)
    unc = torch.rand(n, device="cuda", generator=g) * 3.0
    unc[300:310] = 12.0                                  # a few rays whose softmax weight hits the clamp at 10
    face = torch.rand(n, device="cuda", generator=g) < 0.4
    gt = torch.rand(n, 3, device="cuda", generator=g)
    for bg, sf in ((torch.ones(1, 3, device="cuda"), step_factor), (torch.rand(n, 3, device="cuda", generator=g), torch.tensor([step_factor], device="cuda"))):
        a = [t.clone().requires_grad_(True) for t in (image, ws, aud, eye, unc)]
        b = [t.clone().requires_grad_(True) for t in (image, ws, aud, eye, unc)]
        lf = fused_head_loss(a[0], a[1], a[2], a[3], gt, bg, unc_sum=a[4] if with_unc else None, face_mask=face, step_factor=sf)
        img = (b[0] + (1 - b[1]).unsqueeze(-1) * bg).clamp(0, 1)
        lt = torch_head_loss(img, b[1], b[2], b[3], gt, unc_sum=b[4] if with_unc else None, face_mask=face, step_factor=step_factor)
        assert abs(float(lf) - float(lt)) < 2e-6 * max(1.0, abs(float(lt))), (float(lf), float(lt))
        (lf * 1024.0).backward(); (lt * 1024.0).backward()
        for x, y, name in zip(a, b, ("image", "ws", "aud", "eye", "unc")):
            if y.grad is None:
                assert x.grad is None or float(x.grad.abs().max()) == 0.0, name
                continue
            err = float((x.grad - y.grad).abs().max()) / (float(y.grad.abs().max()) + 1e-20)
            assert err < 2e-5, (name, err)


def test_fused_head_without_unc_net():
    """unc_loss = False: unc_net is not packed / evaluated (network.py:276-278 else-branch); the fused path must run with the unc buffers absent,
    give finite gradients for every other parameter and none for unc_net."""
    from b2nerf.model import HeadModel
    from b2nerf.fused_train import fused_head_train
    torch.manual_seed(6)
    m = HeadModel(audio_in_dim=29).cuda()
    for enc in (m.encoder_xy, m.encoder_yz, m.encoder_xz):
        enc.embeddings.data.uniform_(-0.5, 0.5)
    m.testing, m.unc_loss = False, False
    M = 16384 + 5
    g = torch.Generator(device="cuda").manual_seed(2)
    x = (torch.rand(M, 3, device="cuda", generator=g) * 2 - 1) * torch.tensor([1.0, 0.5, 1.0], device="cuda")
    d = torch.nn.functional.normalize(torch.randn(M, 3, device="cuda", generator=g), dim=1)
    enc_a = (torch.randn(1, 32, device="cuda", generator=g) * 0.5).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        m.pack()
        sig, rgb, aud, eye_att, unc = fused_head_train(m, x, d, enc_a, m.individual_codes[1], torch.tensor([[0.2]], device="cuda"))
        loss = sig.sum() * 1e-3 + rgb.sum() + aud.sum() + eye_att.sum()
    loss.backward()
    assert float((unc - 0.6931471805599453).abs().max()) < 1e-6          # log(1 + e^0)
    for n, p in m.named_parameters():
        if n.startswith("unc_net") or n.startswith("audio"):
            assert p.grad is None or float(p.grad.abs().sum()) == 0, n
        elif n == "individual_codes":
            assert float(p.grad[1].abs().sum()) > 0 and float(p.grad[0].abs().sum()) == 0
        else:
            assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().sum()) > 0, n
    assert torch.isfinite(enc_a.grad).all() and float(enc_a.grad.abs().sum()) > 0


@pytest.mark.parametrize("hubert", [True, False])
def test_fused_audio_backward_matches_autograd(hubert):
    """b2n_audio_encode + b2n_audio_backward inside autograd vs encode_audio (torch AudioNet + AudioAttNet).  Reference = the torch graph in fp32; the
    kernel (fp16-rounded forward operands like autocast, fp32 gradients) must be as close to it as torch's own autocast path is: AudioNet gradients
    3e-3 of their max; the attention net's gradients are ill-conditioned (they pass through a softmax over 8 nearly equal logits: 1e-3 in size against
    1e+3 for the AudioNet, and autocast itself is 2..17 % off the fp32 values), so they are held to 2.5x autocast's own error + 2 %."""
    import copy
    from b2nerf.model import HeadModel
    from b2nerf.fused_train import fused_encode_audio
    torch.manual_seed(9)
    m = HeadModel(audio_in_dim=1024 if hubert else 29).cuda()
    m2, m3 = copy.deepcopy(m), copy.deepcopy(m)
    auds = torch.from_numpy(scene.audio_window(4, hubert=hubert)).cuda()
    up = torch.randn(1, 32, device="cuda")
    with torch.autocast("cuda", dtype=torch.float16):
        e1 = fused_encode_audio(m, auds)
        e2 = m2.encode_audio(auds)
    e3 = m3.encode_audio(auds)                                   # fp32 reference
    for e in (e1, e2, e3):
        (e.float() * up).sum().mul(4096.0).backward()           # loss scale, as under GradScaler
    assert float((e1.float() - e3).abs().max()) < 2e-3 * max(1.0, float(e3.abs().max()))
    n_checked = 0
    for (n, p), (_, q), (_, r) in zip(m.named_parameters(), m2.named_parameters(), m3.named_parameters()):
        if not n.startswith("audio"):
            continue
        mx = float(r.grad.abs().max()) + 1e-12
        err_fused, err_autocast = float((p.grad - r.grad).abs().max()) / mx, float((q.grad - r.grad).abs().max()) / mx
        if n.startswith("audio_net"):
            assert err_fused < 3e-3, (n, err_fused, err_autocast)
        else:
            assert err_fused < 2.5 * err_autocast + 2e-2, (n, err_fused, err_autocast)
        n_checked += 1
    assert n_checked == 24


def test_combined_head_audio_node_equals_separate_nodes():
    """fused_head_audio_train (one autograd node, audio backward on a side stream next to the table-gradient kernel) gives the gradients of
    fused_encode_audio -> fused_head_train (two nodes, one stream)."""
    import copy
    from b2nerf.model import HeadModel
    from b2nerf.fused_train import fused_head_train, fused_encode_audio, fused_head_audio_train
    torch.manual_seed(12)
    m = HeadModel(audio_in_dim=29).cuda()
    for enc in (m.encoder_xy, m.encoder_yz, m.encoder_xz):
        enc.embeddings.data.uniform_(-0.5, 0.5)
    m.testing = False
    m2 = copy.deepcopy(m)
    M = 16384
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.rand(M, 3, device="cuda", generator=g) * 2 - 1) * torch.tensor([1.0, 0.5, 1.0], device="cuda")
    d = torch.nn.functional.normalize(torch.randn(M, 3, device="cuda", generator=g), dim=1)
    auds = torch.from_numpy(scene.audio_window(2, hubert=False)).cuda()
    eye = torch.tensor([[0.3]], device="cuda")
    outs = []
    for model, combined in ((m, True), (m2, False)):
        with torch.autocast("cuda", dtype=torch.float16):
            model.pack()
            if combined:
                o = fused_head_audio_train(model, x, d, auds, model.individual_codes[4], eye)
            else:
                o = fused_head_train(model, x, d, fused_encode_audio(model, auds), model.individual_codes[4], eye)
            loss = (o[0].sum() * 1e-3 + o[1].sum() + o[2].sum() + o[3].sum()) * 64.0
        loss.backward()
        outs.append(o)
    torch.cuda.synchronize()
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)
    for (n, p), (_, q) in zip(m.named_parameters(), m2.named_parameters()):
        if q.grad is None:
            assert p.grad is None or float(p.grad.abs().sum()) == 0, n
            continue
        scale = float(q.grad.abs().max()) + 1e-20
        assert float((p.grad - q.grad).abs().max()) <= 1e-4 * scale, (n, float((p.grad - q.grad).abs().max()) / scale)      # atomics order only


def test_direct_gradient_accumulation_equals_autograd_outputs():
    """Trainer mode (model._direct_grads): the fused backward accumulates every product straight into the parameters' .grad views through ONE
    b2n_wgrad_scatter launch (tables and audio nets by red.global into .grad) — same numbers as returning the gradients through autograd.
    Also checks that it really accumulates (pre-filled .grad)."""
    import copy
    from b2nerf.model import HeadModel
    from b2nerf.dist import FlatGradBuffer
    from b2nerf.fused_train import fused_head_audio_train
    torch.manual_seed(11)
    m = HeadModel(audio_in_dim=29).cuda()
    for enc in (m.encoder_xy, m.encoder_yz, m.encoder_xz):
        enc.embeddings.data.uniform_(-0.5, 0.5)
    m.testing = False
    m2 = copy.deepcopy(m)
    M = 30000 + 11
    g = torch.Generator(device="cuda").manual_seed(2)
    x = (torch.rand(M, 3, device="cuda", generator=g) * 2 - 1) * torch.tensor([1.0, 0.5, 1.0], device="cuda")
    d = torch.nn.functional.normalize(torch.randn(M, 3, device="cuda", generator=g), dim=1)
    auds = torch.randn(8, 29, 16, device="cuda", generator=g)
    eye = torch.tensor([[0.4]], device="cuda")
    ups = [torch.randn(M, device="cuda", generator=g) * 0.1, torch.randn(M, 3, device="cuda", generator=g), torch.randn(M, 1, device="cuda", generator=g) * 0.1,
           torch.randn(M, 1, device="cuda", generator=g) * 0.1, torch.randn(M, 1, 1, device="cuda", generator=g) * 0.1]
    flats = []
    for model, direct in ((m, True), (m2, False)):
        fb = FlatGradBuffer(list(model.parameters()))
        fb.flat.fill_(0.25)                                   # accumulate, do not overwrite
        model._direct_grads = direct
        model.pack()
        with torch.autocast("cuda", dtype=torch.float16):
            outs = fused_head_audio_train(model, x, d, auds, model.individual_codes[3], eye)
            loss = sum((o.float() * u).sum() for o, u in zip(outs, ups))
        loss.backward()
        torch.cuda.synchronize()
        fb.check_attached()
        flats.append(fb.flat.clone())
    a, b = flats
    assert float((b - 0.25).abs().max()) > 1e-3
    # identical kernels produce the products; only the accumulation order of the fp32 sums differs (replica sum + 0.25 vs 0.25 + replica sum,
    # atomics order in the tables / audio nets)
    tol = 2e-5 * float((b - 0.25).abs().max()) + 1e-6
    assert float((a - b).abs().max()) <= tol, (float((a - b).abs().max()), tol)


def test_triplane_grid_backward_fixed_point_vs_float64():
    """b2n_triplane_grid_backward (fixed-point shared-memory privatisation, csrc/gridenc.cu:k_triplane_bwd_fix) against the scatter restated in torch
    with float64 accumulation: index rule of gridencoder.cu:54-72 (D = 2: dense levels i + j * (res + 1), hashed levels (i ^ j * 2654435761) mod
    size), weights / products formed in fp32 exactly like the kernel.  Gradients span 5 decades; tolerance 1e-5 of each level's largest entry
    (the fixed-point sums are exact to 2^-26 of the largest gradient of a slice; fp32 atomics would be ~1e-4 here).  A non-finite gradient must
    poison its table (GradScaler overflow detection)."""
    import math
    from b2nerf import lib
    from b2nerf.model import HeadModel
    from gridencoder.backend import grid_level_scales
    torch.manual_seed(4)
    m = HeadModel().cuda()
    enc = m.encoder_xy
    S, H, L = float(math.log2(enc.per_level_scale)), enc.base_resolution, 12
    offs = enc.offsets.cpu().tolist()
    M = 150003
    g = torch.Generator(device="cuda").manual_seed(3)
    xyz = torch.randn(M, 3, device="cuda", generator=g) * 0.25           # clustered: heavy collisions on the coarse levels
    xyz[:64] = torch.rand(64, 3, device="cuda", generator=g) * 2.4 - 1.2   # some out of range
    xyz[64] = torch.tensor([1.0, 1.0, 1.0]); xyz[65] = torch.tensor([-1.0, -1.0, -1.0])
    xyz = xyz.contiguous()
    mag = 10.0 ** (torch.rand(3, L, M, device="cuda", generator=g) * 5 - 4)
    grad = (torch.randn(3, L, M, device="cuda", generator=g) * mag).contiguous()
    grad[:, :, 100:200] = 0.0
    tabs = [torch.zeros(offs[-1], 1, device="cuda") for _ in range(3)]
    st = torch.cuda.current_stream().cuda_stream
    lib().call("b2n_triplane_grid_backward", grad.data_ptr(), xyz.data_ptr(), enc.offsets.data_ptr(), tabs[0].data_ptr(), tabs[1].data_ptr(), tabs[2].data_ptr(),
               M, L, S, H, 1.0, st)
    scales = grid_level_scales(S, H, L)
    u = (xyz + 1.0) * 0.5                                                  # exact in fp32 for bound = 1
    planes = ((0, 1), (1, 2), (0, 2))
    for p, (ca, cb) in enumerate(planes):
        ref = torch.zeros(offs[-1], dtype=torch.float64, device="cuda")
        ok = (u[:, ca] >= 0) & (u[:, ca] <= 1) & (u[:, cb] >= 0) & (u[:, cb] <= 1)
        for l in range(L):
            sc = scales[l]
            size = offs[l + 1] - offs[l]
            res = int(math.ceil(float(sc))) + 1
            stride = res + 1
            dense = stride * stride <= size
            pa = (u[:, ca].double() * sc.double() + 0.5).float(); pb = (u[:, cb].double() * sc.double() + 0.5).float()     # == fmaf in fp32
            ia, ib = pa.floor(), pb.floor()
            fa, fb = pa - ia, pb - ib
            ia, ib = ia.long(), ib.long()
            for k in range(4):
                wa = fa if (k & 1) else (1.0 - fa); wb = fb if (k & 2) else (1.0 - fb)
                ja, jb = ia + (k & 1), ib + ((k >> 1) & 1)
                if dense:
                    slot = ja + jb * stride
                else:
                    slot = (ja ^ ((jb * 2654435761) & 0xFFFFFFFF)) % size
                t = ((wa * wb) * grad[p, l]).double()
                ref.index_add_(0, (slot + offs[l])[ok], t[ok])
        got = tabs[p].view(-1).double()
        for l in range(L):
            a, b = got[offs[l]:offs[l + 1]], ref[offs[l]:offs[l + 1]]
            tol = 1e-5 * float(b.abs().max())
            assert float((a - b).abs().max()) <= tol, (p, l, float((a - b).abs().max()), tol)
    # overflow: an inf in one (plane, level) slab must reach that level's gradient as a non-finite value
    grad[1, 7, 5000] = float("inf")
    tabs = [torch.zeros(offs[-1], 1, device="cuda") for _ in range(3)]
    lib().call("b2n_triplane_grid_backward", grad.data_ptr(), xyz.data_ptr(), enc.offsets.data_ptr(), tabs[0].data_ptr(), tabs[1].data_ptr(), tabs[2].data_ptr(),
               M, L, S, H, 1.0, st)
    torch.cuda.synchronize()
    assert not bool(torch.isfinite(tabs[1][offs[7]:offs[8]]).all())
    assert bool(torch.isfinite(tabs[0]).all()) and bool(torch.isfinite(tabs[2]).all())


def test_flat_grad_scaler_matches_torch_grad_scaler():
    """FlatGradScaler (scale / found-inf / growth tracker on the device, non-finite check + unscale + skip + scale update inside FlatAdamW's launch chain,
    csrc/optim.cu) against torch.amp.GradScaler + torch.optim.AdamW: same parameters and the same scale after a sequence with two overflows and a growth."""
    from b2nerf.optim import FlatAdamW, FlatGradScaler
    torch.manual_seed(1)
    mk = lambda shapes: [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    a0, a1 = mk([(500, 1), (77, 1)]), mk([(16, 9), (5,)])
    b0, b1 = ([torch.nn.Parameter(p.detach().clone()) for p in a] for a in (a0, a1))
    ref = torch.optim.AdamW([{"params": b0, "lr": 1e-2}, {"params": b1, "lr": 1e-3, "weight_decay": 0}], betas=(0.0, 0.99), eps=1e-8)
    opt = FlatAdamW([{"params": a0, "lr": 1e-2, "weight_decay": 0.01}, {"params": a1, "lr": 1e-3, "weight_decay": 0.0}], betas=(0.0, 0.99), eps=1e-8)
    flat_g = torch.zeros(opt.n, device="cuda")
    off = 0
    for p in a0 + a1:
        p.grad = flat_g[off:off + p.numel()].view_as(p); off += p.numel()
    opt.attach_grads(flat_g)
    sa = FlatGradScaler("cuda", init_scale=1024.0, growth_interval=3)
    sb = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=3)
    for step in range(10):
        gs = [torch.randn_like(p) for p in a0 + a1]
        if step in (2, 7):
            gs[step % 3][1] = float("nan") if step == 2 else float("-inf")
        for p, q, g in zip(a0 + a1, b0 + b1, gs):
            p.grad.copy_(g * sa.get_scale()); q.grad = g * sb.get_scale()
        sb.scale(torch.zeros((), device="cuda"))
        sa.step(opt); sa.update()
        sb.step(ref); sb.update()
        assert sa.get_scale() == sb.get_scale(), (step, sa.get_scale(), sb.get_scale())
    assert float(opt.step_count) == 8.0
    for p, q in zip(a0 + a1, b0 + b1):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), float((p - q).abs().max())
    sd = sa.state_dict()
    sa.state.zero_(); sa.load_state_dict(sd)
    assert sa.get_scale() == sb.get_scale()

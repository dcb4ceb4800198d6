"""Training path of the head network on the fused kernels: forward = k_head_forward<SAVE> (csrc/fused_head.cu), backward = k_head_backward
(csrc/fused_head_bwd.cu) + one b2n_linear_wgrad per weight matrix (csrc/wgrad.cu) + the grid backward fed in its own layout.

Replaces, inside autograd, the graph NeRFNetwork.forward builds under autocast (network.py:215-311): three GridEncoder calls, eleven
nn.Linear (+ two for unc_net), SHEncoder, cat / repeat / relu / sigmoid / exp / norm — ~500 kernels per step in the reference-style path
(profiles/r1_train_step.md).  Numerics: autocast semantics (fp16 operands and activations, fp32 accumulation)."""
import ctypes
import math

import torch
from torch.amp import custom_bwd, custom_fwd

from ._lib import lib


class _GradsC(ctypes.Structure):
    """b2n_head_grads (include/b2nerf_fused.h)"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("d_rl", "d_hc", "d_o", "d_h2", "d_h1", "d_ew", "d_att", "d_ha", "d_el", "d_he", "d_ul", "d_hu", "d_ci", "d_planes")]


WGRAD_REPLICAS = 16      # copies of each weight gradient the CTAs reduce into (summed here): un-serialises the same-address reductions in the L2
GRAD_WIDTHS = dict(d_rl=8, d_hc=64, d_o=72, d_h2=64, d_h1=64, d_ew=32, d_att=32, d_ha=64, d_el=8, d_he=16, d_ul=8, d_hu=32, d_ci=8)


def head_parameters(model):
    """The parameters the fused Function differentiates, in the order its backward returns their gradients."""
    return [model.encoder_xy.embeddings, model.encoder_yz.embeddings, model.encoder_xz.embeddings,
            model.aud_ch_att_net.net[0].weight, model.aud_ch_att_net.net[1].weight, model.eye_att_net.net[0].weight, model.eye_att_net.net[1].weight,
            model.sigma_net.net[0].weight, model.sigma_net.net[1].weight, model.sigma_net.net[2].weight,
            model.color_net.net[0].weight, model.color_net.net[1].weight, model.unc_net.net[0].weight, model.unc_net.net[1].weight]


class _WgradJobC(ctypes.Structure):
    """b2n_wgrad_job (include/b2nerf_fused.h)"""
    _fields_ = [("dy", ctypes.c_void_p), ("x", ctypes.c_void_p), ("dw", ctypes.c_void_p), ("out_dim", ctypes.c_uint32), ("in_dim", ctypes.c_uint32)]


def _wgrad_all(pairs):
    """pairs: [(dY [M,o] fp16, X [M,i] fp16)] -> list of fp32 [o,i] = dY^T X.  All results share one zero-initialised buffer [replicas, total]
    (ONE fill), ONE launch streams all the products (b2n_linear_wgrad_batch), every CTA adds its partials to replica (cta mod replicas), and ONE sum
    over the replica axis finishes all of them."""
    dev = pairs[0][0].device
    sizes = [dy.shape[1] * x.shape[1] for dy, x in pairs]
    total = sum(sizes)
    buf = torch.zeros(WGRAD_REPLICAS, total, dtype=torch.float32, device=dev)
    jobs, off = (_WgradJobC * len(pairs))(), 0
    for k, ((dy, x), sz) in enumerate(zip(pairs, sizes)):
        jobs[k] = _WgradJobC(dy.data_ptr(), x.data_ptr(), buf.data_ptr() + 4 * off, dy.shape[1], x.shape[1])
        off += sz
    lib().call("b2n_linear_wgrad_batch", jobs, len(pairs), pairs[0][1].shape[0], WGRAD_REPLICAS, total, torch.cuda.current_stream().cuda_stream)
    flat = buf.sum(0)
    outs, off = [], 0
    for (dy, x), sz in zip(pairs, sizes):
        outs.append(flat[off:off + sz].view(dy.shape[1], x.shape[1]))
        off += sz
    return outs


class _WgradBlockC(ctypes.Structure):
    """b2n_wgrad_block (include/b2nerf_fused.h)"""
    _fields_ = [("src_off", ctypes.c_uint32), ("src_ld", ctypes.c_uint32), ("rows", ctypes.c_uint32), ("cols", ctypes.c_uint32), ("dst", ctypes.c_void_p),
                ("dst_ld", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


def _wgrad_all_direct(pairs, blocks_of):
    """Same launch as _wgrad_all, but the replicas are summed AND accumulated into their destinations by one b2n_wgrad_scatter launch.
    blocks_of(k, off, in_dim, scratch_ptr) -> [(src_off, src_ld, rows, cols, dst_ptr, dst_ld)] for product k.  Returns the 64-float scratch tail
    (zero-initialised with the replica buffer) that blocks may target."""
    dev = pairs[0][0].device
    sizes = [dy.shape[1] * x.shape[1] for dy, x in pairs]
    total = sum(sizes)
    buf = torch.zeros(WGRAD_REPLICAS * total + 64, dtype=torch.float32, device=dev)
    scratch = buf[WGRAD_REPLICAS * total:]
    jobs, off, blocks = (_WgradJobC * len(pairs))(), 0, []
    for k, ((dy, x), sz) in enumerate(zip(pairs, sizes)):
        jobs[k] = _WgradJobC(dy.data_ptr(), x.data_ptr(), buf.data_ptr() + 4 * off, dy.shape[1], x.shape[1])
        blocks += blocks_of(k, off, x.shape[1], scratch.data_ptr())
        off += sz
    st = torch.cuda.current_stream().cuda_stream
    lib().call("b2n_linear_wgrad_batch", jobs, len(pairs), pairs[0][1].shape[0], WGRAD_REPLICAS, total, st)
    arr = (_WgradBlockC * len(blocks))(*[_WgradBlockC(so, sl, r, c, dst, dl, 0) for so, sl, r, c, dst, dl in blocks])
    lib().call("b2n_wgrad_scatter", buf.data_ptr(), WGRAD_REPLICAS, total, arr, len(blocks), st)
    return scratch


def _direct_targets(m, with_unc, audio):
    """The .grad tensors the fused backward may accumulate into in place (Trainer's flat gradient buffer), or None when the caller asked for
    ordinary autograd outputs (model._direct_grads unset, or some parameter has no contiguous fp32 .grad yet)."""
    if not getattr(m, "_direct_grads", False):
        return None
    ps = head_parameters(m)
    if not with_unc:
        ps = ps[:12]
    if audio:
        ps = ps + audio_parameters(m)
    for p_ in ps:
        g = p_.grad
        if g is None or g.dtype != torch.float32 or not g.is_contiguous() or not g.is_cuda:
            return None
    return True


class _FusedHead(torch.autograd.Function):
    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, model, x, d, enc_a, ind_code, eye, *params):
        sig, rgb, aud, eye_att, unc, saved = model.forward_train_fused(x, d, enc_a, ind_code, eye)
        ctx.model, ctx.saved_acts = model, saved
        ctx.with_unc = "hu" in saved
        x = x.contiguous()
        ctx.save_for_backward(x, enc_a, eye if eye is not None else torch.zeros(1, device=x.device), sig, aud)
        ctx.has_eye = eye is not None
        ctx.mark_non_differentiable()
        return sig, rgb, aud, eye_att, unc

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g_sig, g_rgb, g_aud, g_eye, g_unc):
        d_enc_a, d_ind, head, _ = _head_backward(ctx, g_sig, g_rgb, g_aud, g_eye, g_unc)
        return (None, None, None, d_enc_a, d_ind, None, *head)


_SIDE_STREAMS = {}


def _side_stream(dev, which=0):
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device(), which)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _SIDE_STREAMS[key]


def _head_backward(ctx, g_sig, g_rgb, g_aud, g_eye, g_unc, audio=None):
    """Shared backward of the fused head.  audio = (weights struct, auds): also back-propagate d enc_a through the audio nets, on a side stream,
    concurrently with the table-gradient kernel (8 CTAs next to a 148-SM kernel).  Returns (d_enc_a, d_ind, [14 head gradients], [24 audio gradients] | None)."""
    m, sv = ctx.model, ctx.saved_acts
    x, enc_a, eye, sig, aud = ctx.saved_tensors
    M, dev = x.shape[0], x.device
    f32 = lambda t: None if t is None else t.float().contiguous()
    g_sig, g_rgb, g_aud, g_eye, g_unc = f32(g_sig), f32(g_rgb), f32(g_aud), f32(g_eye), f32(g_unc)
    gr = {k: torch.empty(M, w, dtype=torch.float16, device=dev) for k, w in GRAD_WIDTHS.items() if ctx.with_unc or k not in ("d_ul", "d_hu")}
    planes = torch.empty(3, 12, M, dtype=torch.float32, device=dev)
    gc = _GradsC(*[(gr[k].data_ptr() if k in gr else None) for k, _ in _GradsC._fields_[:-1]], planes.data_ptr())
    from .model import _HeadSavedC
    sc = _HeadSavedC(*[sv[k].data_ptr() if k in sv else None for k, _ in _HeadSavedC._fields_])
    p = lambda t: None if t is None else t.data_ptr()
    enc_a_flat = enc_a.contiguous().view(-1)
    lib().call("b2n_head_backward", m.handle, M, enc_a_flat.data_ptr(), eye.data_ptr() if ctx.has_eye else None, ctypes.byref(sc), sig.data_ptr(),
               aud.data_ptr(), p(g_sig), p(g_rgb), p(g_aud), p(g_eye), p(g_unc), ctypes.byref(gc), torch.cuda.current_stream().cuda_stream)
    # weight gradients: one pass over (dY, X) per matrix
    pairs = [(gr["d_rl"], sv["hc"]), (gr["d_hc"], sv["c_in"]), (gr["d_o"], sv["h2"]), (gr["d_h2"], sv["h1"]), (gr["d_h1"], sv["s_in"]),
             (gr["d_att"], sv["ha"]), (gr["d_ha"], sv["x36"]), (gr["d_el"], sv["he"]), (gr["d_he"], sv["x36"])]
    # two reductions over the samples ride along as products: d enc_a[j] = sum_m d_ew[m,j] att[m,j] = diag(d_ew^T att);
    # d ind_code[i] = sum_m d_ci[m,i] = (d_ci^T misc)[i, 5]  (misc column 5 is a column of ones)
    pairs += [(gr["d_ew"], sv["att"]), (gr["d_ci"], sv["misc"])]
    if ctx.with_unc:
        pairs += [(gr["d_ul"], sv["hu"]), (gr["d_hu"], sv["x36"])]
    direct = _direct_targets(m, ctx.with_unc, audio is not None)
    if direct:
        # every product lands in its parameter's .grad (views of the trainer's flat all-reduce buffer) through ONE finishing launch: no sum / slice /
        # cat / contiguous / AccumulateGrad kernels.  Block = (src_off, src_ld, rows, cols, dst, dst_ld) inside product k (row pitch = its in_dim).
        G = lambda t: t.grad.data_ptr()
        net = dict(c1=m.color_net.net[1].weight, c0=m.color_net.net[0].weight, s2=m.sigma_net.net[2].weight, s1=m.sigma_net.net[1].weight,
                   s0=m.sigma_net.net[0].weight, a1=m.aud_ch_att_net.net[1].weight, a0=m.aud_ch_att_net.net[0].weight, e1=m.eye_att_net.net[1].weight,
                   e0=m.eye_att_net.net[0].weight, u1=m.unc_net.net[1].weight, u0=m.unc_net.net[0].weight)

        def blocks_of(k, off, ld, scratch):
            if k == 0: return [(off, ld, 3, 64, G(net["c1"]), 64)]
            if k == 1: return [(off, ld, 64, 84, G(net["c0"]), 84)]
            if k == 2: return [(off + 64 * ld, ld, 1, 64, G(net["s2"]), 64), (off, ld, 64, 64, G(net["s2"]) + 4 * 64, 64)]      # logit row 64 -> row 0
            if k == 3: return [(off, ld, 64, 64, G(net["s1"]), 64)]
            if k == 4: return [(off, ld, 64, 36, G(net["s0"]), 69), (off + 40, ld, 64, 33, G(net["s0"]) + 4 * 36, 69)]       # s_in = [enc_x 36 | pad 4 | enc_w 32 | e | pad 7]
            if k == 5: return [(off, ld, 32, 64, G(net["a1"]), 64)]
            if k == 6: return [(off, ld, 64, 36, G(net["a0"]), 36)]
            if k == 7: return [(off, ld, 1, 16, G(net["e1"]), 16)]
            if k == 8: return [(off, ld, 16, 36, G(net["e0"]), 36)]
            if k == 9: return [(off, ld + 1, 32, 1, scratch, 1)]                       # d enc_a = diag(d_ew^T att)
            if k == 10: return [(off + 5, ld, 4, 1, scratch + 4 * 32, 1)]              # d ind_code = column 5 of d_ci^T misc
            if k == 11: return [(off, ld, 1, 32, G(net["u1"]), 32)]
            return [(off, ld, 32, 36, G(net["u0"]), 36)]

        scratch = _wgrad_all_direct(pairs, blocks_of)
        d_enc_a, d_ind = scratch[:32].reshape(enc_a.shape), scratch[32:36]
        d_c1 = d_c0 = d_s2 = d_s1 = d_s0 = d_a1 = d_a0 = d_e1 = d_e0 = d_u1 = d_u0 = None
    else:
        w = _wgrad_all(pairs)
        d_c1, d_c0 = w[0][:3], w[1][:, :84]
        d_s2 = torch.cat([w[2][64:65], w[2][:64]], dim=0)                # rows: geo_feat 0..63, density logit 64 -> sigma_net.2's row order
        d_s1, d_s0 = w[3], torch.cat([w[4][:, :36], w[4][:, 40:73]], dim=1)      # s_in = [enc_x 36 | pad 4 | enc_w 32 | e | pad 7]
        d_a1, d_a0 = w[5], w[6][:, :36]
        d_e1, d_e0 = w[7][:1], w[8][:, :36]
        d_u1, d_u0 = (w[11][:1], w[12][:, :36]) if ctx.with_unc else (None, None)
        d_enc_a = torch.diagonal(w[9]).reshape(enc_a.shape)
        d_ind = w[10][:4, 5]
    audio_grads, side = None, None
    if audio is not None:
        aw, auds = audio
        params = audio_parameters(m)
        cur = torch.cuda.current_stream(dev)
        side = _side_stream(dev)
        side.wait_stream(cur)                                  # d enc_a is ready
        with torch.cuda.stream(side):
            if direct:                                         # the kernel accumulates (red.global) straight into the parameters' .grad
                audio_grads, flat = [None] * len(params), None
                ptr = [p_.grad.data_ptr() for p_ in params]
            else:
                flat = torch.zeros(sum(p_.numel() for p_ in params), dtype=torch.float32, device=dev)
                audio_grads, off = [], 0
                for p_ in params:
                    audio_grads.append(flat[off:off + p_.numel()].view_as(p_)); off += p_.numel()
                ptr = [v.data_ptr() for v in audio_grads]
            gs = _AudioGradsC((ctypes.c_void_p * 4)(*ptr[0:4]), (ctypes.c_void_p * 4)(*ptr[4:8]), (ctypes.c_void_p * 2)(*ptr[8:10]), (ctypes.c_void_p * 2)(*ptr[10:12]),
                              (ctypes.c_void_p * 5)(*ptr[12:17]), (ctypes.c_void_p * 5)(*ptr[17:22]), ptr[22], ptr[23])
            g_flat = d_enc_a.float().contiguous().view(-1)
            lib().call("b2n_audio_backward", ctypes.byref(aw), auds.data_ptr(), auds.shape[2], g_flat.data_ptr(), ctypes.byref(gs), side.cuda_stream)
            if flat is not None:
                flat.record_stream(cur)                        # consumed on the main stream after the join
            g_flat.record_stream(side)
    # table gradients: d enc_x is already in the grid backward's [plane][level][sample] layout; the three planes (xy, yz, xz:
    # network.py:208-212) go through one launch that takes its plane coordinates straight from xyz
    enc = m.encoder_xy
    S, H = float(math.log2(enc.per_level_scale)), enc.base_resolution
    if direct:                                                 # accumulate (red.global) into the tables' .grad
        d_tabs = [e_mod.embeddings.grad for e_mod in (m.encoder_xy, m.encoder_yz, m.encoder_xz)]
    else:
        d_tabs = [torch.zeros_like(e_mod.embeddings) for e_mod in (m.encoder_xy, m.encoder_yz, m.encoder_xz)]
    lib().call("b2n_triplane_grid_backward", planes.data_ptr(), x.data_ptr(), enc.offsets.data_ptr(), d_tabs[0].data_ptr(), d_tabs[1].data_ptr(),
               d_tabs[2].data_ptr(), M, 12, S, H, float(m.bound), torch.cuda.current_stream().cuda_stream)
    if direct:
        d_tabs = [None, None, None]
    if side is not None:
        torch.cuda.current_stream(dev).wait_stream(side)      # join before anybody reads the audio gradients
    c = lambda t: None if t is None else t.contiguous()
    return d_enc_a, d_ind, [d_tabs[0], d_tabs[1], d_tabs[2], c(d_a0), c(d_a1), c(d_e0), c(d_e1), c(d_s0), c(d_s1), c(d_s2), c(d_c0), c(d_c1), c(d_u0), c(d_u1)], audio_grads


def fused_head_train(model, x, d, enc_a, ind_code, eye):
    """Drop-in for HeadModel.forward_unfused in a training step: (sigma [M], color [M,3], ambient_aud [M,1], ambient_eye [M,1], unc [M,1,1])."""
    sig, rgb, aud, eye_att, unc = _FusedHead.apply(model, x, d, enc_a, ind_code.view(-1), eye, *head_parameters(model))
    return sig, rgb, aud[:, None], eye_att[:, None], unc[:, None, None]


class _LossArgsC(ctypes.Structure):
    """b2n_loss_args (include/b2nerf_fused.h)"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("image", "weights_sum", "aud_sum", "eye_sum", "unc_sum", "gt_rgb", "bg_color", "face_mask", "step_factor")] + \
               [(n, ctypes.c_float) for n in ("step_factor_host", "lambda_ent", "lambda_amb", "inv_max_steps")] + \
               [(n, ctypes.c_int) for n in ("bg_per_ray", "amb_aud_loss", "amb_eye_loss")]


class _FusedLoss(torch.autograd.Function):
    """The head branch's loss (TrainerUtil.py:238-334) + the background blend (renderer.py:559-561) as three kernels (csrc/loss.cu)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, image, ws, aud_sum, eye_sum, unc_sum, gt, bg, face_mask, step_factor, cfg):
        image, ws, aud_sum, eye_sum, gt, bg = (t.contiguous() for t in (image, ws, aud_sum, eye_sum, gt, bg))
        unc_sum = None if unc_sum is None else unc_sum.contiguous()
        n = ws.shape[0]
        per_ray = int(bg.numel() == 3 * n and n > 1)
        if face_mask is not None:
            face_mask = face_mask.contiguous().view(-1)
            face_mask = face_mask.view(torch.uint8) if face_mask.dtype == torch.bool else face_mask.to(torch.uint8)
        sf_dev = step_factor if torch.is_tensor(step_factor) else None
        p = lambda t: None if t is None else t.data_ptr()
        args = _LossArgsC(p(image), p(ws), p(aud_sum), p(eye_sum), p(unc_sum), p(gt), p(bg), p(face_mask), p(sf_dev), 0.0 if sf_dev is not None else float(step_factor),
                          float(cfg["lambda_ent"]), float(cfg["lambda_amb"]), 1.0 / float(cfg["max_steps"]), per_ray, int(cfg["amb_aud_loss"]), int(cfg["amb_eye_loss"]))
        stats = torch.empty(256, dtype=torch.float32, device=ws.device)
        lib().call("b2n_head_loss_forward", ctypes.byref(args), n, stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ctx.keep = (args, image, ws, aud_sum, eye_sum, unc_sum, gt, bg, face_mask, sf_dev, stats)
        ctx.n = n
        return stats[0]

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g):
        args, image, ws, aud_sum, eye_sum, unc_sum, gt, bg, face_mask, sf_dev, stats = ctx.keep
        g = g.float().contiguous()
        d_img, d_ws, d_aud, d_eye = torch.empty_like(image), torch.empty_like(ws), torch.empty_like(ws), torch.empty_like(ws)
        d_unc = torch.empty_like(ws) if unc_sum is not None else None
        lib().call("b2n_head_loss_backward", ctypes.byref(args), ctx.n, stats.data_ptr(), g.data_ptr(), d_img.data_ptr(), d_ws.data_ptr(), d_aud.data_ptr(),
                   d_eye.data_ptr(), None if d_unc is None else d_unc.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return d_img, d_ws, d_aud, d_eye, d_unc, None, None, None, None, None


def fused_head_loss(image, weights_sum, aud_sum, eye_sum, gt_rgb, bg_color, unc_sum=None, face_mask=None, step_factor=0.0, lambda_ent=1e-4, lambda_amb=1e-4,
                    max_steps=16, amb_aud_loss=True, amb_eye_loss=True):
    """`image` is the composite's UN-blended image; returns the scalar head-branch loss of TrainerUtil.train_step on clamp(image + (1 - ws) * bg, 0, 1).
    unc_sum=None drops the uncertainty terms (opt.unc_loss = 0); face_mask=None treats every ray as a face ray; step_factor: float or 1-element device tensor."""
    cfg = dict(lambda_ent=lambda_ent, lambda_amb=lambda_amb, max_steps=max_steps, amb_aud_loss=amb_aud_loss, amb_eye_loss=amb_eye_loss)
    return _FusedLoss.apply(image, weights_sum, aud_sum, eye_sum, unc_sum, gt_rgb, bg_color, face_mask, step_factor, cfg)


def torch_head_loss(image, weights_sum, aud_sum, eye_sum, gt_rgb, unc_sum=None, face_mask=None, step_factor=0.0, lambda_ent=1e-4, lambda_amb=1e-4, max_steps=16,
                    amb_aud_loss=True, amb_eye_loss=True):
    """The same loss op by op, in the reference's own order of operations (TrainerUtil.py:238-334) on the blended image; the parity partner of the kernels."""
    n = weights_sum.shape[0]
    sf = step_factor.reshape(()) if torch.is_tensor(step_factor) else step_factor          # a device scalar in a replayed CUDA graph
    face = torch.ones(n, dtype=torch.bool, device=image.device) if face_mask is None else face_mask.view(-1).bool()
    loss = ((image - gt_rgb) ** 2).mean(-1)
    if unc_sum is not None:
        alpha = 0.2
        unc_weight = torch.softmax(unc_sum, dim=-1) * n
        loss = loss * (alpha + (1 - alpha) * ((1 - sf) + sf * unc_weight.detach()).clamp(0, 10))
        beta = unc_sum + 1
        norm_rgb = torch.norm(image - gt_rgb, dim=-1).detach()
        loss_u = (norm_rgb / (2 * beta ** 2) + (torch.log(beta) ** 2) / 2) * face
        loss = loss + sf * loss_u + 1e-3 * sf * (unc_sum * (~face))
    loss = loss.mean()
    alphas = weights_sum.clamp(1e-5, 1 - 1e-5)
    loss = loss + lambda_ent * (-alphas * torch.log2(alphas) - (1 - alphas) * torch.log2(1 - alphas)).mean()
    lam = sf * lambda_amb
    if amb_aud_loss:
        loss = loss + lam * (aud_sum * (~face)).mean()
    if amb_eye_loss:
        loss = loss + lam * ((eye_sum / max_steps * aud_sum.detach()) * face).mean()
    return loss


class _AudioGradsC(ctypes.Structure):
    """b2n_audio_grads (include/b2nerf_fused.h)"""
    _fields_ = [("conv_w", ctypes.c_void_p * 4), ("conv_b", ctypes.c_void_p * 4), ("fc_w", ctypes.c_void_p * 2), ("fc_b", ctypes.c_void_p * 2),
                ("att_conv_w", ctypes.c_void_p * 5), ("att_conv_b", ctypes.c_void_p * 5), ("att_fc_w", ctypes.c_void_p), ("att_fc_b", ctypes.c_void_p)]


def audio_parameters(model):
    """AudioNet / AudioAttNet parameters in the order _AudioEncode.backward returns their gradients."""
    conv = [model.audio_net.encoder_conv[i] for i in (0, 2, 4, 6)]
    fc = [model.audio_net.encoder_fc1[i] for i in (0, 2)]
    att = [model.audio_att_net.attentionConvNet[i] for i in (0, 2, 4, 6, 8)]
    lin = model.audio_att_net.attentionNet[0]
    return ([c.weight for c in conv] + [c.bias for c in conv] + [c.weight for c in fc] + [c.bias for c in fc] + [c.weight for c in att] + [c.bias for c in att]
            + [lin.weight, lin.bias])


class _AudioEncode(torch.autograd.Function):
    """encode_audio (network.py:226-240) inside autograd on the two cluster kernels of csrc/fused_audio.cu (~100 cuDNN / elementwise launches per step in
    the torch path for an 8-frame window)."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, model, auds, *params):
        auds = auds.contiguous()
        w = model.audio_weights_struct()
        enc = torch.empty(1, 32, dtype=torch.float32, device=auds.device)
        lib().call("b2n_audio_encode", ctypes.byref(w), auds.data_ptr(), auds.shape[2], enc.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ctx.model, ctx.auds, ctx.w = model, auds, w
        return enc

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g_enc):
        params = audio_parameters(ctx.model)
        flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=g_enc.device)
        views, off = [], 0
        for p in params:
            views.append(flat[off:off + p.numel()].view_as(p)); off += p.numel()
        ptr = [v.data_ptr() for v in views]
        gs = _AudioGradsC((ctypes.c_void_p * 4)(*ptr[0:4]), (ctypes.c_void_p * 4)(*ptr[4:8]), (ctypes.c_void_p * 2)(*ptr[8:10]), (ctypes.c_void_p * 2)(*ptr[10:12]),
                          (ctypes.c_void_p * 5)(*ptr[12:17]), (ctypes.c_void_p * 5)(*ptr[17:22]), ptr[22], ptr[23])
        g = g_enc.float().contiguous().view(-1)
        lib().call("b2n_audio_backward", ctypes.byref(ctx.w), ctx.auds.data_ptr(), ctx.auds.shape[2], g.data_ptr(), ctypes.byref(gs),
                   torch.cuda.current_stream().cuda_stream)
        return (None, None) + tuple(views)


class _FusedHeadAudio(torch.autograd.Function):
    """encode_audio + NeRFNetwork.forward as ONE autograd node, so that the audio nets' backward (a latency-bound 8-CTA cluster kernel) runs on a side
    stream next to the table-gradient kernel instead of after it."""

    @staticmethod
    @custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, model, x, d, auds, ind_code, eye, n_head, ahead, *params):
        if ahead is not None:                                  # audio_encode_ahead: already running (or done) on its side stream
            aw, auds, enc_a, side = ahead
            cur = torch.cuda.current_stream(auds.device)
            cur.wait_stream(side)
            enc_a.record_stream(cur)
        else:
            auds = auds.contiguous()
            aw = model.audio_weights_struct()
            enc_a = torch.empty(1, 32, dtype=torch.float32, device=auds.device)
            lib().call("b2n_audio_encode", ctypes.byref(aw), auds.data_ptr(), auds.shape[2], enc_a.data_ptr(), torch.cuda.current_stream().cuda_stream)
        sig, rgb, aud, eye_att, unc, saved = model.forward_train_fused(x, d, enc_a, ind_code, eye)
        ctx.model, ctx.saved_acts, ctx.with_unc = model, saved, "hu" in saved
        ctx.audio = (aw, auds)
        x = x.contiguous()
        ctx.save_for_backward(x, enc_a, eye if eye is not None else torch.zeros(1, device=x.device), sig, aud)
        ctx.has_eye = eye is not None
        return sig, rgb, aud, eye_att, unc

    @staticmethod
    @custom_bwd(device_type="cuda")
    def backward(ctx, g_sig, g_rgb, g_aud, g_eye, g_unc):
        _, d_ind, head, audio = _head_backward(ctx, g_sig, g_rgb, g_aud, g_eye, g_unc, audio=ctx.audio)
        return (None, None, None, None, d_ind, None, None, None, *head, *audio)


def audio_encode_ahead(model, auds, cur):
    """Start encode_audio(auds) on a side stream forked from `cur` (the audio code does not depend on the step's rays: it runs beside the marcher).
    Returns the handle fused_head_audio_train(ahead=...) joins."""
    dev = auds.device
    auds = auds.contiguous()
    aw = model.audio_weights_struct()
    side = _side_stream(dev, 2)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        enc_a = torch.empty(1, 32, dtype=torch.float32, device=dev)
        lib().call("b2n_audio_encode", ctypes.byref(aw), auds.data_ptr(), auds.shape[2], enc_a.data_ptr(), side.cuda_stream)
    return aw, auds, enc_a, side


def fused_head_audio_train(model, x, d, auds, ind_code, eye, ahead=None):
    """encode_audio(auds) -> forward(x, d, enc_a, ind_code, eye) in a training step, both directions on the fused kernels (att > 0)."""
    hp = head_parameters(model)
    sig, rgb, aud, eye_att, unc = _FusedHeadAudio.apply(model, x, d, auds, ind_code.view(-1), eye, len(hp), ahead, *hp, *audio_parameters(model))
    return sig, rgb, aud[:, None], eye_att[:, None], unc[:, None, None]


def fused_encode_audio(model, auds):
    """Drop-in for HeadModel.encode_audio in a training step (att > 0): [8, dim_in, L] -> [1, 32] with gradients to the audio nets."""
    return _AudioEncode.apply(model, auds, *audio_parameters(model))

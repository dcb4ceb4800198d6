#!/usr/bin/env python
"""TEST INFRASTRUCTURE — executes the REFERENCE's own model code on a GPU and records what it computes.

What runs: the unmodified nerf_triplane/network.py (NeRFNetwork, AudioNet, AudioAttNet, MLP) + renderer.py (NeRFRenderer.run_cuda,
run_cuda_for_inference, run_torso, mark_untrained_grid, update_extra_state) + utils.get_rays + the reference's op wrappers
(raymarching.py, grid.py, sphere_harmonics.py, freq.py), staged byte for byte by oracle/stage_ref_py.sh into oracle/_ref_py/, on top of the
reference's own CUDA extensions (oracle/_ref/*.so built by oracle/build_ref_ext.sh).  No code of lzzx-nerf_b200/ is on that path: only
`b2nerf.scene` (numpy input synthesis) is imported, through tests/refcases.py.

    python tests/golden/make_golden_from_ref_model.py --out tests/golden/refmodel --sizes golden      # committed fixtures (small)
    python tests/golden/make_golden_from_ref_model.py --out /tmp/x --sizes full                       # what tests/test_gpu_refmodel.py runs live

Each case writes one .npz with the reference's OUTPUTS; inputs and weights are regenerated from tests/refcases.py by the consumer.
Large arrays are stored subsampled in golden mode (key + '.stride')."""
import argparse
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_PY = os.path.join(ROOT, "oracle", "_ref_py")
if not os.path.isdir(os.path.join(REF_PY, "nerf_triplane")):
    print("[make_golden_from_ref_model] oracle/_ref_py not staged (oracle/stage_ref_py.sh)"); sys.exit(3)
# the staged reference tree FIRST: `import raymarching` / `gridencoder` / ... must resolve to the reference's wrappers, not to the drop-ins
sys.path[:0] = [REF_PY, os.path.join(ROOT, "tests"), os.path.join(ROOT, "lzzx-nerf_b200")]

import numpy as np  # noqa: E402
import torch  # noqa: E402

for _ in range(32):           # third-party modules the reference imports at module scope but this path never calls (trimesh, lpips, mcubes, ...)
    try:
        import nerf_triplane.renderer as ref_renderer  # noqa: E402
        from nerf_triplane.network import NeRFNetwork  # noqa: E402
        from nerf_triplane.utils import get_rays  # noqa: E402
        break
    except ModuleNotFoundError as e:
        sys.modules[e.name] = types.ModuleType(e.name)
        for k in [k for k in sys.modules if k.startswith("nerf_triplane")]:
            del sys.modules[k]
import raymarching  # noqa: E402
import gridencoder  # noqa: E402

assert raymarching.__file__.startswith(REF_PY) and gridencoder.__file__.startswith(REF_PY), (raymarching.__file__, gridencoder.__file__)
assert raymarching.raymarching._backend.__name__ == "_ref_raymarching_face"

import refcases as rc  # noqa: E402

dev = torch.device("cuda")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
MAX_GOLDEN = 1 << 16


class Out:
    def __init__(self, mode):
        self.mode, self.d = mode, {}

    def put(self, name, t, exact=False):
        a = t.detach().float().cpu().numpy() if torch.is_tensor(t) and t.is_floating_point() else (t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t))
        if self.mode == "golden" and a.size > MAX_GOLDEN and not exact:
            stride = -(-a.size // MAX_GOLDEN)
            self.d[name + ".stride"] = np.int64(stride)
            a = a.reshape(-1)[::stride]
        self.d[name] = a

    def save(self, path):
        np.savez_compressed(path, **self.d)
        print(f"[ref model] wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB, {len(self.d)} arrays)")


def build(tag, torso=False, asr="hubert", smooth_lips=False, table_scale=1.0, seed=0):
    m = NeRFNetwork(rc.ref_opt(torso, asr, smooth_lips))
    rc.load_seeded(m, tag, seed, table_scale)
    m = m.to(dev)
    m.density_bitfield.copy_(T(rc.bitfield()))
    return m


def reset_renderer_globals():
    for g in ("zeroDepth", "zero_amb_aud_sum", "zero_amb_eye_sum", "zero_uncertainty_sum"):      # module-level caches sized by the first frame (renderer.py:477-489)
        setattr(ref_renderer, g, None)


def case_forward(sz, mode, outdir):
    """NeRFNetwork.forward (network.py:252-311) on random samples: autocast(fp16) — how the reference runs it — and fp32."""
    for table_scale, name in ((1.0, "forward"), (1e-4, "forward_refinit")):
        m = build("head_hubert", table_scale=table_scale).eval()
        n = sz["n_fwd"] if name == "forward" else 4096
        x, d, enc_a, eye = (T(a) for a in rc.forward_inputs(n))
        c = m.individual_codes[0]
        o = Out(mode)
        for testing in (True, False):
            m.testing = testing
            with torch.no_grad():
                with torch.autocast("cuda", dtype=torch.float16):
                    ra = m(x, d, enc_a, c, eye)
                rf = m(x, d, enc_a, c, eye)
            for tagp, r in (("amp", ra), ("f32", rf)):
                sig, rgb, aud, eyeo, unc = r
                p = f"{tagp}.testing{int(testing)}."
                o.put(p + "sigma", sig); o.put(p + "rgb", rgb); o.put(p + "amb_aud", aud); o.put(p + "amb_eye", eyeo)
                o.put(p + "unc_shape", np.array(unc.shape, np.int64))
                o.put(p + "unc", unc.reshape(n, -1)[:, 0])
        o.save(os.path.join(outdir, name + ".npz"))


def case_audio(sz, mode, outdir):
    """encode_audio = AudioNet + AudioAttNet (network.py:9-70, 226-240), HuBERT and DeepSpeech window layouts."""
    o = Out(mode)
    for tag, asr, hub in (("head_hubert", "hubert", True), ("head_deepspeech", "deepspeech", False)):
        m = build(tag, asr=asr).eval()
        for f in range(4):
            a = T(rc.audio_window(f, hubert=hub))
            with torch.no_grad():
                with torch.autocast("cuda", dtype=torch.float16):
                    o.put(f"{asr}.{f}.amp", m.encode_audio(a))
                o.put(f"{asr}.{f}.f32", m.encode_audio(a))
    o.save(os.path.join(outdir, "audio.npz"))


class LoopTrace:
    """Records (n_alive, n_step) of every loop iteration and keeps the accumulators the reference does not return, by wrapping the staged module's
    composite_rays_triplane — the reference's files are untouched."""

    def __enter__(self):
        self.orig, self.trace, self.ws = raymarching.composite_rays_triplane, [], None
        def wrapped(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, amb_aud, amb_eye, unc, weights_sum, depth, image, *rest):
            self.trace.append((int(n_alive), int(n_step)))
            self.ws = weights_sum
            return self.orig(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, amb_aud, amb_eye, unc, weights_sum, depth, image, *rest)
        ref_renderer.raymarching.composite_rays_triplane = wrapped
        return self

    def __exit__(self, *a):
        ref_renderer.raymarching.composite_rays_triplane = self.orig


def render_kwargs():
    return dict(dt_gamma=1.0 / 256, max_steps=16, T_thresh=1e-4)


def case_frame(sz, mode, outdir):
    """NeRFRenderer.render -> run_cuda_for_inference (renderer.py:406-570): one whole frame, as TrainerUtil.test_step calls it (no_grad + autocast)."""
    hw = sz["frame_hw"]
    m = build("head_hubert").eval(); m.testing = True
    o = Out(mode)
    for f in range(2):
        reset_renderer_globals()
        ro, rd, auds, eye = (T(a) for a in rc.frame_inputs(hw, f))
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16), LoopTrace() as lt:
            res, _ = m.render(ro[None], rd[None], auds, torch.zeros(1, hw * hw, 2, device=dev), None, eye=eye, index=[0], bg_color=None, perturb=False, **render_kwargs())
        o.put(f"{f}.image", res["image"].view(-1, 3)); o.put(f"{f}.weights_sum", lt.ws); o.put(f"{f}.trace", np.array(lt.trace, np.int64))
    # utils.get_rays (utils.py:227-312): rays of the frame's pose, for the device-side prologue (k_frame_rays)
    from b2nerf import scene
    pose = T(scene.camera_pose(0, seed=5).astype(np.float32))[None]
    r = get_rays(pose, scene.intrinsics(hw, hw), hw, hw, -1)
    o.put("get_rays.o", r["rays_o"].reshape(-1, 3)); o.put("get_rays.d", r["rays_d"].reshape(-1, 3))
    o.save(os.path.join(outdir, "frame.npz"))


def case_smooth_lips(sz, mode, outdir):
    """smooth_lips (renderer.py:456-460): enc_a <- 0.35 * previous + 0.65 * new, carried across consecutive frames."""
    hw = sz["lips_hw"]
    m = NeRFNetwork(rc.ref_opt(False, "hubert", smooth_lips=True)); rc.load_seeded(m, "head_hubert"); m = m.to(dev).eval(); m.testing = True
    m.density_bitfield.copy_(T(rc.bitfield()))
    reset_renderer_globals()
    o = Out(mode)
    for f in range(3):
        ro, rd, auds, eye = (T(a) for a in rc.frame_inputs(hw, 10 + f))
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            res, _ = m.render(ro[None], rd[None], auds, torch.zeros(1, hw * hw, 2, device=dev), None, eye=eye, index=[0], bg_color=None, perturb=False, **render_kwargs())
        o.put(f"{f}.image", res["image"].view(-1, 3)); o.put(f"{f}.enc_a", m.enc_a)
    o.save(os.path.join(outdir, "smooth_lips.npz"))


def case_train(sz, mode, outdir):
    """run_cuda's training branch (renderer.py:279-304, 553-570) + autograd backward of a fixed linear functional of its outputs, under autocast with a
    static loss scale (the reference uses GradScaler): outputs and the gradient of every parameter."""
    n, scale = sz["n_train"], rc.LOSS_SCALE
    m = build("head_deepspeech", asr="deepspeech", table_scale=0.5).train(); m.testing = False
    ro, rd, auds, eye, bg, w = rc.train_inputs(n)
    ro, rd, auds, eye, bg = (T(a) for a in (ro, rd, auds, eye, bg))
    w = {k: T(v) for k, v in w.items()}
    index = [3]
    with torch.autocast("cuda", dtype=torch.float16):
        res, _ = m.render(ro[None], rd[None], auds, torch.zeros(1, n, 2, device=dev), None, eye=eye, index=index, bg_color=bg, perturb=False, force_all_rays=False,
                          **render_kwargs())
        loss = rc.train_loss(res, w, n)
    (loss * scale).backward()
    o = Out(mode)
    for k in ("image", "weights_sum", "ambient_aud", "ambient_eye", "uncertainty", "depth"):
        o.put(k, res[k].reshape(n, -1))
    o.put("loss", loss.detach()); o.put("counter", m.step_counter[0]); o.put("n_samples_buffer", np.int64(res["rays"][0].shape[0]))
    for name, p in m.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.float() / scale
        o.put("grad." + name, g[index[0]] if name == "individual_codes" else g)
    o.save(os.path.join(outdir, "train.npz"))


def case_untrained(sz, mode, outdir):
    """mark_untrained_grid (renderer.py:633-697): cells no training camera sees get density -1."""
    m = build("head_hubert")
    poses, intr = rc.untrained_inputs()
    m.mark_untrained_grid(T(poses), intr)
    o = Out(mode)
    o.put("untrained_bits", np.packbits((m.density_grid < 0).cpu().numpy().reshape(-1)), exact=True)
    o.put("n_untrained", np.int64(int((m.density_grid < 0).sum())))
    o.save(os.path.join(outdir, "untrained.npz"))


def case_extra_state(sz, mode, outdir):
    """update_extra_state, head branch (renderer.py:699-766), called the way train_one_epoch does (TrainerUtil.py:1025-1029: under autocast), twice (the second
    call exercises the decayed EMA-max), after mark_untrained_grid."""
    m = build("head_hubert").train(); m.testing = False
    m.density_bitfield.zero_()
    feats, eye_area = rc.extra_state_inputs()
    m.aud_features, m.eye_area = torch.from_numpy(feats), torch.from_numpy(eye_area)
    poses, intr = rc.untrained_inputs()
    m.mark_untrained_grid(T(poses), intr)
    random.seed(7); torch.manual_seed(1234)
    o = Out(mode)
    for k in range(2):
        with torch.autocast("cuda", dtype=torch.float16):
            m.update_extra_state()
        o.put(f"{k}.density_grid", m.density_grid.view(-1)); o.put(f"{k}.bitfield", m.density_bitfield, exact=True); o.put(f"{k}.mean_density", np.float64(m.mean_density))
    o.save(os.path.join(outdir, "extra_state.npz"))


def case_torso(sz, mode, outdir):
    """run_torso / forward_torso (renderer.py:572-631, network.py:170-205): inference (eval, individual code 0) and a training backward."""
    hw = sz["torso_hw"]
    m = build("torso_hubert", torso=True).eval()
    g = T(rc.torso_grid())
    m.density_grid_torso.copy_(g); m.mean_density_torso = float(g.mean())
    coords, pose = T(rc.bg_coords(hw)), T(rc.torso_pose())
    N = hw * hw
    o = Out(mode)
    r = np.random.default_rng(4000)
    bg_ray = T(r.random((N, 3)).astype(np.float32))
    for name, bg in (("white", None), ("per_ray", bg_ray)):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            res = m.render_torso(torch.zeros(1, N, 3, device=dev), None, None, coords, pose, index=[0], bg_color=bg)
        mask = torch.zeros(N, dtype=torch.bool, device=dev)
        occ = torch.nn.functional.grid_sample(m.density_grid_torso.view(1, 1, 128, 128), coords.view(1, -1, 1, 2), align_corners=True).view(-1)
        mask = occ > min(m.density_thresh_torso, m.mean_density_torso)
        deform = torch.zeros(N, 2, device=dev); deform[mask] = res["deform"].float()
        o.put(name + ".bg_color", res["bg_color"]); o.put(name + ".torso_alpha", res["torso_alpha"]); o.put(name + ".deform", deform)
        o.put(name + ".n_mask", np.int64(int(mask.sum())))
    # training backward (TrainerUtil.py:188-236 torso stage: MSE on torso_color); here a fixed linear functional
    m.train()
    wt = T(r.standard_normal((N, 3)).astype(np.float32))
    scale = rc.LOSS_SCALE
    with torch.autocast("cuda", dtype=torch.float16):
        res = m.render_torso(torch.zeros(1, N, 3, device=dev), None, None, coords, pose, index=[5], bg_color=bg_ray)
        loss = (res["torso_color"] * wt).sum() / N
    (loss * scale).backward()
    o.put("train.loss", loss.detach())
    for name, p in m.named_parameters():
        if p.grad is not None:
            gr = p.grad.float() / scale
            o.put("grad." + name, gr[5] if name == "individual_codes_torso" else gr)
    o.save(os.path.join(outdir, "torso.npz"))


CASES = dict(forward=case_forward, audio=case_audio, frame=case_frame, smooth_lips=case_smooth_lips, train=case_train, untrained=case_untrained,
             extra_state=case_extra_state, torso=case_torso)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--sizes", default="golden", choices=list(rc.SIZES))
    ap.add_argument("--cases", default=",".join(CASES))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for name in a.cases.split(","):
        CASES[name](rc.SIZES[a.sizes], a.sizes, a.out)
    torch.cuda.synchronize()
    print("[ref model] done:", a.cases)

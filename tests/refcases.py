"""Seeded inputs + a deterministic weight recipe shared by
  * tests/golden/make_golden_from_ref_model.py — runs the REFERENCE's own NeRFNetwork / NeRFRenderer (staged by oracle/stage_ref_py.sh, on the
    reference's own CUDA extensions oracle/_ref/*.so) on a B200 and records its outputs, and
  * tests/test_gpu_refmodel.py — runs this repo's fused path on the same inputs and compares.
Pure numpy (no torch RNG): both sides, on any box, see identical bits.  Only OUTPUTS are stored in the golden files; inputs and
weights are regenerated from here."""
import json
import os
import types
import zlib

import numpy as np

from b2nerf import scene

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = json.load(open(os.path.join(HERE, "golden", "ref_state_dict_keys.json")))

# sizes: "golden" = what is committed under tests/golden/refmodel/ (small), "full" = the live run on the GPU box (BASELINE sizes)
SIZES = {
    "golden": dict(n_fwd=20000, frame_hw=160, n_train=8192, torso_hw=128, lips_hw=96),
    "full": dict(n_fwd=100003, frame_hw=512, n_train=65536, torso_hw=512, lips_hw=256),
}


LOSS_SCALE = 65536.0           # static loss scale of the backward goldens = GradScaler's initial scale (the reference trains under autocast + GradScaler)


def grad_errors(named_ours, golden, take, index_rows=None):
    """{name: (max error, L1 error)} of our gradients against the golden ones, each RELATIVE TO THE LARGEST GRADIENT OF THE PARAMETER'S NETWORK (audio_net,
    audio_att_net, sigma_net, ...; a tri-plane table is its own network): under autocast the reference's backward runs in fp16, where the smallest tensors of a
    network (the first attention convolutions: 1e-9 against 1e-7 for the network) underflow to exactly 0 — an error bar per tensor would compare against that
    underflow, one per network compares against what the optimizer step of that network sees."""
    rows = {}
    for name, gr in named_ours:
        key = "grad." + name
        if key not in golden:
            continue
        if index_rows and name in index_rows:
            gr = gr[index_rows[name]]
        rows[name] = take(golden, key, gr)
    net_of = lambda n: n.split(".")[0]
    net_max, net_l1 = {}, {}
    for name, (o, w) in rows.items():
        k = net_of(name)
        net_max[k] = max(net_max.get(k, 0.0), float(w.abs().max()))
        net_l1[k] = net_l1.get(k, 0.0) + float(w.abs().sum())
    return {name: (float((o - w).abs().max()) / (net_max[net_of(name)] + 1e-30), float((o - w).abs().sum()) / (net_l1[net_of(name)] + 1e-30)) for name, (o, w) in rows.items()}


def ref_opt(torso=False, asr="hubert", smooth_lips=False):
    """The option namespace the reference's NeRFNetwork / NeRFRenderer read (train.py:18-141 defaults with -O)."""
    return types.SimpleNamespace(bound=1, min_near=0.05, density_thresh=10, density_thresh_torso=0.01, exp_eye=True, test_train=False, smooth_lips=smooth_lips,
                                 torso=torso, cuda_ray=True, ind_num=10000, ind_dim=4, ind_dim_torso=8, train_camera=False, emb=False, asr_model=asr, att=2,
                                 torso_shrink=0.8, fix_eye=-1, smooth_eye=False, amb_dim=2, part=False, part2=False, unc_loss=1, lambda_amb=1e-4)


def _rng(name, seed):
    return np.random.default_rng((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)


def seeded_state_dict(tag="head_hubert", seed=0, table_scale=1.0):
    """{name: float32 ndarray} for every PARAMETER of the reference model `tag` (buffers and anchor_points keep their defaults); load with strict=False.
    Weights ~ U(+-gain/sqrt(fan_in)) like nn.Linear / Conv1d, tables ~ U(+-table_scale) (table_scale=1e-4 is the reference's own init, grid.py:132-134)."""
    out = {}
    for name, (shape, dtype) in KEYS[tag].items():
        if "float" not in dtype or name in ("aabb_train", "aabb_infer", "density_grid", "density_grid_torso", "anchor_points"):
            continue
        r = _rng(name, seed)
        if name.startswith("encoder_") and name.endswith("embeddings"):
            v = r.uniform(-table_scale, table_scale, shape)
        elif name == "torso_encoder.embeddings":
            v = r.uniform(-0.5, 0.5, shape)
        elif name.startswith("individual_codes"):
            v = r.standard_normal(shape) * 0.1
        elif name.endswith(".bias"):
            v = r.uniform(-0.1, 0.1, shape)
        else:
            fan_in = int(np.prod(shape[1:]))
            gain = 0.05 if name == "torso_deform_net.net.2.weight" else 1.5
            v = r.uniform(-1, 1, shape) * gain / np.sqrt(fan_in)
        out[name] = v.astype(np.float32)
    return out


def load_seeded(model, tag="head_hubert", seed=0, table_scale=1.0):
    import torch
    sd = {k: torch.from_numpy(v) for k, v in seeded_state_dict(tag, seed, table_scale).items()}
    own = model.state_dict()
    sd = {k: v for k, v in sd.items() if k in own}
    missing = [k for k, p in model.named_parameters() if k not in sd and k != "anchor_points"]
    assert not missing, missing
    model.load_state_dict(sd, strict=False)
    return model


# ---- case inputs -----------------------------------------------------------------------------------------------------------------------------
def forward_inputs(n, seed=0):
    r = np.random.default_rng(1000 + seed)
    x = (r.random((n, 3)) * 2 - 1).astype(np.float32)
    x[0] = [1.0, -1.0, 1.0]; x[1] = 0.0; x[2] = [-1.0, 0.3, 0.999]
    d = r.standard_normal((n, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    enc_a = (r.standard_normal((1, 32)) * 0.5).astype(np.float32)
    eye = np.array([[0.37]], np.float32)
    return x, d, enc_a, eye


def audio_window(frame, hubert=True):
    return scene.audio_window(frame, hubert=hubert, seed=3)


def frame_inputs(hw, frame=0):
    o, d = scene.frame_rays(frame=frame, H=hw, W=hw, seed=5)
    return o, d, audio_window(frame), np.array([[0.4]], np.float32)


def bitfield():
    return scene.bitfield_from_grid(scene.density_grid())


def train_inputs(n, step=0):
    o, d = scene.train_rays(step=step, n=n, seed=9)
    r = np.random.default_rng(2000 + step)
    bg = r.random((n, 3)).astype(np.float32)
    w = {k: r.standard_normal(s).astype(np.float32) for k, s in (("image", (n, 3)), ("weights_sum", (n,)), ("ambient_aud", (n,)), ("ambient_eye", (n,)),
                                                                 ("uncertainty", (n,)))}
    return o, d, audio_window(100 + step, hubert=False), np.array([[0.55]], np.float32), bg, w


def train_loss(out, w, n):
    """Scalar the training golden differentiates: a fixed random linear functional of every output of run_cuda's training branch."""
    return sum((out[k].reshape(w[k].shape).float() * w[k]).sum() for k in w) / n


def untrained_inputs():
    poses = np.stack([scene.camera_pose(f, jitter_deg=25.0, seed=11) for f in range(6)]).astype(np.float32)
    return poses, np.array(scene.intrinsics(512, 512), np.float32)


def extra_state_inputs():
    r = np.random.default_rng(3000)
    feats = r.standard_normal((16, 1024, 2)).astype(np.float32)
    eye_area = r.random((16, 1)).astype(np.float32)
    return feats, eye_area


def get_audio_features(features, index):
    """att_mode 2 window (utils.py:34-50): 8 frames centred on `index`, zero padded."""
    left, right = index - 4, index + 4
    pad_l, pad_r = max(0, -left), max(0, right - features.shape[0])
    a = features[max(left, 0):min(right, features.shape[0])]
    return np.concatenate([np.zeros((pad_l,) + a.shape[1:], a.dtype), a, np.zeros((pad_r,) + a.shape[1:], a.dtype)], 0)


def torso_grid(G=128):
    yy, xx = np.meshgrid(np.linspace(-1, 1, G), np.linspace(-1, 1, G), indexing="ij")
    return (0.05 * np.exp(-((xx * 1.2) ** 2 + ((yy - 0.5) * 1.5) ** 2) / 0.3)).reshape(-1).astype(np.float32)


def torso_pose():
    return scene.camera_pose(3, jitter_deg=8.0, seed=13).astype(np.float32)[None]


def bg_coords(hw):
    """utils.py:218-223 get_bg_coords: [1, H*W, 2] in [-1, 1]."""
    X = np.arange(hw, dtype=np.float32) / np.float32(hw - 1) * 2 - 1
    xs, ys = np.meshgrid(X, X, indexing="ij")
    return np.stack([xs.reshape(-1), ys.reshape(-1)], -1)[None].astype(np.float32)

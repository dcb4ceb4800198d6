#!/usr/bin/env python
"""Instantiates the REFERENCE's own NeRFNetwork (nerf_triplane/network.py, imported from /root/reference with this repo's drop-in encoder packages
on sys.path in front of it — the reference's extension packages cannot be imported without a GPU build) on the CPU and records the name, shape and dtype
of every state_dict entry, for the head model (torso=False) and the torso stage (torso=True).  Run in the build container:

    python tests/golden/make_ref_state_dict_keys.py          ->  tests/golden/ref_state_dict_keys.json

tests/test_cabi_and_surface.py checks HeadModel + TorsoModel against it (checkpoint compatibility)."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "lzzx-nerf_b200"))          # drop-in gridencoder / shencoder / freqencoder / raymarching

import torch  # noqa: E402

# third-party modules the reference imports at module scope but the network class never uses (trimesh, mcubes, cv2, ...): stub what is not installed
for _ in range(32):
    try:
        from nerf_triplane.network import NeRFNetwork  # noqa: E402
        break
    except ModuleNotFoundError as e:
        sys.modules[e.name] = types.ModuleType(e.name)
        for k in [k for k in sys.modules if k.startswith("nerf_triplane")]:
            del sys.modules[k]


def opt(torso, asr):
    return types.SimpleNamespace(bound=1, min_near=0.05, density_thresh=10, density_thresh_torso=0.01, exp_eye=True, test_train=False, smooth_lips=False,
                                 torso=torso, cuda_ray=True, ind_num=10000, ind_dim=4, ind_dim_torso=8, train_camera=False, emb=False, asr_model=asr, att=2,
                                 torso_shrink=0.8, fix_eye=-1, smooth_eye=False, amb_dim=2, part=False, part2=False, unc_loss=1, lambda_amb=1e-4)


out = {}
for tag, torso, asr in (("head_hubert", False, "hubert"), ("head_deepspeech", False, "deepspeech"), ("torso_hubert", True, "hubert")):
    m = NeRFNetwork(opt(torso, asr))
    out[tag] = {k: [list(v.shape), str(v.dtype)] for k, v in m.state_dict().items()}
    out[tag + "_n_params"] = sum(p.numel() for p in m.parameters())
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_state_dict_keys.json"), "w"), indent=0, sort_keys=True)
print({k: (len(v) if isinstance(v, dict) else v) for k, v in out.items()})

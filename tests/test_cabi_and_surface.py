"""CPU: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls without a GPU), and the
drop-in packages expose the reference's operator surface (names, positional signatures, state_dict layout)."""
import ctypes
import inspect
import os

import pytest
import torch

from b2nerf import _lib as L


def test_library_exports_every_declared_symbol():
    assert os.path.exists(L.LIB_PATH), "build with __graft_entry__.build()"
    dll = ctypes.CDLL(L.LIB_PATH)
    protos = L.declared_symbols()
    assert len(protos) >= 31
    for name in protos:
        assert hasattr(dll, name), f"{name} declared in include/*.h but not exported"
    for must in ("b2n_march_rays_train", "b2n_composite_rays_train_triplane_backward", "b2n_composite_rays_triplane", "b2n_grid_encode_forward",
                 "b2n_sh_encode_forward", "b2n_freq_encode_backward", "b2n_near_far_from_aabb", "b2n_packbits", "b2n_morton3D"):
        assert must in protos
    lib = L.lib()                      # binds argtypes from the header; raises if anything is missing
    assert lib.raw("b2n_version")() >= 100 and lib.launch_count() == 0


def test_header_prototypes_mirror_reference_argument_lists():
    protos = L.declared_symbols()
    # raymarching.h:14 march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises)
    assert protos["b2n_march_rays_train"][2] == ["rays_o", "rays_d", "grid", "bound", "dt_gamma", "max_steps", "N", "C", "H", "M", "nears", "fars",
                                                "xyzs", "dirs", "deltas", "rays", "counter", "noises", "stream"]
    assert protos["b2n_composite_rays_triplane"][2] == ["n_alive", "n_step", "T_thresh", "rays_alive", "rays_t", "sigmas", "rgbs", "deltas", "ambs_aud",
                                                       "ambs_eye", "uncertainties", "weights_sum", "depth", "image", "amb_aud_sum", "amb_eye_sum",
                                                       "uncertainty_sum", "stream"]
    assert protos["b2n_grid_encode_forward"][2][:13] == ["inputs", "embeddings", "offsets", "outputs", "B", "D", "C", "L", "S", "H", "dy_dx", "gridtype",
                                                         "align_corners"]


def test_dropin_operator_surface():
    import raymarching
    import gridencoder
    import shencoder
    import freqencoder
    from encoding import get_encoder
    names = ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "morton3D_dilation", "march_rays_train", "composite_rays_train",
             "march_rays", "composite_rays", "composite_rays_ambient", "composite_rays_train_sigma", "composite_rays_ambient_sigma",
             "composite_rays_train_uncertainty", "composite_rays_uncertainty", "composite_rays_train_triplane", "composite_rays_triplane"]
    for n in names:                                          # raymarching.py:48..671
        assert callable(getattr(raymarching, n)), n
    from raymarching.backend import _backend as rb
    for n in ["packbits", "near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "morton3D_dilation", "march_rays_train",
              "march_rays_train_backward", "composite_rays_train_forward", "composite_rays_train_backward", "march_rays", "composite_rays",
              "composite_rays_ambient", "composite_rays_train_sigma_forward", "composite_rays_train_sigma_backward", "composite_rays_ambient_sigma",
              "composite_rays_train_uncertainty_forward", "composite_rays_train_uncertainty_backward", "composite_rays_uncertainty",
              "composite_rays_train_triplane_forward", "composite_rays_train_triplane_backward", "composite_rays_triplane"]:
        assert callable(getattr(rb, n)), n                   # raymarching/src/bindings.cpp:5-38 (22 functions)
    sig = inspect.signature(raymarching.raymarching._march_rays_train.forward)
    assert list(sig.parameters)[1:] == ["rays_o", "rays_d", "bound", "density_bitfield", "C", "H", "nears", "fars", "step_counter", "mean_count",
                                         "perturb", "align", "force_all_rays", "dt_gamma", "max_steps"]
    assert sig.parameters["max_steps"].default == 1024 and sig.parameters["mean_count"].default == -1
    sig = inspect.signature(raymarching.raymarching._march_rays.forward)
    assert list(sig.parameters)[1:13] == ["n_alive", "n_step", "rays_alive", "rays_t", "rays_o", "rays_d", "bound", "density_bitfield", "C", "H", "near", "far"]
    sig = inspect.signature(gridencoder.grid._grid_encode.forward)
    assert list(sig.parameters)[1:] == ["inputs", "embeddings", "offsets", "per_level_scale", "base_resolution", "calc_grad_inputs", "gridtype", "align_corners"]
    enc, dim = get_encoder("hashgrid", input_dim=2, num_levels=12, level_dim=1, base_resolution=64, log2_hashmap_size=14, desired_resolution=512)
    assert dim == 12 and tuple(enc.embeddings.shape) == (163584, 1) and enc.offsets.dtype == torch.int32
    assert float(enc.embeddings.abs().max()) <= 1e-4                      # grid.py:132-134
    t, tdim = get_encoder("tiledgrid", input_dim=2, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16, desired_resolution=2048)
    assert tdim == 32 and t.embeddings.numel() == 1111040                 # network.py:166
    sh, shd = get_encoder("spherical_harmonics"); fq, fqd = get_encoder("frequency", input_dim=2, multires=8)
    assert shd == 16 and fqd == 34
    assert get_encoder("None", input_dim=5)[1] == 5
    with pytest.raises(NotImplementedError):
        get_encoder("ash")
    assert isinstance(shencoder.SHEncoder(degree=8), torch.nn.Module) and isinstance(freqencoder.FreqEncoder(), torch.nn.Module)


def test_no_cpu_fallback():
    """The product path must fail loudly without a CUDA device / with CPU tensors — never route to the oracle."""
    from gridencoder.backend import _backend as gb
    x = torch.rand(4, 2); emb = torch.rand(8, 1); offs = torch.tensor([0, 8], dtype=torch.int32); out = torch.empty(1, 4, 1)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        gb.grid_encode_forward(x, emb, offs, out, 4, 2, 1, 1, 0.0, 4, None, 0, False)
    import b2nerf
    pkg = os.path.dirname(os.path.dirname(b2nerf.__file__))
    for root, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(root, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{fn} imports the test oracle"


def test_head_model_state_dict_matches_reference_names_and_count():
    """#parameters = 683 509 for the HuBERT head model (test.ipynb:269) and reference parameter names."""
    from b2nerf.model import HeadModel
    m = HeadModel(audio_in_dim=1024)
    assert sum(p.numel() for p in m.parameters()) == 683509
    assert sum(p.numel() for p in HeadModel(audio_in_dim=29).parameters()) == 587989
    sd = m.state_dict()
    for k, shape in {"encoder_xy.embeddings": (163584, 1), "encoder_yz.offsets": (13,), "sigma_net.net.0.weight": (64, 69), "sigma_net.net.2.weight": (65, 64),
                     "color_net.net.0.weight": (64, 84), "color_net.net.1.weight": (3, 64), "unc_net.net.1.weight": (1, 32), "aud_ch_att_net.net.1.weight": (32, 64),
                     "eye_att_net.net.0.weight": (16, 36), "audio_net.encoder_conv.0.weight": (32, 1024, 3), "audio_net.encoder_fc1.2.weight": (32, 64),
                     "audio_att_net.attentionConvNet.8.weight": (1, 2, 3), "audio_att_net.attentionNet.0.weight": (8, 8), "individual_codes": (10000, 4),
                     "density_bitfield": (262144,), "density_grid": (1, 2097152), "step_counter": (16, 2), "aabb_infer": (6,)}.items():
        assert tuple(sd[k].shape) == shape, k


def test_torso_model_state_dict_matches_reference_names_and_count():
    """Torso branch (network.py:156-167, renderer.py:123-149): reference parameter / buffer names and shapes; the tiled grid D=2 L=16 C=2
    holds 1 111 040 parameters (network.py:166)."""
    from b2nerf.torso import TorsoModel, get_bg_coords
    m = TorsoModel()
    sd = m.state_dict()
    for k, shape in {"anchor_points": (3, 4), "torso_deform_net.net.0.weight": (32, 84), "torso_deform_net.net.1.weight": (32, 32), "torso_deform_net.net.2.weight": (2, 32),
                     "torso_encoder.embeddings": (555520, 2), "torso_encoder.offsets": (17,), "torso_net.net.0.weight": (32, 116), "torso_net.net.2.weight": (4, 32),
                     "individual_codes_torso": (10000, 8), "density_grid_torso": (128 * 128,)}.items():
        assert tuple(sd[k].shape) == shape, k
    assert m.torso_encoder.embeddings.numel() == 1111040
    assert m.torso_deform_encoder.output_dim == 34 and m.anchor_encoder.output_dim == 42
    c = get_bg_coords(4, 6, "cpu")                          # utils.py:218-223: first coordinate runs over the rows
    assert tuple(c.shape) == (1, 24, 2) and float(c[0, 0, 0]) == -1 and float(c[0, 5, 1]) == 1 and float(c[0, 6, 0]) > -1
    import pytest
    with pytest.raises(RuntimeError):                        # no CPU path
        m.run_torso_fused(c, None, 0, None, h_const=__import__("torch").zeros(50))


def test_state_dicts_match_the_reference_network_class():
    """Checkpoint compatibility pinned on the reference's OWN NeRFNetwork: tests/golden/ref_state_dict_keys.json lists every state_dict entry
    (name, shape, dtype) of nerf_triplane/network.py:NeRFNetwork instantiated from /root/reference (tests/golden/make_ref_state_dict_keys.py).  HeadModel must
    hold exactly the head entries; TorsoModel exactly what the torso stage adds."""
    import json
    import os
    import torch
    from b2nerf.model import HeadModel
    from b2nerf.torso import TorsoModel
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_state_dict_keys.json")))
    desc = lambda sd: {k: [list(v.shape), str(v.dtype)] for k, v in sd.items()}
    for tag, dim in (("head_hubert", 1024), ("head_deepspeech", 29)):
        assert desc(HeadModel(audio_in_dim=dim).state_dict()) == ref[tag], tag
        assert sum(p.numel() for p in HeadModel(audio_in_dim=dim).parameters()) == ref[tag + "_n_params"]
    extra = {k: v for k, v in ref["torso_hubert"].items() if k not in ref["head_hubert"]}
    assert desc(TorsoModel().state_dict()) == extra
    assert sum(p.numel() for p in TorsoModel().parameters()) == ref["torso_hubert_n_params"] - ref["head_hubert_n_params"]
    # a reference-style checkpoint dict loads into both models
    sd = {k: torch.zeros(s, dtype=getattr(torch, d.split(".")[1])) for k, (s, d) in ref["torso_hubert"].items()}
    HeadModel(audio_in_dim=1024).load_state_dict(sd, strict=False)
    missing, unexpected = TorsoModel().load_state_dict(sd, strict=False)
    assert not missing

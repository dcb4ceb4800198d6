"""`_backend` of the drop-in `raymarching` package.

Replaces the JIT-built pybind module `_raymarching_face` of the reference (raymarching/backend.py:31-38,
raymarching/src/bindings.cpp:5-38): same 22 function names and positional arguments, implemented on
libb2nerf.so through the C ABI of include/b2nerf.h.  All float tensors must be float32 (the reference's
wrappers cast with custom_fwd(cast_inputs=torch.float32); "scalar_t should always be float in use",
raymarching.cu:90).
"""
import types

import torch

from b2nerf.shim import call, dev_ptr, stream_ptr
from b2nerf._lib import lib

f32, i32, u8 = torch.float32, torch.int32, torch.uint8


def _p(t, name, dt=f32, optional=False):
    return dev_ptr(t, name, dt, optional)


def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
    call("b2n_near_far_from_aabb", _p(rays_o, "rays_o"), _p(rays_d, "rays_d"), _p(aabb, "aabb"), N, min_near,
         _p(nears, "nears"), _p(fars, "fars"), stream_ptr(rays_o))


def sph_from_ray(rays_o, rays_d, radius, N, coords):
    call("b2n_sph_from_ray", _p(rays_o, "rays_o"), _p(rays_d, "rays_d"), radius, N, _p(coords, "coords"), stream_ptr(rays_o))


def morton3D(coords, N, indices):
    call("b2n_morton3D", _p(coords, "coords", i32), N, _p(indices, "indices", i32), stream_ptr(coords))


def morton3D_invert(indices, N, coords):
    call("b2n_morton3D_invert", _p(indices, "indices", i32), N, _p(coords, "coords", i32), stream_ptr(indices))


def packbits(grid, N, density_thresh, bitfield):
    call("b2n_packbits", _p(grid, "grid"), N, density_thresh, _p(bitfield, "bitfield", u8), stream_ptr(grid))


def morton3D_dilation(grid, C, H, grid_dilation):
    call("b2n_morton3D_dilation", _p(grid, "grid"), C, H, _p(grid_dilation, "grid_dilation"), stream_ptr(grid))


def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays, counter, noises):
    # the marcher's scratch (occupied box, per-CTA totals, per-sample t cache) is a torch allocation of THIS call: stream-ordered by the caching allocator,
    # so concurrent marches on different streams and CUDA graphs that captured an earlier call never share or lose it
    ws = torch.empty(int(lib().raw("b2n_march_rays_train_workspace_bytes")(N, max_steps)), dtype=u8, device=rays_o.device)
    call("b2n_march_rays_train_ws", _p(rays_o, "rays_o"), _p(rays_d, "rays_d"), _p(grid, "grid", u8), bound, dt_gamma, max_steps,
         N, C, H, M, _p(nears, "nears"), _p(fars, "fars"), _p(xyzs, "xyzs"), _p(dirs, "dirs"), _p(deltas, "deltas"),
         _p(rays, "rays", i32), _p(counter, "counter", i32), _p(noises, "noises"), ws.data_ptr(), stream_ptr(rays_o))


def march_rays_train_backward(grad_xyzs, grad_dirs, rays, deltas, N, M, grad_rays_o, grad_rays_d):
    call("b2n_march_rays_train_backward", _p(grad_xyzs, "grad_xyzs"), _p(grad_dirs, "grad_dirs"), _p(rays, "rays", i32),
         _p(deltas, "deltas"), N, M, _p(grad_rays_o, "grad_rays_o"), _p(grad_rays_d, "grad_rays_d"), stream_ptr(rays))


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars, xyzs, dirs, deltas, noises):
    ws = torch.empty(int(lib().raw("b2n_march_rays_workspace_bytes")()), dtype=u8, device=rays_o.device)      # occupied box of the bitfield (see march_rays_train)
    call("b2n_march_rays_ws", n_alive, n_step, _p(rays_alive, "rays_alive", i32), _p(rays_t, "rays_t"), _p(rays_o, "rays_o"),
         _p(rays_d, "rays_d"), bound, dt_gamma, max_steps, C, H, _p(grid, "grid", u8), _p(nears, "nears"), _p(fars, "fars"),
         _p(xyzs, "xyzs"), _p(dirs, "dirs"), _p(deltas, "deltas"), _p(noises, "noises"), ws.data_ptr(), stream_ptr(rays_o))


# ---- composites: (C-ABI suffix, extra per-sample inputs, extra per-ray accumulators) ----------------------------
def _train_pair(suffix, n_extra):
    """Build <suffix>_forward / <suffix>_backward with the reference's argument order (raymarching.h:16-36)."""
    extra = ["ambient", "uncertainty"][:n_extra] if suffix != "triplane" else ["amb_aud", "amb_eye", "uncertainty"]
    k = len(extra)

    def forward(sigmas, rgbs, *rest):
        ex, (deltas, rays, M, N, T_thresh, weights_sum), tail = rest[:k], rest[k:k + 6], rest[k + 6:]
        sums, (depth, image) = tail[:k], tail[k:]
        call(f"b2n_composite_rays_train{'_' + suffix if suffix else ''}_forward", _p(sigmas, "sigmas"), _p(rgbs, "rgbs"),
             *[_p(t, n) for t, n in zip(ex, extra)], _p(deltas, "deltas"), _p(rays, "rays", i32), M, N, T_thresh,
             _p(weights_sum, "weights_sum"), *[_p(t, n + "_sum") for t, n in zip(sums, extra)], _p(depth, "depth"),
             _p(image, "image"), stream_ptr(sigmas))

    def backward(grad_weights_sum, *rest):
        g_ex, rest = rest[:k], rest[k:]
        grad_image, sigmas, rgbs = rest[:3]
        ex, rest = rest[3:3 + k], rest[3 + k:]
        deltas, rays, weights_sum = rest[:3]
        sums, rest = rest[3:3 + k], rest[3 + k:]
        image, M, N, T_thresh, grad_sigmas, grad_rgbs = rest[:6]
        g_out = rest[6:]
        call(f"b2n_composite_rays_train{'_' + suffix if suffix else ''}_backward", _p(grad_weights_sum, "grad_weights_sum"),
             *[_p(t, "grad_" + n + "_sum") for t, n in zip(g_ex, extra)], _p(grad_image, "grad_image"), _p(sigmas, "sigmas"),
             _p(rgbs, "rgbs"), *[_p(t, n) for t, n in zip(ex, extra)], _p(deltas, "deltas"), _p(rays, "rays", i32),
             _p(weights_sum, "weights_sum"), *[_p(t, n + "_sum") for t, n in zip(sums, extra)], _p(image, "image"), M, N, T_thresh,
             _p(grad_sigmas, "grad_sigmas"), _p(grad_rgbs, "grad_rgbs"), *[_p(t, "grad_" + n) for t, n in zip(g_out, extra)],
             stream_ptr(sigmas))

    return forward, backward


composite_rays_train_forward, composite_rays_train_backward = _train_pair("", 1)
composite_rays_train_sigma_forward, composite_rays_train_sigma_backward = _train_pair("sigma", 1)
composite_rays_train_uncertainty_forward, composite_rays_train_uncertainty_backward = _train_pair("uncertainty", 2)
composite_rays_train_triplane_forward, composite_rays_train_triplane_backward = _train_pair("triplane", 3)


def _infer(cname, extra):
    """Inference composites (raymarching.h:20-38): (n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas,
    <extra inputs>, weights, depth, image, <extra sums>)."""
    k = len(extra)

    def fn(n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, *rest):
        ex, (weights, depth, image), sums = rest[:k], rest[k:k + 3], rest[k + 3:]
        call(cname, n_alive, n_step, T_thresh, _p(rays_alive, "rays_alive", i32), _p(rays_t, "rays_t"), _p(sigmas, "sigmas"),
             _p(rgbs, "rgbs"), _p(deltas, "deltas"), *[_p(t, n) for t, n in zip(ex, extra)], _p(weights, "weights_sum"),
             _p(depth, "depth"), _p(image, "image"), *[_p(t, n + "_sum") for t, n in zip(sums, extra)], stream_ptr(sigmas))

    return fn


composite_rays = _infer("b2n_composite_rays", [])
composite_rays_ambient = _infer("b2n_composite_rays_ambient", ["ambients"])
composite_rays_ambient_sigma = _infer("b2n_composite_rays_ambient_sigma", ["ambients"])
composite_rays_uncertainty = _infer("b2n_composite_rays_uncertainty", ["ambients", "uncertainties"])
composite_rays_triplane = _infer("b2n_composite_rays_triplane", ["ambs_aud", "ambs_eye", "uncertainties"])

_NAMES = [
    "packbits", "near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "morton3D_dilation",
    "march_rays_train", "march_rays_train_backward", "composite_rays_train_forward", "composite_rays_train_backward",
    "march_rays", "composite_rays", "composite_rays_ambient",
    "composite_rays_train_sigma_forward", "composite_rays_train_sigma_backward", "composite_rays_ambient_sigma",
    "composite_rays_train_uncertainty_forward", "composite_rays_train_uncertainty_backward", "composite_rays_uncertainty",
    "composite_rays_train_triplane_forward", "composite_rays_train_triplane_backward", "composite_rays_triplane",
]
_backend = types.SimpleNamespace(**{n: globals()[n] for n in _NAMES})

__all__ = ["_backend"]

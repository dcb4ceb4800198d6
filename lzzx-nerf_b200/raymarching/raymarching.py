"""Drop-in `raymarching` operator surface on the B200-native library.

Same 17 callables, positional signatures, defaults, return shapes and autocast contract as the reference's
raymarching/raymarching.py (cited per function); the implementation is libb2nerf.so via `.backend._backend`.
Deliberate differences, none visible to renderer.py:
  * march_rays_train's sample allocation is deterministic (rays[] in ray order, contiguous segments);
  * the `torch.cuda.empty_cache()` the reference issues in the warm-up path (raymarching.py:256) is dropped;
  * inputs on the CPU are moved to the current CUDA device exactly like the reference does; there is no CPU path.
"""
import torch
from torch.autograd import Function
from torch.amp import custom_bwd, custom_fwd

from .backend import _backend

_fwd32 = custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_bwd = custom_bwd(device_type="cuda")


def _cuda(t):
    return t if t.is_cuda else t.cuda()


def _rays3(t):
    return _cuda(t).contiguous().view(-1, 3)


def _pad_to(m, align):
    # NB: adds a full `align` when already aligned — reference behaviour (raymarching.py:226-227, 250-251, 381-382)
    return m + (align - m % align) if align > 0 else m


# ----------------------------------------------------------------------------------------------------------------
# utils
# ----------------------------------------------------------------------------------------------------------------
class _near_far_from_aabb(Function):
    """raymarching.py:18-48 — (rays_o [N,3], rays_d [N,3], aabb [6], min_near) -> nears [N], fars [N]"""

    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        rays_o, rays_d = _rays3(rays_o), _rays3(rays_d)
        n = rays_o.shape[0]
        nears, fars = rays_o.new_empty(n), rays_o.new_empty(n)
        _backend.near_far_from_aabb(rays_o, rays_d, _cuda(aabb).contiguous(), n, min_near, nears, fars)
        return nears, fars


class _sph_from_ray(Function):
    """raymarching.py:51-79 — background-sphere (theta, phi) in [-1,1]^2"""

    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, radius):
        rays_o, rays_d = _rays3(rays_o), _rays3(rays_d)
        n = rays_o.shape[0]
        coords = rays_o.new_empty(n, 2)
        _backend.sph_from_ray(rays_o, rays_d, radius, n, coords)
        return coords


class _morton3D(Function):
    """raymarching.py:82-103 — int32 coords [N,3] -> int32 Morton indices [N] (no autocast decorator, like the reference)"""

    @staticmethod
    def forward(ctx, coords):
        coords = _cuda(coords).int().contiguous()
        n = coords.shape[0]
        indices = torch.empty(n, dtype=torch.int32, device=coords.device)
        _backend.morton3D(coords, n, indices)
        return indices


class _morton3D_invert(Function):
    """raymarching.py:105-125"""

    @staticmethod
    def forward(ctx, indices):
        indices = _cuda(indices).int().contiguous()
        n = indices.shape[0]
        coords = torch.empty(n, 3, dtype=torch.int32, device=indices.device)
        _backend.morton3D_invert(indices, n, coords)
        return coords


class _packbits(Function):
    """raymarching.py:128-154 — grid [C, H^3] > thresh -> bitfield uint8 [C*H^3/8] (optionally into `bitfield`)"""

    @staticmethod
    @_fwd32
    def forward(ctx, grid, thresh, bitfield=None):
        grid = _cuda(grid).contiguous()
        n = grid.shape[0] * grid.shape[1] // 8
        if bitfield is None:
            bitfield = torch.empty(n, dtype=torch.uint8, device=grid.device)
        _backend.packbits(grid, n, thresh, bitfield)
        return bitfield


class _morton3D_dilation(Function):
    """raymarching.py:157-180 — 6-neighbour max pool of a Morton-ordered grid [C, H^3]"""

    @staticmethod
    @_fwd32
    def forward(ctx, grid):
        grid = _cuda(grid).contiguous()
        cascades, h3 = grid.shape
        h = round(h3 ** (1.0 / 3.0))
        out = torch.empty_like(grid)
        _backend.morton3D_dilation(grid, cascades, h, out)
        return out


near_far_from_aabb = _near_far_from_aabb.apply
sph_from_ray = _sph_from_ray.apply
morton3D = _morton3D.apply
morton3D_invert = _morton3D_invert.apply
packbits = _packbits.apply
morton3D_dilation = _morton3D_dilation.apply


# ----------------------------------------------------------------------------------------------------------------
# marching
# ----------------------------------------------------------------------------------------------------------------
class _march_rays_train(Function):
    """raymarching.py:186-280.  Returns xyzs [M,3], dirs [M,3], deltas [M,2], rays int32 [N,3] = (id, offset, count)."""

    @staticmethod
    @_fwd32
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        rays_o, rays_d = _rays3(rays_o), _rays3(rays_d)
        density_bitfield = _cuda(density_bitfield).contiguous()
        n = rays_o.shape[0]
        use_estimate = (not force_all_rays) and mean_count > 0
        m = _pad_to(mean_count, align) if use_estimate else n * max_steps
        # zero-filled like the reference's (raymarching.py:231-233), in ONE allocation / fill: rows past the counter stay zero
        zbuf = rays_o.new_zeros(m * 8)
        xyzs, dirs, deltas = zbuf[:3 * m].view(m, 3), zbuf[3 * m:6 * m].view(m, 3), zbuf[6 * m:].view(m, 2)
        rays = torch.empty(n, 3, dtype=torch.int32, device=rays_o.device)
        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=rays_o.device)
        noises = torch.rand(n, dtype=rays_o.dtype, device=rays_o.device) if perturb else rays_o.new_zeros(n)
        _backend.march_rays_train(rays_o, rays_d, density_bitfield, bound, dt_gamma, max_steps, n, C, H, m,
                                  nears, fars, xyzs, dirs, deltas, rays, step_counter, noises)
        if not use_estimate:      # warm-up epochs only: one D2H read to trim the worst-case buffers
            used = _pad_to(int(step_counter[0].item()), align)
            xyzs, dirs, deltas = xyzs[:used], dirs[:used], deltas[:used]
        ctx.save_for_backward(rays, deltas)
        return xyzs, dirs, deltas, rays

    @staticmethod
    @_bwd
    def backward(ctx, grad_xyzs, grad_dirs, grad_deltas, grad_rays):
        rays, deltas = ctx.saved_tensors
        n, m = rays.shape[0], grad_xyzs.shape[0]
        g_o = torch.zeros(n, 3, device=rays.device)
        g_d = torch.zeros(n, 3, device=rays.device)
        _backend.march_rays_train_backward(grad_xyzs.contiguous(), grad_dirs.contiguous(), rays, deltas, n, m, g_o, g_d)
        return (g_o, g_d) + (None,) * 13


class _march_rays(Function):
    """raymarching.py:347-398 — inference march of the compacted alive rays, <= n_step samples each."""

    @staticmethod
    @_fwd32
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
                align=-1, perturb=False, dt_gamma=0, max_steps=1024):
        rays_o, rays_d = _rays3(rays_o), _rays3(rays_d)
        m = _pad_to(n_alive * n_step, align)
        # zero fill is part of the contract: deltas == 0 marks "ray ended" for the composite (raymarching.cu:982)
        # zero-filled like the reference's (raymarching.py:231-233), in ONE allocation / fill: rows past the counter stay zero
        zbuf = rays_o.new_zeros(m * 8)
        xyzs, dirs, deltas = zbuf[:3 * m].view(m, 3), zbuf[3 * m:6 * m].view(m, 3), zbuf[6 * m:].view(m, 2)
        noises = torch.rand(n_alive, dtype=rays_o.dtype, device=rays_o.device) if perturb else rays_o.new_zeros(n_alive)
        _backend.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
                            density_bitfield, near, far, xyzs, dirs, deltas, noises)
        return xyzs, dirs, deltas


march_rays_train = _march_rays_train.apply
march_rays = _march_rays.apply


# ----------------------------------------------------------------------------------------------------------------
# composites — one factory per family instead of nine hand-written classes
# ----------------------------------------------------------------------------------------------------------------
def _make_train_composite(tag, extras):
    """Training composite `composite_rays_train[_<tag>]` (raymarching.py:283-341, 442-500, 515-578, 594-660).

    forward(sigmas [M], rgbs [M,3], *extras [M], deltas [M,2], rays [N,3], T_thresh=1e-4)
        -> weights_sum [N], *extra_sums [N], depth [N], image [N,3]
    backward ignores grad_depth, like the reference (raymarching.py:316)."""
    k = len(extras)
    fwd_fn = getattr(_backend, f"composite_rays_train{'_' + tag if tag else ''}_forward")
    bwd_fn = getattr(_backend, f"composite_rays_train{'_' + tag if tag else ''}_backward")

    class _Composite(Function):
        @staticmethod
        @_fwd32
        def forward(ctx, sigmas, rgbs, *args):
            per_sample = [t.contiguous() for t in args[:k]]
            deltas, rays = args[k], args[k + 1]
            t_thresh = args[k + 2] if len(args) > k + 2 else 1e-4
            sigmas, rgbs = sigmas.contiguous(), rgbs.contiguous()
            m, n = sigmas.shape[0], rays.shape[0]
            weights_sum = sigmas.new_empty(n)
            sums = [sigmas.new_empty(n) for _ in range(k)]
            depth, image = sigmas.new_empty(n), sigmas.new_empty(n, 3)
            fwd_fn(sigmas, rgbs, *per_sample, deltas, rays, m, n, t_thresh, weights_sum, *sums, depth, image)
            ctx.save_for_backward(sigmas, rgbs, *per_sample, deltas, rays, weights_sum, *sums, image)
            ctx.dims = (m, n, t_thresh)
            return (weights_sum, *sums, depth, image)

        @staticmethod
        @_bwd
        def backward(ctx, grad_weights_sum, *grads):
            g_sums, grad_image = [g.contiguous() for g in grads[:k]], grads[k + 1].contiguous()   # grads[k] = grad_depth (unused)
            saved = ctx.saved_tensors
            sigmas, rgbs = saved[0], saved[1]
            per_sample, (deltas, rays, weights_sum) = saved[2:2 + k], saved[2 + k:5 + k]
            sums, image = saved[5 + k:5 + 2 * k], saved[5 + 2 * k]
            m, n, t_thresh = ctx.dims
            # pre-zeroed gradients (raymarching.py:627-633), one fill for all of them
            shapes = [sigmas.shape, rgbs.shape] + [t.shape for t in per_sample]
            sizes = [int(torch.Size(sh).numel()) for sh in shapes]
            zbuf = sigmas.new_zeros(sum(sizes))
            views, off = [], 0
            for sh, sz in zip(shapes, sizes):
                views.append(zbuf[off:off + sz].view(sh)); off += sz
            g_sigmas, g_rgbs, g_extra = views[0], views[1], views[2:]
            bwd_fn(grad_weights_sum.contiguous(), *g_sums, grad_image, sigmas, rgbs, *per_sample, deltas, rays,
                   weights_sum, *sums, image, m, n, t_thresh, g_sigmas, g_rgbs, *g_extra)
            return (g_sigmas, g_rgbs, *g_extra, None, None, None)

    _Composite.__name__ = f"_composite_rays_train{'_' + tag if tag else ''}"
    _Composite.__qualname__ = _Composite.__name__
    return _Composite


def _make_infer_composite(name, n_extra):
    """Inference composite (raymarching.py:401-434, 503-511, 581-589, 663-671): mutates the accumulators,
    `rays_alive` and `rays_t` in place and returns an empty tuple.

    forward(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, *extras, weights_sum, depth, image, *extra_sums, T_thresh=1e-2)"""
    fn = getattr(_backend, name)
    n_pos = 7 + n_extra + 3 + n_extra

    class _Composite(Function):
        @staticmethod
        @_fwd32
        def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, *args):
            t_thresh = args[n_pos - 7] if len(args) > n_pos - 7 else 1e-2
            fn(n_alive, n_step, t_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, *args[:n_pos - 7])
            return tuple()

    _Composite.__name__ = "_" + name
    _Composite.__qualname__ = _Composite.__name__
    return _Composite


_composite_rays_train = _make_train_composite("", ["ambient"])
_composite_rays_train_sigma = _make_train_composite("sigma", ["ambient"])
_composite_rays_train_uncertainty = _make_train_composite("uncertainty", ["ambient", "uncertainty"])
_composite_rays_train_triplane = _make_train_composite("triplane", ["amb_aud", "amb_eye", "uncertainty"])
_composite_rays = _make_infer_composite("composite_rays", 0)
_composite_rays_ambient = _make_infer_composite("composite_rays_ambient", 1)
_composite_rays_ambient_sigma = _make_infer_composite("composite_rays_ambient_sigma", 1)
_composite_rays_uncertainty = _make_infer_composite("composite_rays_uncertainty", 2)
_composite_rays_triplane = _make_infer_composite("composite_rays_triplane", 3)

composite_rays_train = _composite_rays_train.apply
composite_rays_train_sigma = _composite_rays_train_sigma.apply
composite_rays_train_uncertainty = _composite_rays_train_uncertainty.apply
composite_rays_train_triplane = _composite_rays_train_triplane.apply
composite_rays = _composite_rays.apply
composite_rays_ambient = _composite_rays_ambient.apply
composite_rays_ambient_sigma = _composite_rays_ambient_sigma.apply
composite_rays_uncertainty = _composite_rays_uncertainty.apply
composite_rays_triplane = _composite_rays_triplane.apply

__all__ = [
    "near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "morton3D_dilation",
    "march_rays_train", "composite_rays_train", "march_rays", "composite_rays", "composite_rays_ambient",
    "composite_rays_train_sigma", "composite_rays_ambient_sigma", "composite_rays_train_uncertainty",
    "composite_rays_uncertainty", "composite_rays_train_triplane", "composite_rays_triplane",
]

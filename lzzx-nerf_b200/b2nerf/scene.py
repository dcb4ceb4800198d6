"""Seeded synthetic inputs for the hot path: a head-sized density blob in the 128^3 occupancy grid,
a pinhole camera orbiting it, HuBERT-shaped audio windows and random-init head-model weights.
Pure numpy (float64 math, narrowed once) so the CPU oracle, the GPU path and both boxes see identical bits.
"""
import math

import numpy as np

GRID = 128
FOVY_DEG = 21.24          # train.py:97
RADIUS = 3.35             # train.py:96
AABB = np.array([-1.0, -0.5, -1.0, 1.0, 0.5, 1.0], np.float32)   # renderer.py:110 with bound=1
MIN_NEAR = 0.05           # train.py:77
DT_GAMMA = 1.0 / 256      # train.py:75
MAX_STEPS = 16            # train.py:35
BOUND = 1.0
DENSITY_THRESH = 10.0     # train.py:78


def _spread3(v):
    v = v.astype(np.uint64)
    v = (v * 0x00010001) & 0xFF0000FF
    v = (v * 0x00000101) & 0x0F00F00F
    v = (v * 0x00000011) & 0xC30C30C3
    v = (v * 0x00000005) & 0x49249249
    return v


def morton3d(x, y, z):
    return (_spread3(x) | (_spread3(y) << 1) | (_spread3(z) << 2)).astype(np.int64)


def density_grid(h=GRID, sigma=0.35, peak=20.0):
    """density[0, morton(x,y,z)] = peak * exp(-|p|^2 / (2 sigma^2)), p = voxel centre in [-1,1]^3 (Morton order, [1,h^3] fp32)."""
    c = (2.0 * np.arange(h, dtype=np.float64) + 1.0) / h - 1.0
    x, y, z = np.meshgrid(np.arange(h), np.arange(h), np.arange(h), indexing="ij")
    r2 = c[x] ** 2 + c[y] ** 2 + c[z] ** 2
    dens = peak * np.exp(-r2 / (2.0 * sigma * sigma))
    out = np.zeros(h ** 3, np.float32)
    out[morton3d(x.ravel(), y.ravel(), z.ravel())] = dens.ravel().astype(np.float32)
    return out.reshape(1, -1)


def bitfield_from_grid(grid, thresh=DENSITY_THRESH):
    """Reference packbits semantics (raymarching.cu:283-288): bit i of byte n = cell 8n+i, strict >."""
    bits = (grid.reshape(-1, 8) > np.float32(thresh)).astype(np.uint8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)


def camera_pose(frame=0, radius=RADIUS, jitter_deg=5.0, seed=0):
    """c2w [4,4]: camera on a sphere of `radius` looking at the origin, yaw/pitch jittered +-jitter_deg (seeded per frame)."""
    rng = np.random.default_rng(seed * 100003 + frame)
    yaw, pitch = np.deg2rad(rng.uniform(-jitter_deg, jitter_deg, 2))
    eye = radius * np.array([math.sin(yaw) * math.cos(pitch), math.sin(pitch), math.cos(yaw) * math.cos(pitch)])
    fwd = -eye / np.linalg.norm(eye)              # camera looks along +z of the camera frame (get_rays convention, utils.py:300)
    right = np.cross(np.array([0.0, 1.0, 0.0]), fwd)
    right /= np.linalg.norm(right)
    up = np.cross(fwd, right)
    pose = np.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, up, fwd, eye
    return pose


def rays_for_pixels(pose, H, W, pix_i, pix_j, fovy_deg=FOVY_DEG):
    """get_rays semantics (nerf_triplane/utils.py:227-312): pixel centres +0.5, zs = 1, normalised directions. Returns fp32 [N,3] x2."""
    focal = H / (2.0 * math.tan(math.radians(fovy_deg) / 2.0))
    cx, cy = W / 2.0, H / 2.0
    xs = (pix_i.astype(np.float64) + 0.5 - cx) / focal
    ys = (pix_j.astype(np.float64) + 0.5 - cy) / focal
    d = np.stack([xs, ys, np.ones_like(xs)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    d = d @ pose[:3, :3].T
    o = np.broadcast_to(pose[:3, 3], d.shape)
    return np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32)


def intrinsics(H=512, W=512, fovy_deg=FOVY_DEG):
    """(fx, fy, cx, cy) of the synthetic pinhole camera used by rays_for_pixels."""
    focal = H / (2.0 * math.tan(math.radians(fovy_deg) / 2.0))
    return focal, focal, W / 2.0, H / 2.0


def frame_rays(frame=0, H=512, W=512, seed=0):
    j, i = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    return rays_for_pixels(camera_pose(frame, seed=seed), H, W, i.ravel(), j.ravel())


def train_rays(step=0, n=65536, H=512, W=512, seed=0):
    rng = np.random.default_rng(seed * 7919 + step + 1)
    idx = rng.integers(0, H * W, n)           # duplicates allowed, like provider.py:669
    return rays_for_pixels(camera_pose(step, seed=seed), H, W, idx % W, idx // W)


def audio_window(frame=0, hubert=True, seed=0):
    rng = np.random.default_rng(seed * 31337 + frame + 17)
    return rng.standard_normal((8, 1024, 2) if hubert else (8, 29, 16)).astype(np.float32)

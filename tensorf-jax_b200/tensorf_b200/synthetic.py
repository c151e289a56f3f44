"""Seeded synthetic inputs of the reference's shapes and distributions.

There is no network for datasets or checkpoints, so benches and parity tests use random-init
factors (`training.py:44`, `:69-82`: N(0, 0.1^2)), MLP leaves with the variances
`networks.py:9-10` states, and rays drawn like the lego (Blender, pinhole) and dozer
(nerfstudio, poses normalised into [-1,1]^3) datasets.  Everything is NumPy fp32 so the very
same arrays can be handed to the CUDA path and to the CPU oracle.
Seeds: 0 parameters, 1 rays, 2 noise.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional, Tuple

import numpy as np

LEGO_AABB = np.array([[-0.6585, -1.1833, -0.4651], [0.6636, 1.1929, 1.0512]], dtype=np.float32)  # train_lego.py:28-29
DOZER_AABB = np.array([[-2.0, -2.0, -2.0], [2.0, 2.0, 2.0]], dtype=np.float32)  # train_nerfstudio.py:29-30


@dataclasses.dataclass(frozen=True)
class Workload:
    """One BASELINE.json configuration as concrete shapes."""

    name: str
    R: int
    G: int
    cd: int
    ca: int
    N: int
    K: int
    feat_freqs: int
    view_freqs: int
    contracted: bool = False
    num_cameras: Optional[int] = None
    near: float = 0.05
    far: float = 200.0

    def aabb(self) -> np.ndarray:
        return DOZER_AABB if self.contracted else LEGO_AABB

    def encoded_dim(self, squash: int = 27) -> int:
        return squash + 3 + 2 * self.feat_freqs * squash + 2 * self.view_freqs * 3

    def fwd_bytes_per_ray(self) -> int:
        """Gather-model algorithmic bytes: 6 taps x 4 B per sample-channel."""
        return 24 * (self.N * 3 * self.cd + self.K * 3 * self.ca)

    def train_bytes_per_ray(self) -> int:
        return 3 * self.fwd_bytes_per_ray()


def sample_counts(grid_dim: int, multiplier: float = 1.0) -> Tuple[int, int]:
    """training.py:115-118."""
    n = int(math.sqrt(3 * grid_dim**2) * multiplier)
    return n, int(0.15 * n)


def lego_workload(R: int = 4096, G: int = 128, N: Optional[int] = None, K: Optional[int] = None, name: str = "") -> Workload:
    n, k = sample_counts(G)
    N = n if N is None else N
    K = k if K is None else K
    return Workload(name or f"lego_G{G}_R{R}_N{N}_K{K}", R, G, 16, 48, N, K, 2, 2)


def dozer_workload(R: int = 2048, G: int = 128, ncam: int = 256) -> Workload:
    n, k = sample_counts(G, 3.0)
    return Workload(f"dozer_G{G}_R{R}_N{n}_K{k}", R, G, 32, 48, n, k, 6, 6, contracted=True, num_cameras=ncam)


def render360_workload(R: int = 16384, G: int = 300) -> Workload:
    """render_360.py:43-51 with the BASELINE sizes: N=512, K=128, chunks of 16384 rays."""
    return Workload(f"render360_G{G}_R{R}_N512_K128", R, G, 16, 48, 512, 128, 2, 2)


def make_params(G: int, cd: int, ca: int, feat_freqs: int, view_freqs: int, num_cameras: Optional[int] = None,
                seed: int = 0, squash: int = 27, units: int = 128, bias_std: float = 0.0) -> Dict[str, np.ndarray]:
    """Leaves of `LearnableParams` (render.py:39-46) in the reference's layouts."""
    rng = np.random.default_rng(seed)
    enc = squash + 3 + 2 * feat_freqs * squash + 2 * view_freqs * 3

    def n(shape, std):
        return rng.normal(0.0, std, size=shape).astype(np.float32)

    p = {
        "density_vector": n((3, cd, G), 0.1),
        "density_matrix": n((3, cd, G, G), 0.1),
        "appearance_vector": n((3, ca, G), 0.1),
        "appearance_matrix": n((3, ca, G, G), 0.1),
        "w0": n((3 * ca, squash), math.sqrt(1.0 / (3 * ca))),
        "w1": n((enc, units), math.sqrt(2.0 / enc)),
        "b1": n((units,), bias_std) if bias_std > 0 else np.zeros(units, np.float32),
        "w2": n((units, units), math.sqrt(2.0 / units)),
        "b2": n((units,), bias_std) if bias_std > 0 else np.zeros(units, np.float32),
        "w3": n((units, 3), math.sqrt(1.0 / units)),
        "b3": n((3,), bias_std) if bias_std > 0 else np.zeros(3, np.float32),
    }
    if num_cameras is not None:
        p["embed"] = n((num_cameras, units), math.sqrt(1.0 / units))
    return p


def _look_at_rotation(origin: np.ndarray) -> np.ndarray:
    """R_world_camera (…,3,3) for cameras at `origin` looking at the world origin, +z forward
    (OpenCV convention, as produced by data.py's T_blendercam_camera flip), world up = +z."""
    fwd = -origin / np.linalg.norm(origin, axis=-1, keepdims=True)
    up = np.broadcast_to(np.array([0.0, 0.0, 1.0]), fwd.shape)
    right = np.cross(fwd, up)
    right /= np.linalg.norm(right, axis=-1, keepdims=True) + 1e-12
    down = np.cross(fwd, right)
    return np.stack([right, down, fwd], axis=-1)


def lego_rays(R: int, seed: int = 1, width: int = 800, height: int = 800, fov_x: float = 0.6911, radius: float = 4.0311,
              num_cameras: int = 100):
    """Training-like rays: shuffled across images (training.py:318-322), i.e. no inter-ray
    coherence.  Pinhole model of cameras.py:44-76, :100-121."""
    rng = np.random.default_rng(seed)
    # uniform on the upper hemisphere
    v = rng.normal(size=(R, 3))
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    v[:, 2] = np.abs(v[:, 2]) * 0.9 + 0.05
    v /= np.linalg.norm(v, axis=-1, keepdims=True)
    origins = v * radius
    rot = _look_at_rotation(origins)
    fx = (width / 2.0) / math.tan(fov_x / 2.0)
    cx, cy = width / 2.0 - 0.5, height / 2.0 - 0.5
    u = rng.integers(0, width, size=R).astype(np.float64)
    w = rng.integers(0, height, size=R).astype(np.float64)
    d_cam = np.stack([(u - cx) / fx, (w - cy) / fx, np.ones(R)], axis=-1)
    d = np.einsum("rij,rj->ri", rot, d_cam)
    d /= np.linalg.norm(d, axis=-1, keepdims=True) + 1e-8
    cams = rng.integers(0, num_cameras, size=R).astype(np.uint32)
    return origins.astype(np.float32), d.astype(np.float32), cams


def frame_rays(width: int, height: int, angle: float = 0.3, elevation: float = 0.5, fov_x: float = 0.6911,
               radius: float = 4.0311, rows: Optional[Tuple[int, int]] = None):
    """Raster-ordered rays of one render_360-style frame (render_360.py:118-170,
    cameras.py:124-143); `rows=(r0, r1)` returns one image tile (row band)."""
    o = radius * np.array([math.cos(angle) * math.cos(elevation), math.sin(angle) * math.cos(elevation), math.sin(elevation)])
    rot = _look_at_rotation(o[None])[0]
    fx = (width / 2.0) / math.tan(fov_x / 2.0)
    cx, cy = width / 2.0 - 0.5, height / 2.0 - 0.5
    r0, r1 = rows if rows is not None else (0, height)
    vv, uu = np.mgrid[r0:r1, :width]
    d_cam = np.stack([(uu - cx) / fx, (vv - cy) / fx, np.ones_like(uu, dtype=np.float64)], axis=-1).reshape(-1, 3)
    d = d_cam @ rot.T
    d /= np.linalg.norm(d, axis=-1, keepdims=True) + 1e-8
    n = d.shape[0]
    return (np.broadcast_to(o.astype(np.float32), (n, 3)).copy(), d.astype(np.float32), np.zeros(n, np.uint32))


def frame_camera(width: int, height: int, angle: float = 0.3, elevation: float = 0.5, fov_x: float = 0.6911, radius: float = 4.0311):
    """The `cameras.Camera` whose pixel rays are `frame_rays` (one pose of a render_360-style orbit)."""
    from .cameras import Camera
    o = radius * np.array([math.cos(angle) * math.cos(elevation), math.sin(angle) * math.cos(elevation), math.sin(elevation)])
    rot = _look_at_rotation(o[None])[0]  # R_world_camera
    T = np.eye(4)
    T[:3, :3] = rot.T
    T[:3, 3] = -rot.T @ o
    return Camera.from_fov(T.astype(np.float32), width, height, fov_x_radians=fov_x)


def dozer_rays(R: int, ncam: int = 256, seed: int = 1):
    """Unbounded-scene rays: origins uniform in [-1,1]^3 (poses are normalised into that cube,
    data.py:167-181), forward axis toward the scene centre plus a fisheye-like direction inside
    a 90-degree half-angle cone, unit norm (data.py:106-120)."""
    rng = np.random.default_rng(seed)
    origins = rng.uniform(-1.0, 1.0, size=(R, 3))
    origins[np.linalg.norm(origins, axis=-1) < 0.05] += 0.1
    rot = _look_at_rotation(origins)
    theta = np.arccos(rng.uniform(0.0, 1.0, size=R))  # uniform on the hemisphere cap (<= 90 deg)
    phi = rng.uniform(0.0, 2 * math.pi, size=R)
    d_cam = np.stack([np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)], axis=-1)
    d = np.einsum("rij,rj->ri", rot, d_cam)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    cams = rng.integers(0, ncam, size=R).astype(np.uint32)
    return origins.astype(np.float32), d.astype(np.float32), cams


def make_colors(R: int, seed: int = 1) -> np.ndarray:
    return np.random.default_rng(seed + 1000).uniform(0.0, 1.0, size=(R, 3)).astype(np.float32)


def make_noise(N: int, R: int, contracted: bool, seed: int = 2):
    """jitter U[0,1): (N,) shared by all rays for the bounded scene (render.py:177-183,
    :373-379) or (R,N) for the contracted scene (render.py:158-161); gumbel (N,) shared by all
    rays (render.py:461-469) = -log(-log U), U in [tiny, 1)."""
    rng = np.random.default_rng(seed)
    jitter = rng.uniform(0.0, 1.0, size=(R, N) if contracted else (N,)).astype(np.float32)
    jitter = np.minimum(jitter, np.nextafter(np.float32(1.0), np.float32(0.0)))
    tiny = np.finfo(np.float32).tiny
    u = np.maximum(rng.uniform(0.0, 1.0, size=(N,)).astype(np.float32), tiny)
    u = np.minimum(u, np.nextafter(np.float32(1.0), np.float32(0.0)))
    gumbel = (-np.log(-np.log(u.astype(np.float64)))).astype(np.float32)
    return jitter, gumbel


def make_inputs(w: Workload, seed_params: int = 0, seed_rays: int = 1, seed_noise: int = 2, R: Optional[int] = None,
                bias_std: float = 0.0):
    """Everything one hot-path invocation needs, as NumPy arrays."""
    R = w.R if R is None else R
    params = make_params(w.G, w.cd, w.ca, w.feat_freqs, w.view_freqs, w.num_cameras, seed_params, bias_std=bias_std)
    if w.contracted:
        o, d, cams = dozer_rays(R, w.num_cameras or 256, seed_rays)
    else:
        o, d, cams = lego_rays(R, seed_rays)
    jitter, gumbel = make_noise(w.N, R, w.contracted, seed_noise)
    return dict(params=params, aabb=w.aabb().copy(), origins=o, directions=d, camera_indices=cams,
                colors=make_colors(R, seed_rays), jitter=jitter, gumbel=gumbel)

"""Mirror of tensorf/render.py (render_rays, render_rays_batched, RenderConfig, RenderMode,
LearnableParams, compute_segment_probabilities) over the CUDA library.

Same call signatures and argument meaning as the reference; arrays are torch CUDA tensors
instead of JAX arrays (JAX is not in this image — the JAX binding in `jax_ffi/` wraps the very
same C ABI).  `render_rays` is differentiable w.r.t. every leaf of `LearnableParams` through
torch.autograd, the analogue of the `jax.custom_vjp` that keeps `jax.value_and_grad`
(training.py:153-156) working.
"""
from __future__ import annotations

import dataclasses
import enum
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import cameras, networks, ops, prng, tensor_vm


class RenderMode(enum.Enum):  # render.py:18-23
    RGB = enum.auto()
    DIST_MEDIAN = enum.auto()
    DIST_MEAN = enum.auto()


_MODE_CODE = {RenderMode.RGB: ops.MODE_RGB, RenderMode.DIST_MEDIAN: ops.MODE_DIST_MEDIAN, RenderMode.DIST_MEAN: ops.MODE_DIST_MEAN}


@dataclasses.dataclass(frozen=True)
class RenderConfig:  # render.py:26-36
    near: float
    far: float
    mode: RenderMode
    density_samples_per_ray: int
    appearance_samples_per_ray: int


@dataclasses.dataclass
class LearnableParams:  # render.py:39-46
    appearance_mlp_params: Dict
    appearance_tensor: tensor_vm.TensorVM
    density_tensor: tensor_vm.TensorVM
    scene_contraction: bool = False

    def flat(self) -> Dict[str, torch.Tensor]:
        """Leaves under the flat names of `struct tensorf_params`."""
        out = {
            "density_vector": self.density_tensor.stacked_single_vm.vector,
            "density_matrix": self.density_tensor.stacked_single_vm.matrix,
            "appearance_vector": self.appearance_tensor.stacked_single_vm.vector,
            "appearance_matrix": self.appearance_tensor.stacked_single_vm.matrix,
        }
        out.update(networks.flatten_mlp_params(self.appearance_mlp_params))
        return out

    @staticmethod
    def from_flat(flat: Dict[str, torch.Tensor], scene_contraction: bool) -> "LearnableParams":
        return LearnableParams(
            appearance_mlp_params=networks.unflatten_mlp_params(flat),
            appearance_tensor=tensor_vm.TensorVM(tensor_vm.TensorVMSingle(flat["appearance_vector"], flat["appearance_matrix"])),
            density_tensor=tensor_vm.TensorVM(tensor_vm.TensorVMSingle(flat["density_vector"], flat["density_matrix"])),
            scene_contraction=scene_contraction)


from .schedule import contracted_schedule  # noqa: E402,F401  (numpy only; shared with the JAX binding)


# ---- workspace pool: one RenderCall per (shape, device), never shared by two live graphs ------
_POOL: Dict[tuple, list] = {}


def _acquire(desc, device) -> ops.RenderCall:
    key = (tuple(getattr(desc, f) for f, _ in desc._fields_ if f != "loss_scale"), str(device))
    # a new grid size (every upsampling step changes G, and with it N and K) retires the workspaces of the old one: they
    # are 1-8 GB each and would otherwise stay cached for the life of the process
    for k in [k for k in _POOL if k[1] == str(device) and k[0][3] != desc.G]:
        del _POOL[k]
    pool = _POOL.setdefault(key, [])
    call = pool.pop() if pool else ops.RenderCall(desc, device)
    call.desc = desc
    call._pool_key = key
    return call


def _release(call: ops.RenderCall) -> None:
    pool = _POOL.setdefault(call._pool_key, [])
    if len(pool) < 2:
        pool.append(call)


class _RenderRgb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, desc, inputs, names, *leaves):
        call = _acquire(desc, leaves[0].device)
        rgb, _ = call.forward({n: t.contiguous() for n, t in zip(names, leaves)}, inputs)
        ctx.call, ctx.names = call, names
        return rgb

    @staticmethod
    def backward(ctx, d_rgb):
        grads = ctx.call.backward(d_rgb.contiguous())
        _release(ctx.call)
        return (None, None, None) + tuple(grads[n] for n in ctx.names)


def _device_noise(noise: prng.RenderNoise, device) -> Dict[str, torch.Tensor]:
    def dv(x):
        return x.to(device=device, dtype=torch.float32) if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device)
    out = {"jitter": dv(noise.jitter)}
    if noise.gumbel is not None:
        out["gumbel"] = dv(noise.gumbel)
    return out


def _noise_inputs(prng_key, ray_count: int, density_samples: int, contracted: bool, need_gumbel: bool, device,
                  first_ray: int = 0) -> Dict[str, torch.Tensor]:
    """The jitter / Gumbel input arrays of the C ABI for a key.  A `prng.Key` is expanded on the device
    (`tensorf_prng_uniform/gumbel`: no host draw, no H2D copy of the (R,N) jitter); explicit `RenderNoise`
    arrays and JAX keys take the host route."""
    if isinstance(prng_key, prng.Key):
        d = prng.render_noise_device(prng_key, ray_count, density_samples, contracted, device, need_gumbel, first_ray=first_ray)
        return {k: v for k, v in d.items() if v is not None}
    return _device_noise(prng.render_noise(prng_key, ray_count, density_samples, contracted, need_gumbel=need_gumbel), device)


def render_rays(appearance_mlp: networks.FeatureMlp, learnable_params: LearnableParams, aabb: torch.Tensor,
                rays_wrt_world: cameras.Rays3D, prng_key, config: RenderConfig) -> torch.Tensor:
    """render.py:105-279. Output (ray_count, 3) for RGB, (ray_count,) for the distance modes."""
    (ray_count,) = rays_wrt_world.get_batch_axes()
    device = aabb.device
    if tuple(aabb.shape) != (2, 3):
        raise ValueError(f"aabb must be (2,3), got {tuple(aabb.shape)}")
    flat = learnable_params.flat()
    cd, G = flat["density_vector"].shape[1], flat["density_vector"].shape[2]
    ca = flat["appearance_vector"].shape[1]
    if flat["appearance_vector"].shape[2] != G:
        raise ValueError("density and appearance grids must have the same dimension")  # training.py:114
    N, K = config.density_samples_per_ray, config.appearance_samples_per_ray
    contracted = bool(learnable_params.scene_contraction)
    mode = _MODE_CODE[config.mode]
    desc = ops.make_desc(R=ray_count, N=N, K=K if config.mode is RenderMode.RGB else 1, G=G, cd=cd, ca=ca, mode=mode,
                         contracted=contracted, feat_freqs=appearance_mlp.feature_n_freqs,
                         view_freqs=appearance_mlp.viewdir_n_freqs, num_cameras=appearance_mlp.num_cameras,
                         squash=appearance_mlp.feature_squash_dim, units=appearance_mlp.units)
    inputs = {"origins": rays_wrt_world.origins.to(torch.float32).contiguous(),
              "directions": rays_wrt_world.directions.to(torch.float32).contiguous(),
              "camera_indices": rays_wrt_world.camera_indices.to(torch.int32).contiguous(),
              "aabb": aabb.to(torch.float32).contiguous()}
    inputs.update(_noise_inputs(prng_key, ray_count, N, contracted, config.mode is RenderMode.RGB, device))
    if contracted:
        base, delta = contracted_schedule(config.near, config.far, N)
        inputs["base_ts"] = torch.from_numpy(base).to(device)
        inputs["deltas"] = torch.from_numpy(delta).to(device)
    if config.mode is RenderMode.RGB:
        names = tuple(n for n in ops.param_shapes(desc))
        if not (torch.is_grad_enabled() and any(flat[n].requires_grad for n in names)):
            # nothing will be differentiated (render_360.py, validation): forward-only call, no residuals kept
            desc.flags = ops.FLAG_INFERENCE
            call = _acquire(desc, device)
            rgb, _ = call.forward({n: flat[n].detach().contiguous() for n in names}, inputs)
            _release(call)
            return rgb
        return _RenderRgb.apply(desc, inputs, names, *[flat[n] for n in names])
    call = _acquire(desc, device)
    out = call.depth({k: flat[k].contiguous() for k in ("density_vector", "density_matrix")}, inputs)
    _release(call)
    return out


def render_rays_batched(appearance_mlp: networks.FeatureMlp, learnable_params: LearnableParams, aabb: torch.Tensor,
                        rays_wrt_world: cameras.Rays3D, prng_key, config: RenderConfig, *, batch_size: int = 4096,
                        use_tqdm: bool = False) -> np.ndarray:
    """render.py:49-102: arbitrary batch axes, chunked, concatenated on the host as numpy.
    Every chunk uses the same key (same shared jitter/Gumbel vectors), as the reference does."""
    batch_axes = rays_wrt_world.get_batch_axes()
    rays = rays_wrt_world.reshape(-1).to(aabb.device)
    (total,) = rays.get_batch_axes()
    contracted = bool(learnable_params.scene_contraction)
    out = []
    it = range(math.ceil(total / batch_size))
    if use_tqdm:
        from tqdm.auto import tqdm
        it = tqdm(it)
    shared = None
    if not contracted and not isinstance(prng_key, prng.RenderNoise):
        shared = prng.render_noise(prng_key, batch_size, config.density_samples_per_ray, False,
                                   need_gumbel=config.mode is RenderMode.RGB)
    with torch.no_grad():
        for i in it:
            a, b = i * batch_size, min(total, (i + 1) * batch_size)
            out.append(render_rays(appearance_mlp, learnable_params, aabb, rays.slice(a, b),
                                   shared if shared is not None else prng_key, config))
    res = torch.cat(out, dim=0).cpu().numpy()
    return res.reshape(batch_axes + res.shape[1:])


class FrameRenderer:
    """One rank's share of render_360.py:118-161 (a camera path of full frames through render.py:49-102), kept on the
    device from pixel to pixel: `tensorf_pixel_rays` fills this rank's rows of the frame's ray table, the table is
    rendered in chunks of `batch_size` rays (the last chunk is ragged, as in the reference's chunk loop), every chunk
    writes straight into the rank's slice of a device frame buffer, and ONE device-to-host copy per frame lands it
    in pinned memory.  Rows are dealt to ranks as interleaved stripes (`dist.stripe_rows`); there is no collective:
    rank r's rows of the final image are `rows` and the host just places them (`scatter_into`).
    Every chunk uses the same jitter / Gumbel vectors (same key), as render_rays_batched does."""

    def __init__(self, appearance_mlp: networks.FeatureMlp, learnable_params: LearnableParams, aabb: torch.Tensor, config: RenderConfig,
                 image_height: int, image_width: int, *, batch_size: int = 16384, rank: int = 0, world: int = 1,
                 stripe: Optional[int] = None):
        from . import dist as tdist
        if stripe is None:  # the largest stripe (<= 16 rows) that deals every rank the same number of rows
            stripe = tdist.balanced_stripe(image_height, world)
        self.stripe, self.rank, self.world = stripe, rank, world
        self.config, self.aabb, self.device = config, aabb.to(torch.float32).contiguous(), aabb.device
        self.H, self.W, self.batch = image_height, image_width, batch_size
        self.rows: List[Tuple[int, int]] = tdist.stripe_rows(image_height, rank, world, stripe)
        self.n = sum(b - a for a, b in self.rows) * image_width
        flat = learnable_params.flat()
        if bool(learnable_params.scene_contraction):
            raise NotImplementedError("FrameRenderer: per-ray jitter of contracted scenes goes through render_rays_batched")
        self.rgb = config.mode is RenderMode.RGB
        names = tuple(ops.param_shapes(self._desc(appearance_mlp, flat, 1))) if self.rgb else ("density_vector", "density_matrix")
        self.params = {k: flat[k].detach().contiguous() for k in names}
        self.calls: Dict[int, ops.RenderCall] = {}
        for R in {min(batch_size, self.n), self.n % batch_size} - {0}:
            self.calls[R] = ops.RenderCall(self._desc(appearance_mlp, flat, R), self.device)
        dev = self.device
        self.o = torch.empty((self.n, 3), dtype=torch.float32, device=dev)
        self.d = torch.empty((self.n, 3), dtype=torch.float32, device=dev)
        self.c = torch.empty((self.n,), dtype=torch.int32, device=dev)
        shape = (self.n, 3) if self.rgb else (self.n,)
        self.frame_dev = torch.empty(shape, dtype=torch.float32, device=dev)
        self.frame_host = torch.empty(shape, dtype=torch.float32, pin_memory=True)

    def _desc(self, mlp, flat, R):
        c = self.config
        return ops.make_desc(R=R, N=c.density_samples_per_ray, K=c.appearance_samples_per_ray if c.mode is RenderMode.RGB else 1,
                             G=flat["density_vector"].shape[2], cd=flat["density_vector"].shape[1], ca=flat["appearance_vector"].shape[1],
                             mode=_MODE_CODE[c.mode], feat_freqs=mlp.feature_n_freqs, view_freqs=mlp.viewdir_n_freqs,
                             num_cameras=mlp.num_cameras, squash=mlp.feature_squash_dim, units=mlp.units, inference=True)

    def render(self, camera: cameras.Camera, camera_index: int, noise: Dict[str, torch.Tensor], sync: bool = True) -> torch.Tensor:
        """Renders this rank's rows of one frame; returns the pinned host tensor (n, 3) or (n,) in `rows` order.
        `noise`: device tensors 'jitter' (N,) and, for RGB, 'gumbel' (N,)."""
        M, origin = camera.ray_matrices()
        # this rank's stripes of the pixel grid -> one contiguous ray table, one launch
        ops.pixel_rays_striped(M.reshape(-1), origin, self.W, self.H, self.stripe, self.rank, self.world, int(camera_index),
                               (self.o, self.d, self.c))
        for a in range(0, self.n, self.batch):
            b = min(self.n, a + self.batch)
            call = self.calls[b - a]
            ins = dict(noise, aabb=self.aabb, origins=self.o[a:b], directions=self.d[a:b], camera_indices=self.c[a:b])
            if self.rgb:
                call.forward(self.params, ins, out=self.frame_dev[a:b])
            else:
                call.depth(self.params, ins, out=self.frame_dev[a:b])
        self.frame_host.copy_(self.frame_dev, non_blocking=True)
        if sync:
            torch.cuda.current_stream().synchronize()
        return self.frame_host

    def scatter_into(self, image: np.ndarray, part: Optional[torch.Tensor] = None) -> np.ndarray:
        """Places this rank's rows into a full (H, W[, 3]) image (host)."""
        src = (part if part is not None else self.frame_host).numpy()
        a = 0
        for r0, r1 in self.rows:
            b = a + (r1 - r0) * self.W
            image[r0:r1] = src[a:b].reshape((r1 - r0, self.W) + src.shape[1:])
            a = b
        return image


@dataclasses.dataclass
class SegmentProbabilities:  # render.py:282-297
    p_exits: torch.Tensor
    p_terminates: torch.Tensor

    def get_batch_axes(self) -> Tuple[int, ...]:
        return tuple(self.p_exits.shape)


def compute_segment_probabilities(sigmas: torch.Tensor, step_sizes: torch.Tensor) -> SegmentProbabilities:
    """render.py:300-347, arbitrary leading batch axes."""
    if sigmas.shape != step_sizes.shape:
        raise ValueError("sigmas and step_sizes must have the same shape")
    shape = sigmas.shape
    pe, pt = ops.segment_probabilities(sigmas.reshape(-1, shape[-1]).contiguous().float(),
                                       step_sizes.reshape(-1, shape[-1]).contiguous().float())
    return SegmentProbabilities(pe.reshape(shape), pt.reshape(shape))

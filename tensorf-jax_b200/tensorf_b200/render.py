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
from typing import Dict, Optional, Tuple

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


def contracted_schedule(near: float, far: float, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """Host constants of the contracted branch (render.py:127-155): the `ts` schedule and step
    sizes, computed as the reference does (close_ts fp32 linspace, far_ts float64 numpy)."""
    nc = n // 2
    nf = n - nc
    f32 = np.float32
    if nc > 1:  # jnp.linspace [upstream]: start*(1-s_i) + stop*s_i, s_i = i/(num-1); endpoint appended exactly
        sv = np.arange(nc - 1, dtype=np.float32) * f32(1.0 / (nc - 1))
        close = (f32(near) * (f32(1.0) - sv) + f32(near + 1.0) * sv).astype(np.float32)
        close = np.concatenate([close, np.array([near + 1.0], dtype=np.float32)])
    else:
        close = np.full((nc,), near, dtype=np.float32)
    far_start = near + 1.0 + 1.0 / nc
    k = 10.0
    far_deltas = (1.0 / (1.0 - np.linspace(0.0, 1.0 - 1 / ((far - far_start) / k + 1), nf)) - 1.0) * np.linspace(1.0, k, nf)
    base = np.concatenate([close, (far_start + far_deltas).astype(np.float32)]).astype(np.float32)
    delta = np.roll(base, -1) - base
    delta[-1] = delta[-2]
    return base, delta.astype(np.float32)


# ---- workspace pool: one RenderCall per (shape, device), never shared by two live graphs ------
_POOL: Dict[tuple, list] = {}


def _acquire(desc, device) -> ops.RenderCall:
    key = (tuple(getattr(desc, f) for f, _ in desc._fields_ if f != "loss_scale"), str(device))
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


def _noise_inputs(prng_key, ray_count: int, density_samples: int, contracted: bool, need_gumbel: bool, device) -> Dict[str, torch.Tensor]:
    """The jitter / Gumbel input arrays of the C ABI for a key.  A `prng.Key` is expanded on the device
    (`tensorf_prng_uniform/gumbel`: no host draw, no H2D copy of the (R,N) jitter); explicit `RenderNoise`
    arrays and JAX keys take the host route."""
    if isinstance(prng_key, prng.Key):
        d = prng.render_noise_device(prng_key, ray_count, density_samples, contracted, device, need_gumbel)
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

"""JAX binding of the hot path: keeps `tensorf.render.render_rays`' signature and makes
`jax.value_and_grad` (training.py:153-156) work through a `jax.custom_vjp`.

Import-guarded: JAX is not installable in this image, so this module is exercised only where
JAX and the compiled shim (xla_ffi_shim.cc) exist.  It is a thin layer: every computation is
one `jax.ffi.ffi_call` into libtensorf_b200.so; randomness is drawn in JAX exactly as
render.py:120, :158-160, :375-379 and :461-468 do and handed to the kernels as arrays.
"""
from __future__ import annotations

import ctypes
import pathlib

try:
    import jax
    import jax.numpy as jnp
    import numpy as onp
except ImportError as e:  # pragma: no cover
    raise ImportError("tensorf_jax needs JAX; in this image use tensorf_b200.render (torch front end)") from e

_HERE = pathlib.Path(__file__).resolve().parent
_shim = ctypes.CDLL(str(_HERE / "libtensorf_b200_xla.so"))
_core = ctypes.CDLL(str(_HERE.parent / "tensorf_b200" / "libtensorf_b200.so"))
jax.ffi.register_ffi_target("tensorf_render_rgb_fwd", jax.ffi.pycapsule(_shim.TensorfRenderRgbFwd), platform="CUDA")


def _leaves(lp):
    m = lp.appearance_mlp_params["params"]
    embed = m["Embed_0"]["embedding"] if "Embed_0" in m else jnp.zeros((1, 128), jnp.float32)
    return (m["Dense_0"]["kernel"], m["Dense_1"]["bias"], m["Dense_1"]["kernel"], m["Dense_2"]["bias"], m["Dense_2"]["kernel"],
            m["Dense_3"]["bias"], m["Dense_3"]["kernel"], embed,
            lp.appearance_tensor.stacked_single_vm.vector, lp.appearance_tensor.stacked_single_vm.matrix,
            lp.density_tensor.stacked_single_vm.vector, lp.density_tensor.stacked_single_vm.matrix)


def render_rays_rgb(appearance_mlp, learnable_params, aabb, rays_wrt_world, prng_key, config):
    """Drop-in for tensorf.render.render_rays (RGB mode), forward pass."""
    from tensorf_b200 import render as host  # host constants (contracted schedule) shared with the torch front end

    R = rays_wrt_world.origins.shape[0]
    N, K = config.density_samples_per_ray, config.appearance_samples_per_ray
    contracted = bool(learnable_params.scene_contraction)
    k_sample, k_rgb = jax.random.split(prng_key)                       # render.py:120
    jitter = jax.random.uniform(k_sample, (R, N) if contracted else (N,))   # :158-160 / :375-379
    gumbel = jax.random.gumbel(k_rgb, (N,))                             # :462-468 (random.choice's noise)
    base, delta = host.contracted_schedule(config.near, config.far, N) if contracted else (onp.zeros(N, onp.float32),) * 2
    leaves = _leaves(learnable_params)
    desc = _core_desc(R, N, K, leaves, contracted, appearance_mlp)
    nbytes = ctypes.c_int64()
    _core.tensorf_render_workspace_bytes(ctypes.byref(desc), ctypes.byref(nbytes))
    rgb, _ws = jax.ffi.ffi_call(
        "tensorf_render_rgb_fwd",
        (jax.ShapeDtypeStruct((R, 3), jnp.float32), jax.ShapeDtypeStruct((nbytes.value,), jnp.uint8)),
    )(*leaves, aabb, rays_wrt_world.origins, rays_wrt_world.directions, rays_wrt_world.camera_indices, jitter, gumbel,
      jnp.asarray(base), jnp.asarray(delta), density_samples_per_ray=N, appearance_samples_per_ray=K,
      scene_contraction=int(contracted), feature_n_freqs=appearance_mlp.feature_n_freqs,
      viewdir_n_freqs=appearance_mlp.viewdir_n_freqs, num_cameras=appearance_mlp.num_cameras or 0)
    return rgb


def _core_desc(R, N, K, leaves, contracted, mlp):
    from tensorf_b200 import ops
    den_mat, app_mat = leaves[11], leaves[9]
    return ops.make_desc(R=R, N=N, K=K, G=den_mat.shape[2], cd=den_mat.shape[1], ca=app_mat.shape[1], contracted=contracted,
                         feat_freqs=mlp.feature_n_freqs, view_freqs=mlp.viewdir_n_freqs, num_cameras=mlp.num_cameras)

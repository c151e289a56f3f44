"""JAX binding of the hot path over libtensorf_b200.so: `render_rays` with the signature of tensorf.render.render_rays
(render.py:105-113) for all three RenderModes, differentiable w.r.t. every leaf of LearnableParams through a
`jax.custom_vjp` (what training.py:153-156 `jax.value_and_grad` needs), `interpolate` (TensorVM.interpolate,
tensor_vm.py:42-89, also a custom_vjp) and `adam_step` (training.py:158-243).

Every computation is ONE `jax.ffi.ffi_call` into a handler of xla_ffi_shim.cc; randomness is drawn in JAX exactly as
render.py:120, :158-160, :375-379 and :461-468 do and handed to the kernels as arrays, so keys stay JAX keys.
The workspace the C ABI asks for is an extra result of the forward call and the residual of the custom_vjp.

Import-guarded and torch-free.  JAX is not installable in the image this repository is developed in: this module has
never been executed there (tests/test_jax_ffi_syntax.py only compiles it and checks the shim's handler signatures
against the C ABI); it runs where JAX and the compiled shim exist.
"""
from __future__ import annotations

import ctypes
import functools
import importlib.util
import pathlib

try:
    import jax
    import jax.numpy as jnp
    import numpy as onp
except ImportError as e:  # pragma: no cover
    raise ImportError("tensorf_jax needs JAX; in this image use tensorf_b200.render (torch front end)") from e

_HERE = pathlib.Path(__file__).resolve().parent
_PKG = _HERE.parent / "tensorf_b200"
_shim = ctypes.CDLL(str(_HERE / "libtensorf_b200_xla.so"))
_core = ctypes.CDLL(str(_PKG / "libtensorf_b200.so"))
_core.tensorf_vm_packed_floats.restype = ctypes.c_int64
_core.tensorf_adam_scratch_bytes.restype = ctypes.c_int64

# host constants of the contracted branch (numpy only), loaded by path so that the torch front end is not imported
_spec = importlib.util.spec_from_file_location("tensorf_b200_schedule", _PKG / "schedule.py")
_schedule = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_schedule)

for _name, _sym in (("tensorf_render_rgb_fwd", "TensorfRenderRgbFwd"), ("tensorf_render_rgb_bwd", "TensorfRenderRgbBwd"),
                    ("tensorf_render_depth", "TensorfRenderDepth"), ("tensorf_vm_interp_fwd", "TensorfVmInterpFwd"),
                    ("tensorf_vm_interp_bwd", "TensorfVmInterpBwd"), ("tensorf_adam_step", "TensorfAdamStep")):
    jax.ffi.register_ffi_target(_name, jax.ffi.pycapsule(getattr(_shim, _sym)), platform="CUDA")

MODE_RGB, MODE_DIST_MEDIAN, MODE_DIST_MEAN = 0, 1, 2


class _Desc(ctypes.Structure):  # struct tensorf_render_desc (include/tensorf_b200.h)
    _fields_ = [(n, ctypes.c_int32) for n in ("R", "N", "K", "G", "cd", "ca", "mode", "contracted", "squash", "units", "feat_freqs",
                                              "view_freqs", "num_cameras", "mlp_impl")] + [("loss_scale", ctypes.c_float), ("flags", ctypes.c_int32)]


def _workspace_bytes(R, N, K, G, cd, ca, mode, contracted, feat_freqs, view_freqs, num_cameras, inference) -> int:
    d = _Desc(R, N, K, G, cd, ca, mode, int(contracted), 27, 128, feat_freqs, view_freqs, num_cameras, 0, 0.0, 1 if inference else 0)
    n = ctypes.c_int64()
    if _core.tensorf_render_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) != 0:
        raise ValueError("tensorf_render_workspace_bytes rejected the configuration")
    return int(n.value)


def _leaves(lp):
    """The 12 leaves in the order the handlers take them; a (1,128) zero stands in for a missing embedding."""
    m = lp.appearance_mlp_params["params"]
    embed = m["Embed_0"]["embedding"] if "Embed_0" in m else jnp.zeros((1, 128), jnp.float32)
    return (m["Dense_0"]["kernel"], m["Dense_1"]["bias"], m["Dense_1"]["kernel"], m["Dense_2"]["bias"], m["Dense_2"]["kernel"],
            m["Dense_3"]["bias"], m["Dense_3"]["kernel"], embed,
            lp.appearance_tensor.stacked_single_vm.vector, lp.appearance_tensor.stacked_single_vm.matrix,
            lp.density_tensor.stacked_single_vm.vector, lp.density_tensor.stacked_single_vm.matrix)


def _attrs(static):
    N, K, contracted, ff, vf, ncam = static
    return dict(density_samples_per_ray=N, appearance_samples_per_ray=K, scene_contraction=int(contracted), feature_n_freqs=ff,
                viewdir_n_freqs=vf, num_cameras=ncam)


def _ws_bytes_for(static, leaves, R, inference):
    N, K, contracted, ff, vf, ncam = static
    den_mat, app_mat = leaves[11], leaves[9]
    return _workspace_bytes(R, N, K, den_mat.shape[2], den_mat.shape[1], app_mat.shape[1], MODE_RGB, contracted, ff, vf, ncam, inference)


def _rgb_call(static, leaves, rays, inference):
    R = rays[1].shape[0]
    nbytes = _ws_bytes_for(static, leaves, R, inference)
    return jax.ffi.ffi_call(
        "tensorf_render_rgb_fwd",
        (jax.ShapeDtypeStruct((R, 3), jnp.float32), jax.ShapeDtypeStruct((nbytes,), jnp.uint8)),
    )(*leaves, *rays, inference=int(inference), **_attrs(static))


@functools.partial(jax.custom_vjp, nondiff_argnums=(0,))
def _render_rgb(static, leaves, rays):
    """rays = (aabb, origins, directions, camera_indices, jitter, gumbel, base_ts, deltas)."""
    rgb, _ = _rgb_call(static, leaves, rays, inference=True)  # not differentiated: no residuals kept
    return rgb


def _render_rgb_fwd(static, leaves, rays):
    rgb, ws = _rgb_call(static, leaves, rays, inference=False)
    return rgb, (leaves, rays, ws)


def _render_rgb_bwd(static, res, d_rgb):
    leaves, rays, ws = res
    outs = jax.ffi.ffi_call(
        "tensorf_render_rgb_bwd",
        tuple(jax.ShapeDtypeStruct(x.shape, jnp.float32) for x in leaves) + (jax.ShapeDtypeStruct(ws.shape, jnp.uint8),),
        input_output_aliases={len(leaves) + len(rays): len(leaves)},  # the workspace is consumed in place
    )(*leaves, *rays, ws, d_rgb.astype(jnp.float32), **_attrs(static))
    grads = tuple(outs[:len(leaves)])
    zeros = tuple(onp.zeros(r.shape, jax.dtypes.float0) if not jnp.issubdtype(r.dtype, jnp.floating) else jnp.zeros_like(r) for r in rays)
    return grads, zeros


_render_rgb.defvjp(_render_rgb_fwd, _render_rgb_bwd)


def render_rays(appearance_mlp, learnable_params, aabb, rays_wrt_world, prng_key, config):
    """Drop-in for tensorf.render.render_rays (render.py:105-279): (ray_count, 3) for RenderMode.RGB (differentiable
    w.r.t. `learnable_params`), (ray_count,) for DIST_MEDIAN / DIST_MEAN."""
    R = rays_wrt_world.origins.shape[0]
    N, K = config.density_samples_per_ray, config.appearance_samples_per_ray
    contracted = bool(learnable_params.scene_contraction)
    k_sample, k_rgb = jax.random.split(prng_key)                            # render.py:120
    jitter = jax.random.uniform(k_sample, (R, N) if contracted else (N,))    # :158-160 / :375-379
    if contracted:
        base, delta = _schedule.contracted_schedule(config.near, config.far, N)   # :127-155 (host, float64 numpy)
    else:
        base = delta = onp.zeros(N, onp.float32)
    mode = getattr(config.mode, "name", str(config.mode))
    if mode.endswith("RGB"):
        gumbel = jax.random.gumbel(k_rgb, (N,))                              # :462-468 (the noise inside random.choice)
        static = (N, K, contracted, appearance_mlp.feature_n_freqs, appearance_mlp.viewdir_n_freqs, appearance_mlp.num_cameras or 0)
        rays = (aabb, rays_wrt_world.origins, rays_wrt_world.directions, rays_wrt_world.camera_indices.astype(jnp.uint32), jitter,
                gumbel, jnp.asarray(base), jnp.asarray(delta))
        leaves = _leaves(learnable_params)
        rgb = _render_rgb(static, leaves, rays)
        return rgb
    code = MODE_DIST_MEDIAN if mode.endswith("MEDIAN") else MODE_DIST_MEAN
    vm = learnable_params.density_tensor.stacked_single_vm
    nbytes = _workspace_bytes(R, N, 1, vm.matrix.shape[2], vm.matrix.shape[1], vm.matrix.shape[1], code, contracted, 0, 0, 0, True)
    depth, _ = jax.ffi.ffi_call(
        "tensorf_render_depth", (jax.ShapeDtypeStruct((R,), jnp.float32), jax.ShapeDtypeStruct((nbytes,), jnp.uint8)),
    )(vm.vector, vm.matrix, aabb, rays_wrt_world.origins, rays_wrt_world.directions, jitter, jnp.asarray(base), jnp.asarray(delta),
      density_samples_per_ray=N, mode=code, scene_contraction=int(contracted))
    return depth


# ---- TensorVM.interpolate (tensor_vm.py:42-89) -----------------------------------------------------------------------
@jax.custom_vjp
def interpolate(vector, matrix, ijk):
    """vector (3,C,G), matrix (3,C,G,G), ijk (3,*) in [-1,1] -> (3C,*)."""
    return _interp_fwd(vector, matrix, ijk)[0]


def _interp_call(vector, matrix, ijk):
    C, G = vector.shape[1], vector.shape[2]
    flat = ijk.reshape(3, -1).astype(jnp.float32)
    n_packed = int(_core.tensorf_vm_packed_floats(C, G))
    out, packed = jax.ffi.ffi_call(
        "tensorf_vm_interp_fwd", (jax.ShapeDtypeStruct((3 * C, flat.shape[1]), jnp.float32), jax.ShapeDtypeStruct((n_packed,), jnp.float32)),
    )(vector, matrix, flat)
    return out.reshape((3 * C,) + ijk.shape[1:]), packed, flat


def _interp_fwd(vector, matrix, ijk):
    out, packed, flat = _interp_call(vector, matrix, ijk)
    return out, (packed, flat, vector.shape, matrix.shape, ijk.shape)


def _interp_bwd(res, d_out):
    packed, flat, vshape, mshape, ishape = res
    d_vec, d_mat, _ = jax.ffi.ffi_call(
        "tensorf_vm_interp_bwd",
        (jax.ShapeDtypeStruct(vshape, jnp.float32), jax.ShapeDtypeStruct(mshape, jnp.float32), jax.ShapeDtypeStruct(packed.shape, jnp.float32)),
    )(packed, flat, d_out.reshape(d_out.shape[0], -1).astype(jnp.float32))
    return d_vec, d_mat, jnp.zeros(ishape, jnp.float32)  # the reference never differentiates w.r.t. the coordinates


interpolate.defvjp(_interp_fwd, _interp_bwd)


# ---- optimiser step (training.py:158-243) ---------------------------------------------------------------------------------
def adam_step(params, grads, mu, nu, neg_lrs, count: int, lr_decay: float = 1.0, b1=0.9, b2=0.99, eps=1e-8, eps_root=0.0):
    """optax.chain(scale_by_adam, masked(scale(-lr)) ...) + apply_updates over lists of leaves, in place (every output is
    aliased to its input; donate the state as training.py:101 does).  `count` = steps taken so far (a Python int: the
    bias corrections are host values).  Returns (params, mu, nu, global_norm(grads))."""
    n = len(params)
    sizes = (ctypes.c_int64 * n)(*[int(p.size) for p in params])
    scratch = int(_core.tensorf_adam_scratch_bytes(sizes, n))
    t = count + 1
    outs = jax.ffi.ffi_call(
        "tensorf_adam_step",
        tuple(jax.ShapeDtypeStruct(x.shape, jnp.float32) for x in (*params, *mu, *nu)) +
        (jax.ShapeDtypeStruct((), jnp.float32), jax.ShapeDtypeStruct((max(scratch, 16),), jnp.uint8)),
        input_output_aliases={**{i: i for i in range(n)}, **{2 * n + i: n + i for i in range(2 * n)}},
    )(*params, *grads, *mu, *nu, neg_lrs=onp.asarray(neg_lrs, onp.float32), b1=onp.float32(b1), b2=onp.float32(b2),
      eps=onp.float32(eps), eps_root=onp.float32(eps_root), bias_correction1=onp.float32(1.0 - b1 ** t),
      bias_correction2=onp.float32(1.0 - b2 ** t), lr_decay=onp.float32(lr_decay))
    return list(outs[:n]), list(outs[n:2 * n]), list(outs[2 * n:3 * n]), outs[3 * n]

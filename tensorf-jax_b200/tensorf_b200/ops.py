"""Torch-tensor front end of the C ABI (include/tensorf_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream, every computation is a
call into libtensorf_b200.so.  All tensors must be CUDA, contiguous, fp32 (int32 / uint32-as-
int32 for indices).  Work is enqueued on `torch.cuda.current_stream()`.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import PARAM_FIELDS, AdamDesc, Params, RenderDesc, RenderInputs, check

MODE_RGB, MODE_DIST_MEDIAN, MODE_DIST_MEAN = 0, 1, 2
MLP_AUTO, MLP_SIMT_FP32, MLP_TCGEN05, MLP_FUSED = 0, 1, 2, 3
FLAG_INFERENCE = 1  # forward only: residuals of the reverse pass are not kept
FLAG_PACKED_FACTORS = 2  # TENSORF_FLAG_PACKED_FACTORS: factor leaves are the kernel-native packed copies


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor], dtype=torch.float32, name: str = "tensor") -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t.data_ptr()


def make_desc(R: int, N: int, K: int, G: int, cd: int, ca: int, mode: int = MODE_RGB, contracted: bool = False,
              feat_freqs: int = 6, view_freqs: int = 6, num_cameras: Optional[int] = None, loss_scale: float = 0.0,
              squash: int = 27, units: int = 128, mlp_impl: int = MLP_AUTO, inference: bool = False,
              packed_factors: bool = False) -> RenderDesc:
    """`packed_factors`: the factor leaves of params / grads are the kernel-native packed copies (`vm_pack`), named
    'density_packed' / 'appearance_packed' (1-D); the pack and unpack passes of every step are skipped."""
    return RenderDesc(R=R, N=N, K=K, G=G, cd=cd, ca=ca, mode=mode, contracted=int(bool(contracted)), squash=squash,
                      units=units, feat_freqs=feat_freqs, view_freqs=view_freqs, num_cameras=int(num_cameras or 0),
                      mlp_impl=mlp_impl, loss_scale=loss_scale,
                      flags=(FLAG_INFERENCE if inference else 0) | (FLAG_PACKED_FACTORS if packed_factors else 0))


def encoded_dim(desc: RenderDesc) -> int:
    return desc.squash + 3 + 2 * desc.feat_freqs * desc.squash + 2 * desc.view_freqs * 3


def param_shapes(desc: RenderDesc) -> Dict[str, Tuple[int, ...]]:
    """Leaf shapes of LearnableParams (render.py:39-46) in the reference's layouts."""
    G, cd, ca, u = desc.G, desc.cd, desc.ca, desc.units
    if desc.flags & FLAG_PACKED_FACTORS:
        s = {"density_packed": (vm_packed_floats(cd, G),), "appearance_packed": (vm_packed_floats(ca, G),)}
    else:
        s = {"density_vector": (3, cd, G), "density_matrix": (3, cd, G, G),
             "appearance_vector": (3, ca, G), "appearance_matrix": (3, ca, G, G)}
    s.update({
        "w0": (3 * ca, desc.squash), "w1": (encoded_dim(desc), u), "b1": (u,), "w2": (u, u), "b2": (u,),
        "w3": (u, 3), "b3": (3,),
    })
    if desc.num_cameras > 0:
        s["embed"] = (desc.num_cameras, u)
    return s


def _params_struct(desc: RenderDesc, params: Dict[str, torch.Tensor], what: str = "params", factors: bool = True) -> Params:
    shapes = param_shapes(desc)
    ps = Params()
    packed = bool(desc.flags & FLAG_PACKED_FACTORS)
    for name in PARAM_FIELDS:
        if packed and factors and name in ("density_vector", "appearance_vector"):
            # TENSORF_FLAG_PACKED_FACTORS: `*_vector` carries the packed buffer, `*_matrix` stays NULL
            key = name.replace("_vector", "_packed")
            if key not in params:
                raise KeyError(f"{what} is missing leaf '{key}'")
            t = params[key]
            if tuple(t.shape) != shapes[key]:
                raise ValueError(f"{what}['{key}'] has shape {tuple(t.shape)}, expected {shapes[key]}")
            setattr(ps, name, _ptr(t, name=f"{what}['{key}']"))
            continue
        if name not in shapes or (not factors and name.startswith(("density_", "appearance_"))):
            setattr(ps, name, None)
            continue
        if name not in params:
            raise KeyError(f"{what} is missing leaf '{name}'")
        t = params[name]
        if tuple(t.shape) != shapes[name]:
            raise ValueError(f"{what}['{name}'] has shape {tuple(t.shape)}, expected {shapes[name]}")
        setattr(ps, name, _ptr(t, name=f"{what}['{name}']"))
    return ps


def pack_params(params: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Reference-layout leaves -> the leaves a `packed_factors` call takes (MLP leaves are shared, not copied)."""
    out = {k: v for k, v in params.items() if not k.startswith(("density_", "appearance_"))}
    out["density_packed"] = vm_pack(params["density_vector"], params["density_matrix"])
    out["appearance_packed"] = vm_pack(params["appearance_vector"], params["appearance_matrix"])
    return out


def unpack_params(packed: Dict[str, torch.Tensor], cd: int, ca: int, G: int) -> Dict[str, torch.Tensor]:
    """Inverse of `pack_params` (checkpoints, grid resampling, comparison with the reference layout)."""
    out = {k: v for k, v in packed.items() if not k.endswith("_packed")}
    out["density_vector"], out["density_matrix"] = vm_unpack(packed["density_packed"], cd, G)
    out["appearance_vector"], out["appearance_matrix"] = vm_unpack(packed["appearance_packed"], ca, G)
    return out


# ---------------------------------------------------------------------------------------------
# TensorVM
# ---------------------------------------------------------------------------------------------
def vm_packed_floats(C_: int, G: int) -> int:
    return int(_lib.load().tensorf_vm_packed_floats(C_, G))


def vm_pack(vector: torch.Tensor, matrix: torch.Tensor) -> torch.Tensor:
    """(3,C,G), (3,C,G,G) channel-first (tensor_vm.py:129-138) -> packed texel-major copy."""
    if vector.dim() != 3 or matrix.dim() != 4 or vector.shape[0] != 3 or matrix.shape[0] != 3:
        raise ValueError(f"expected vector (3,C,G) and matrix (3,C,G,G), got {tuple(vector.shape)}, {tuple(matrix.shape)}")
    _, Cc, G = vector.shape
    if tuple(matrix.shape) != (3, Cc, G, G):
        raise ValueError(f"matrix shape {tuple(matrix.shape)} does not match vector {tuple(vector.shape)}")
    packed = torch.empty(vm_packed_floats(Cc, G), dtype=torch.float32, device=vector.device)
    check(_lib.load().tensorf_vm_pack(_stream(), _ptr(vector, name="vector"), _ptr(matrix, name="matrix"), _ptr(packed), Cc, G))
    return packed


def vm_unpack(packed: torch.Tensor, C_: int, G: int) -> Tuple[torch.Tensor, torch.Tensor]:
    if packed.numel() != vm_packed_floats(C_, G):
        raise ValueError("packed buffer has the wrong size")
    vector = torch.empty((3, C_, G), dtype=torch.float32, device=packed.device)
    matrix = torch.empty((3, C_, G, G), dtype=torch.float32, device=packed.device)
    check(_lib.load().tensorf_vm_unpack(_stream(), _ptr(packed), _ptr(vector), _ptr(matrix), C_, G))
    return vector, matrix


def vm_interp_fwd(packed: torch.Tensor, ijk: torch.Tensor, C_: int, G: int, feature_major: bool = False) -> torch.Tensor:
    """ijk (3,B) -> (3C,B) or (B,3C)."""
    if ijk.dim() != 2 or ijk.shape[0] != 3:
        raise ValueError(f"ijk must be (3,B), got {tuple(ijk.shape)}")
    B = ijk.shape[1]
    out = torch.empty((B, 3 * C_) if feature_major else (3 * C_, B), dtype=torch.float32, device=ijk.device)
    check(_lib.load().tensorf_vm_interp_fwd(_stream(), _ptr(packed), _ptr(ijk, name="ijk"), _ptr(out), C_, G, B, int(feature_major)))
    return out


def vm_interp_bwd(packed: torch.Tensor, ijk: torch.Tensor, d_out: torch.Tensor, C_: int, G: int,
                  feature_major: bool = False, d_packed: Optional[torch.Tensor] = None) -> torch.Tensor:
    B = ijk.shape[1]
    expect = (B, 3 * C_) if feature_major else (3 * C_, B)
    if tuple(d_out.shape) != expect:
        raise ValueError(f"d_out must be {expect}, got {tuple(d_out.shape)}")
    if d_packed is None:
        d_packed = torch.zeros_like(packed)
    check(_lib.load().tensorf_vm_interp_bwd(_stream(), _ptr(packed), _ptr(ijk, name="ijk"), _ptr(d_out, name="d_out"),
                                            _ptr(d_packed), C_, G, B, int(feature_major)))
    return d_packed


def topk_select(g: torch.Tensor, K: int) -> torch.Tensor:
    """render.py:461-469 selection stage: (R,N) -> (R,K) int32, ascending index order."""
    if g.dim() != 2:
        raise ValueError("g must be (R,N)")
    R, N = g.shape
    idx = torch.empty((R, K), dtype=torch.int32, device=g.device)
    check(_lib.load().tensorf_topk_select(_stream(), _ptr(g, name="g"), R, N, K, _ptr(idx, torch.int32)))
    return idx


def segment_probabilities(sigmas: torch.Tensor, step_sizes: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """render.py:300-347: (R,N),(R,N) -> p_exits, p_terminates."""
    if sigmas.dim() != 2 or sigmas.shape != step_sizes.shape:
        raise ValueError("sigmas and step_sizes must both be (R,N)")
    R, N = sigmas.shape
    pe, pt = torch.empty_like(sigmas), torch.empty_like(sigmas)
    check(_lib.load().tensorf_segment_probabilities(_stream(), _ptr(sigmas, name="sigmas"), _ptr(step_sizes, name="step_sizes"),
                                                    R, N, _ptr(pe), _ptr(pt)))
    return pe, pt


# ---------------------------------------------------------------------------------------------
# FeatureMlp
# ---------------------------------------------------------------------------------------------
class MlpCall:
    """One FeatureMlp.apply (networks.py:46-121) with its saved activations."""

    def __init__(self, desc: RenderDesc, M: int, device):
        self.desc, self.M = desc, M
        nbytes = int(_lib.load().tensorf_mlp_workspace_bytes(C.byref(desc), M))
        if nbytes < 0:
            raise ValueError("bad MLP description")
        self.workspace = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=device)
        self.rgb = None

    def forward(self, params, features, viewdirs, camera_indices, rows_per_ray: int = 1) -> torch.Tensor:
        d = self.desc
        if tuple(features.shape) != (self.M, 3 * d.ca):
            raise ValueError(f"features must be {(self.M, 3 * d.ca)}, got {tuple(features.shape)}")
        if tuple(viewdirs.shape) != (self.M // rows_per_ray, 3):
            raise ValueError(f"viewdirs must be {(self.M // rows_per_ray, 3)}, got {tuple(viewdirs.shape)}")
        ps = _params_struct(d, params, factors=False)
        rgb = torch.empty((self.M, 3), dtype=torch.float32, device=features.device)
        cams = _ptr(camera_indices, torch.int32, "camera_indices") if d.num_cameras > 0 else None
        check(_lib.load().tensorf_mlp_fwd(_stream(), C.byref(d), C.byref(ps), _ptr(features, name="features"),
                                          _ptr(viewdirs, name="viewdirs"), cams, self.M, rows_per_ray,
                                          _ptr(self.workspace), _ptr(rgb)))
        self._saved = (params, features, viewdirs, camera_indices, rows_per_ray)
        self.rgb = rgb
        return rgb

    def backward(self, d_rgb: torch.Tensor):
        d = self.desc
        params, features, viewdirs, camera_indices, rows_per_ray = self._saved
        ps = _params_struct(d, params, factors=False)
        grads = {k: torch.empty_like(v) for k, v in params.items() if not k.startswith(("density_", "appearance_"))}
        gs = _params_struct(d, grads, "grads", factors=False)
        d_feat = torch.empty_like(features)
        cams = _ptr(camera_indices, torch.int32, "camera_indices") if d.num_cameras > 0 else None
        check(_lib.load().tensorf_mlp_bwd(_stream(), C.byref(d), C.byref(ps), _ptr(features), _ptr(viewdirs), cams, self.M,
                                          rows_per_ray, _ptr(self.workspace), _ptr(self.rgb), _ptr(d_rgb, name="d_rgb"),
                                          _ptr(d_feat), C.byref(gs)))
        return d_feat, grads


# ---------------------------------------------------------------------------------------------
# render_rays
# ---------------------------------------------------------------------------------------------
class RenderCall:
    """One invocation of render.py:105-279 (forward, optional fused loss, reverse).  Owns the
    workspace the C ABI asks the caller to provide; reusable across calls of the same shape."""

    def __init__(self, desc: RenderDesc, device):
        self.desc = desc
        nbytes = C.c_int64(0)
        check(_lib.load().tensorf_render_workspace_bytes(C.byref(desc), C.byref(nbytes)))
        self.workspace_bytes = int(nbytes.value)
        self.workspace = torch.empty(max(self.workspace_bytes // 4, 1), dtype=torch.float32, device=device)
        self.device = device
        self._keep = None

    def _inputs(self, inputs: Dict[str, Optional[torch.Tensor]]) -> RenderInputs:
        d = self.desc
        R, N = d.R, d.N
        exp = {
            "origins": ((R, 3), torch.float32), "directions": ((R, 3), torch.float32), "aabb": ((2, 3), torch.float32),
            "jitter": ((R, N) if d.contracted else (N,), torch.float32),
        }
        if d.mode == MODE_RGB:
            exp["gumbel"] = ((N,), torch.float32)
            if d.num_cameras > 0:
                exp["camera_indices"] = ((R,), torch.int32)
        if d.contracted:
            exp["base_ts"] = ((N,), torch.float32)
            exp["deltas"] = ((N,), torch.float32)
        ri = RenderInputs()
        for name, (shape, dt) in exp.items():
            t = inputs.get(name)
            if t is None:
                raise KeyError(f"inputs is missing '{name}'")
            if tuple(t.shape) != shape:
                raise ValueError(f"inputs['{name}'] has shape {tuple(t.shape)}, expected {shape}")
            setattr(ri, name, _ptr(t, dt, f"inputs['{name}']"))
        colors = inputs.get("colors")
        if colors is not None:
            if tuple(colors.shape) != (R, 3):
                raise ValueError(f"colors must be {(R, 3)}")
            ri.colors = _ptr(colors, name="colors")
        return ri

    def forward(self, params: Dict[str, torch.Tensor], inputs: Dict[str, Optional[torch.Tensor]],
                loss_out: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
        """Returns (rgb (R,3), loss or None). `inputs['colors']` enables the fused MSE; `loss_out` (one fp32
        element, e.g. the loss slot of `dist.FlatGrads`) receives it in place of a fresh scalar; `out` (R,3)
        receives the colours in place of a fresh tensor (a slice of a frame buffer)."""
        d = self.desc
        ps = _params_struct(d, params)
        ri = self._inputs(inputs)
        if out is not None and (tuple(out.shape) != (d.R, 3) or not out.is_contiguous()):
            raise ValueError(f"out must be a contiguous {(d.R, 3)} tensor")
        rgb = out if out is not None else torch.empty((d.R, 3), dtype=torch.float32, device=self.device)
        loss = None
        if inputs.get("colors") is not None:
            if loss_out is not None and loss_out.numel() != 1:
                raise ValueError("loss_out must hold exactly one element")
            loss = loss_out if loss_out is not None else torch.empty((), dtype=torch.float32, device=self.device)
        check(_lib.load().tensorf_render_rgb_fwd(_stream(), C.byref(d), C.byref(ps), C.byref(ri), _ptr(self.workspace),
                                                 _ptr(rgb), _ptr(loss)))
        self._keep = (params, inputs)
        return rgb, loss

    def backward(self, d_rgb: Optional[torch.Tensor] = None, grads: Optional[Dict[str, torch.Tensor]] = None, phase: int = 0):
        """Gradients w.r.t. every leaf of LearnableParams. d_rgb=None uses the fused loss cotangent.
        phase 1 / 2 = the appearance / density half of the pass (`tensorf_render_rgb_bwd_phase`): after phase 1
        every leaf but the density factors is final (sharded training starts their exchange there)."""
        d = self.desc
        params, inputs = self._keep
        ps = _params_struct(d, params)
        ri = self._inputs(inputs)
        if grads is None:
            grads = {k: torch.empty_like(v) for k, v in params.items() if k in param_shapes(d)}
        gs = _params_struct(d, grads, "grads")
        if d_rgb is not None and tuple(d_rgb.shape) != (d.R, 3):
            raise ValueError(f"d_rgb must be {(d.R, 3)}")
        check(_lib.load().tensorf_render_rgb_bwd_phase(_stream(), C.byref(d), C.byref(ps), C.byref(ri), _ptr(self.workspace),
                                                       _ptr(d_rgb, name="d_rgb"), C.byref(gs), int(phase)))
        return grads

    def depth(self, params: Dict[str, torch.Tensor], inputs: Dict[str, Optional[torch.Tensor]],
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
        d = self.desc
        ps = Params()
        if d.flags & FLAG_PACKED_FACTORS:
            t = params["density_packed"]
            if tuple(t.shape) != param_shapes(d)["density_packed"]:
                raise ValueError(f"params['density_packed'] has shape {tuple(t.shape)}")
            ps.density_vector = _ptr(t, name="density_packed")
        else:
            for name in ("density_vector", "density_matrix"):
                t = params[name]
                if tuple(t.shape) != param_shapes(d)[name]:
                    raise ValueError(f"params['{name}'] has shape {tuple(t.shape)}")
                setattr(ps, name, _ptr(t, name=name))
        ri = self._inputs(inputs)
        if out is not None and (tuple(out.shape) != (d.R,) or not out.is_contiguous()):
            raise ValueError(f"out must be a contiguous {(d.R,)} tensor")
        out = out if out is not None else torch.empty((d.R,), dtype=torch.float32, device=self.device)
        check(_lib.load().tensorf_render_depth(_stream(), C.byref(d), C.byref(ps), C.byref(ri), _ptr(self.workspace), _ptr(out)))
        return out

    def view(self, name: str) -> torch.Tensor:
        """A copy of one residual stored in the workspace (tests / debugging)."""
        ptr, cnt = C.c_void_p(), C.c_int64()
        check(_lib.load().tensorf_render_workspace_view(C.byref(self.desc), _ptr(self.workspace), name.encode(), C.byref(ptr),
                                                        C.byref(cnt)))
        off = (ptr.value - self.workspace.data_ptr()) // 4
        t = self.workspace[off:off + cnt.value]
        return t.view(torch.int32).clone() if name == "idx" else t.clone()


# ---------------------------------------------------------------------------------------------
# measurement hooks
# ---------------------------------------------------------------------------------------------
class AdamCall:
    """`tensorf_adam_step` over a fixed list of leaves (training.py:158-243): the pointer tables and the
    scratch buffer are built once, every step is one kernel launch.  `neg_lrs[i]` = -(group learning rate)."""

    def __init__(self, params, mu, nu, neg_lrs, b1: float = 0.9, b2: float = 0.99, eps: float = 1e-8, eps_root: float = 0.0):
        self.lib = _lib.load()
        n = len(params)
        if not (n == len(mu) == len(nu) == len(neg_lrs)) or n > _lib.ADAM_MAX_LEAVES:
            raise ValueError(f"adam: {n} leaves (max {_lib.ADAM_MAX_LEAVES}) with mismatched state lists")
        self.params, self.mu, self.nu = list(params), list(mu), list(nu)
        for i, (p, m, v) in enumerate(zip(self.params, self.mu, self.nu)):
            if m.shape != p.shape or v.shape != p.shape:
                raise ValueError(f"adam: leaf {i} state shape mismatch")
        self.n = n
        self.b1, self.b2, self.eps, self.eps_root = b1, b2, eps, eps_root
        self.sizes = (C.c_int64 * n)(*[p.numel() for p in self.params])
        self.neg_lrs = (C.c_float * n)(*[float(x) for x in neg_lrs])
        self._p = (C.c_void_p * n)(*[_ptr(t, name="param") for t in self.params])
        self._m = (C.c_void_p * n)(*[_ptr(t, name="mu") for t in self.mu])
        self._v = (C.c_void_p * n)(*[_ptr(t, name="nu") for t in self.nu])
        nbytes = self.lib.tensorf_adam_scratch_bytes(self.sizes, n)
        dev = self.params[0].device
        self.scratch = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
        self.grad_norm = torch.zeros((), dtype=torch.float32, device=dev)
        self._g_key, self._g_ptrs = None, None

    def step(self, grads, count: int, lr_decay: float = 1.0) -> torch.Tensor:
        """One optimiser step; `count` = steps taken so far (optax's count before increment).  Returns the
        device scalar global_norm(grads)."""
        import numpy as np
        if len(grads) != self.n:
            raise ValueError("adam: gradient list does not match the leaves")
        for g, p in zip(grads, self.params):
            if g.shape != p.shape:
                raise ValueError("adam: gradient shape mismatch")
        key = tuple(t.data_ptr() for t in grads)
        if self._g_key != key:  # pointer table rebuilt only when the gradient buffers change
            self._g_ptrs = (C.c_void_p * self.n)(*[_ptr(t, name="grad") for t in grads])
            self._g_key = key
        g_ptrs = self._g_ptrs
        t = np.float32(count + 1)
        # optax bias_correction: 1 - decay**count in fp32
        bc1 = np.float32(1) - np.power(np.float32(self.b1), t)
        bc2 = np.float32(1) - np.power(np.float32(self.b2), t)
        d = AdamDesc(n_leaves=self.n, reserved=0, b1=self.b1, b2=self.b2, eps=self.eps, eps_root=self.eps_root,
                     bias_correction1=float(bc1), bias_correction2=float(bc2), lr_decay=float(lr_decay), reserved2=0.0)
        check(self.lib.tensorf_adam_step(_stream(), C.byref(d), self.sizes, self._p, g_ptrs, self._m, self._v, self.neg_lrs,
                                         self.grad_norm.data_ptr(), self.scratch.data_ptr(), self.scratch.numel()))
        return self.grad_norm


def vm_resize(vector: torch.Tensor, matrix: torch.Tensor, grid_dim: int, out=None, scratch: Optional[torch.Tensor] = None
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """tensor_vm.py:183-223: (3,C,G) / (3,C,G,G) -> (3,C,grid_dim) / (3,C,grid_dim,grid_dim).  `out` = (vector, matrix)
    tensors to fill and `scratch` (uint8, `vm_resize_scratch_bytes`) let a caller that resizes several leaves of the same
    shape (parameters and both Adam moments, training.py:245-276) allocate once."""
    lib = _lib.load()
    three, C_, G = vector.shape
    if three != 3 or tuple(matrix.shape) != (3, C_, G, G):
        raise ValueError(f"vm_resize: vector {tuple(vector.shape)} / matrix {tuple(matrix.shape)} are not a TensorVM")
    if out is not None:
        vo, mo = out
        if tuple(vo.shape) != (3, C_, grid_dim) or tuple(mo.shape) != (3, C_, grid_dim, grid_dim):
            raise ValueError("vm_resize: `out` tensors have the wrong shape")
    else:
        vo = torch.empty((3, C_, grid_dim), dtype=torch.float32, device=vector.device)
        mo = torch.empty((3, C_, grid_dim, grid_dim), dtype=torch.float32, device=vector.device)
    nbytes = vm_resize_scratch_bytes(C_, G, grid_dim)
    if scratch is None or scratch.numel() < nbytes:
        scratch = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=vector.device)
    check(lib.tensorf_vm_resize(_stream(), _ptr(vector, name="vector"), _ptr(matrix, name="matrix"), C_, G, grid_dim,
                                vo.data_ptr(), mo.data_ptr(), scratch.data_ptr(), scratch.numel()))
    return vo, mo


def vm_resize_scratch_bytes(C_: int, G: int, grid_dim: int) -> int:
    nbytes = int(_lib.load().tensorf_vm_resize_scratch_bytes(C_, G, grid_dim))
    if nbytes < 0:
        raise ValueError(f"vm_resize: unsupported shape C={C_} G={G} -> {grid_dim}")
    return nbytes


def threefry2x32(k0: int, k1: int, x0: int, x1: int) -> Tuple[int, int]:
    """The library's host threefry block (for known-answer checks)."""
    out = (C.c_uint32 * 2)()
    _lib.load().tensorf_threefry2x32(k0, k1, x0, x1, out)
    return int(out[0]), int(out[1])


def prng_uniform(k0: int, k1: int, shape, device, minval: float = 0.0, maxval: float = 1.0, first: int = 0) -> torch.Tensor:
    """jax.random.uniform(key, shape, float32, minval, maxval) drawn on the device (`tensorf_prng_uniform`); with
    `first` > 0 the result is elements [first, first + prod(shape)) of a larger flat draw (a rank's row slice of a
    sharded (R,N) jitter, `tensorf_prng_uniform_slice`)."""
    out = torch.empty(tuple(shape), dtype=torch.float32, device=device)
    check(_lib.load().tensorf_prng_uniform_slice(_stream(), k0, k1, int(first), out.numel(), minval, maxval, out.data_ptr()))
    return out


def prng_gumbel(k0: int, k1: int, shape, device) -> torch.Tensor:
    """jax.random.gumbel(key, shape) drawn on the device (`tensorf_prng_gumbel`)."""
    out = torch.empty(tuple(shape), dtype=torch.float32, device=device)
    check(_lib.load().tensorf_prng_gumbel(_stream(), k0, k1, out.numel(), out.data_ptr()))
    return out


def pixel_rays(M, origin, width: int, rows: Tuple[int, int], camera_index: int, device, out=None):
    """cameras.py:124-143 for image rows [rows[0], rows[1]): M = R_world_camera @ K^-1 (3x3), origin (3,), host
    values. Returns (origins (n,3), directions (n,3), camera_indices (n,) int32 with uint32 bits); `out` = such a
    triple of contiguous tensors to fill (slices of a frame's ray table) instead of fresh ones."""
    r0, r1 = rows
    n = (r1 - r0) * width
    if out is not None:
        o, d, c = out
        if tuple(o.shape) != (n, 3) or tuple(d.shape) != (n, 3) or tuple(c.shape) != (n,):
            raise ValueError("pixel_rays: `out` tensors do not match the row range")
    else:
        o = torch.empty((n, 3), dtype=torch.float32, device=device)
        d = torch.empty((n, 3), dtype=torch.float32, device=device)
        c = torch.empty((n,), dtype=torch.int32, device=device)
    Mh = (C.c_float * 9)(*[float(x) for x in M])
    oh = (C.c_float * 3)(*[float(x) for x in origin])
    check(_lib.load().tensorf_pixel_rays(_stream(), Mh, oh, width, r0, r1, camera_index, o.data_ptr(), d.data_ptr(), c.data_ptr()))
    return o, d, c


def pixel_rays_striped(M, origin, width: int, height: int, stripe: int, rank: int, world: int, camera_index: int, out) -> int:
    """cameras.py:124-143 for the rows `rank` owns when a frame is dealt to `world` ranks in interleaved `stripe`-row stripes
    (`dist.stripe_rows`), written contiguously into `out` = (origins (n,3), directions (n,3), camera_indices (n,)) by ONE
    launch (`tensorf_pixel_rays_striped`).  Returns n."""
    o, d, c = out
    Mh = (C.c_float * 9)(*[float(x) for x in M])
    oh = (C.c_float * 3)(*[float(x) for x in origin])
    n = C.c_int64(0)
    lib = _lib.load()
    check(lib.tensorf_pixel_rays_striped(_stream(), Mh, oh, width, height, stripe, rank, world, camera_index, None, None, None, C.byref(n)))
    if tuple(o.shape) != (n.value, 3) or tuple(d.shape) != (n.value, 3) or tuple(c.shape) != (n.value,):
        raise ValueError(f"pixel_rays_striped: outputs must hold {n.value} rays")
    check(lib.tensorf_pixel_rays_striped(_stream(), Mh, oh, width, height, stripe, rank, world, camera_index, o.data_ptr(), d.data_ptr(),
                                         c.data_ptr(), None))
    return int(n.value)


def launch_count() -> int:
    """Kernels this thread has enqueued through the library so far."""
    return int(_lib.load().tensorf_launch_count())


def profile_enable(enable: bool = True) -> None:
    check(_lib.load().tensorf_profile_enable(int(enable)))


def profile_read(max_entries: int = 64) -> Dict[str, Tuple[float, int]]:
    """{stage: (total_ms, calls)} measured with CUDA events the library recorded on the launch
    stream since profile_enable()/the last read. Blocks until those events completed."""
    names = C.create_string_buffer(32 * max_entries)
    ms = (C.c_float * max_entries)()
    calls = (C.c_int * max_entries)()
    n = C.c_int(0)
    check(_lib.load().tensorf_profile_read(max_entries, names, ms, calls, C.byref(n)))
    out = {}
    for i in range(n.value):
        out[names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode()] = (float(ms[i]), int(calls[i]))
    return out

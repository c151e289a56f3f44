"""Mirror of tensorf/tensor_vm.py:14-199 over the CUDA library.

`TensorVM` keeps the reference's parameter pytree layout (`stacked_single_vm.vector (3,C,G)`,
`.matrix (3,C,G,G)`, channel-first, tensor_vm.py:129-138); `interpolate` runs the fused
sm_100a gather on a texel-major packed copy and is differentiable w.r.t. the factors through
torch.autograd (the torch analogue of the jax.custom_vjp the JAX binding registers).
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional

import torch

from . import ops


@dataclasses.dataclass
class TensorVMSingle:
    """Vector-matrix pair(s). In a `TensorVM` the three pairs are stacked along axis 0."""

    vector: torch.Tensor
    matrix: torch.Tensor

    def grid_dim(self) -> int:  # tensor_vm.py:169-174
        r0, r1 = self.matrix.shape[-2:]
        r2 = self.vector.shape[-1]
        assert r0 == r1 == r2
        return int(r0)

    def channel_dim(self) -> int:  # tensor_vm.py:176-181
        c0, c1 = self.matrix.shape[-3], self.vector.shape[-2]
        assert c0 == c1
        return int(c0)


class _Interpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vector, matrix, ijk):
        C, G = vector.shape[1], vector.shape[2]
        packed = ops.vm_pack(vector.contiguous(), matrix.contiguous())
        flat = ijk.reshape(3, -1).contiguous()
        out = ops.vm_interp_fwd(packed, flat, C, G, feature_major=False)
        ctx.save_for_backward(packed, flat)
        ctx.dims = (C, G)
        return out.reshape((3 * C,) + tuple(ijk.shape[1:]))

    @staticmethod
    def backward(ctx, d_out):
        packed, flat = ctx.saved_tensors
        C, G = ctx.dims
        d_packed = ops.vm_interp_bwd(packed, flat, d_out.reshape(3 * C, -1).contiguous(), C, G, feature_major=False)
        dv, dm = ops.vm_unpack(d_packed, C, G)
        return dv, dm, None  # sample positions carry no gradient on this path


@dataclasses.dataclass
class TensorVM:
    """Three vector-matrix pairs (tensor_vm.py:14-19)."""

    stacked_single_vm: TensorVMSingle

    @staticmethod
    def initialize(grid_dim: int, per_axis_channel_dim: int, init: Optional[Callable] = None,
                   prng_key: Optional[torch.Generator] = None, dtype=torch.float32, device="cuda") -> "TensorVM":
        """tensor_vm.py:21-40 / :116-138. `init(generator, shape, dtype, device)`; default N(0, 0.1^2)
        as training.py:44 uses."""
        if init is None:
            def init(gen, shape, dt, dev):
                return torch.randn(shape, generator=gen, dtype=dt, device=dev) * 0.1
        return TensorVM(TensorVMSingle(
            vector=init(prng_key, (3, per_axis_channel_dim, grid_dim), dtype, device),
            matrix=init(prng_key, (3, per_axis_channel_dim, grid_dim, grid_dim), dtype, device)))

    def interpolate(self, ijk: torch.Tensor) -> torch.Tensor:
        """tensor_vm.py:42-89: ijk (3, *) in [-1, 1] -> (3*channel, *)."""
        if ijk.shape[0] != 3:
            raise ValueError(f"ijk must have shape (3, *), got {tuple(ijk.shape)}")
        if not ijk.dtype.is_floating_point:
            raise TypeError("ijk must be floating point")
        return _Interpolate.apply(self.stacked_single_vm.vector, self.stacked_single_vm.matrix, ijk.to(torch.float32))

    def grid_dim(self) -> int:
        return self.stacked_single_vm.grid_dim()

    def channel_dim(self) -> int:
        return self.stacked_single_vm.channel_dim() * 3

    def resize(self, grid_dim: int, scratch: Optional[torch.Tensor] = None) -> "TensorVM":
        """tensor_vm.py:91-100 / :183-223: align-corners linear resampling of all three pairs
        (`tensorf_vm_resize`).  `scratch`: a uint8 buffer to reuse across several resizes (`ops.vm_resize_scratch_bytes`)."""
        sv = self.stacked_single_vm
        with torch.no_grad():
            v, m = ops.vm_resize(sv.vector.detach().contiguous(), sv.matrix.detach().contiguous(), int(grid_dim), scratch=scratch)
        v.requires_grad_(sv.vector.requires_grad)
        m.requires_grad_(sv.matrix.requires_grad)
        return TensorVM(stacked_single_vm=TensorVMSingle(vector=v, matrix=m))

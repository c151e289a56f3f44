"""Mirror of tensorf/networks.py:38-121 (FeatureMlp) over the CUDA library."""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional

import torch

from . import ops

# flax leaf path -> flat leaf name used at the C ABI (include/tensorf_b200.h, struct tensorf_params)
FLAX_TO_FLAT = {
    ("Dense_0", "kernel"): "w0", ("Dense_1", "kernel"): "w1", ("Dense_1", "bias"): "b1",
    ("Dense_2", "kernel"): "w2", ("Dense_2", "bias"): "b2", ("Dense_3", "kernel"): "w3", ("Dense_3", "bias"): "b3",
    ("Embed_0", "embedding"): "embed",
}


def flatten_mlp_params(variables: Dict) -> Dict[str, torch.Tensor]:
    """{'params': {'Dense_0': {'kernel': ...}, ...}} -> {'w0': ..., ...}."""
    p = variables["params"] if "params" in variables else variables
    return {flat: p[mod][leaf] for (mod, leaf), flat in FLAX_TO_FLAT.items() if mod in p}


def unflatten_mlp_params(flat: Dict[str, torch.Tensor]) -> Dict:
    out: Dict = {}
    for (mod, leaf), name in FLAX_TO_FLAT.items():
        if name in flat:
            out.setdefault(mod, {})[leaf] = flat[name]
    return {"params": out}


class _MlpApply(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mlp, names, features, viewdirs, camera_indices, *leaves):
        flat = dict(zip(names, leaves))
        M = features.shape[0]
        desc = ops.make_desc(R=M, N=1, K=1, G=2, cd=1, ca=features.shape[1] // 3, feat_freqs=mlp.feature_n_freqs,
                             view_freqs=mlp.viewdir_n_freqs, num_cameras=mlp.num_cameras, squash=mlp.feature_squash_dim,
                             units=mlp.units)
        call = ops.MlpCall(desc, M, features.device)
        rgb = call.forward({k: v.contiguous() for k, v in flat.items()}, features.contiguous(), viewdirs.contiguous(),
                           camera_indices, 1)
        ctx.call, ctx.names = call, names
        return rgb

    @staticmethod
    def backward(ctx, d_rgb):
        d_feat, grads = ctx.call.backward(d_rgb.contiguous())
        return (None, None, d_feat, None, None) + tuple(grads[n] for n in ctx.names)


@dataclasses.dataclass(frozen=True)
class FeatureMlp:
    """networks.py:38-43."""

    feature_squash_dim: int = 27
    units: int = 128
    feature_n_freqs: int = 6
    viewdir_n_freqs: int = 6
    num_cameras: Optional[int] = None

    def encoded_dim(self) -> int:  # networks.py:77-82
        return (self.feature_squash_dim + 3 + 2 * self.feature_n_freqs * self.feature_squash_dim
                + 2 * self.viewdir_n_freqs * 3)

    def init(self, prng_key: Optional[torch.Generator], features: torch.Tensor, viewdirs=None, camera_indices=None) -> Dict:
        """flax `Module.init`: kaiming/lecun normal kernels (networks.py:9-10), zero biases, N(0,1/units)
        embeddings. Only shapes of `features` matter."""
        dev, fin = features.device, features.shape[-1]
        enc, u = self.encoded_dim(), self.units

        def n(shape, var):
            # flax kaiming_normal / lecun_normal = variance_scaling(..., "truncated_normal"): N(0, 1) truncated to
            # [-2, 2], scaled by sqrt(var) / 0.87962566 (the truncated distribution's standard deviation)
            x = torch.empty(shape, device=dev)
            torch.nn.init.trunc_normal_(x, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=prng_key)
            return x * (math.sqrt(var) / 0.87962566103423978)

        p = {"Dense_0": {"kernel": n((fin, self.feature_squash_dim), 1.0 / fin)},
             "Dense_1": {"kernel": n((enc, u), 2.0 / enc), "bias": torch.zeros(u, device=dev)},
             "Dense_2": {"kernel": n((u, u), 2.0 / u), "bias": torch.zeros(u, device=dev)},
             "Dense_3": {"kernel": n((u, 3), 1.0 / u), "bias": torch.zeros(3, device=dev)}}
        if self.num_cameras is not None:
            # flax nn.Embed: variance_scaling(1.0, "fan_in", "normal", out_axis=0) = N(0, 1/units), not truncated
            p["Embed_0"] = {"embedding": torch.randn((self.num_cameras, u), generator=prng_key, device=dev) * math.sqrt(1.0 / u)}
        return {"params": p}

    def apply(self, variables: Dict, features: torch.Tensor, viewdirs: torch.Tensor, camera_indices: torch.Tensor) -> torch.Tensor:
        """networks.py:46-121: features (*, 3ca), viewdirs (*, 3), camera_indices (*,) -> rgb (*, 3)."""
        batch = tuple(features.shape[:-1])
        if tuple(viewdirs.shape) != batch + (3,):
            raise ValueError(f"viewdirs must have shape {batch + (3,)}, got {tuple(viewdirs.shape)}")
        flat = flatten_mlp_params(variables)
        names = tuple(flat)
        cams = camera_indices.reshape(-1).to(torch.int32) if self.num_cameras is not None else None
        rgb = _MlpApply.apply(self, names, features.reshape(-1, features.shape[-1]), viewdirs.reshape(-1, 3), cams,
                              *[flat[n] for n in names])
        return rgb.reshape(batch + (3,))

    __call__ = apply

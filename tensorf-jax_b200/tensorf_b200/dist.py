"""Multi-GPU plumbing for the hot path. The reference is single-device
(README.md:128 lists multi-GPU as a TODO); the path shards naturally:

  * training: the global ray batch is split contiguously, R/P rays per rank, parameters are
    replicated, the loss is the mean over the GLOBAL batch (training.py:140), every rank uses the
    same PRNG key (shared jitter / Gumbel vectors; contracted scenes take the matching row
    slice of the global (R,N) jitter), and ONE sum-allreduce of the gradient pytree follows the
    reverse pass. All gradient leaves live in one flat buffer so that is a single NCCL call.
  * rendering: frames are split into row bands (image tiles), no collective.

torch.distributed is plumbing (NCCL on GPUs, gloo in the CPU tests); nothing here computes.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np
import torch


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `total` units for `rank`; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_rays(arrays: Dict[str, np.ndarray], rank: int, world: int, per_ray: Sequence[str]) -> Dict[str, np.ndarray]:
    """Slice the per-ray arrays (origins, directions, camera_indices, colors, contracted
    jitter) of a global batch; shared arrays (aabb, bounded jitter, gumbel) pass through."""
    R = arrays[per_ray[0]].shape[0]
    a, b = shard_range(R, rank, world)
    return {k: (v[a:b] if k in per_ray else v) for k, v in arrays.items()}


def tile_rows(height: int, rank: int, world: int) -> Tuple[int, int]:
    """Row band of a frame rendered by `rank` (render_360-style frames, no collective)."""
    return shard_range(height, rank, world)


class FlatGrads:
    """All gradient leaves as views into one contiguous fp32 buffer (one allreduce per step)."""

    def __init__(self, shapes: Dict[str, Tuple[int, ...]], device):
        self.shapes = dict(shapes)
        self.total = int(sum(int(np.prod(s)) for s in shapes.values()))
        self.flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.leaves: Dict[str, torch.Tensor] = {}
        off = 0
        for k, s in shapes.items():
            n = int(np.prod(s))
            self.leaves[k] = self.flat[off:off + n].view(s)
            off += n

    def allreduce(self, group=None) -> None:
        """Sum over ranks (training.py:140 is a mean over the whole batch: every rank already
        scaled its cotangent by 1/(3*R_global), so the reduction is a plain sum)."""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)


def global_loss_scale(local_rays: int, world: int) -> float:
    """1/(3*R_global) for equal shards."""
    return 1.0 / (3.0 * local_rays * world)

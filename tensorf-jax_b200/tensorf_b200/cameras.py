"""Mirror of tensorf/cameras.py:10-20 (the `Rays3D` input type of the hot path)."""
from __future__ import annotations

import dataclasses
from typing import Tuple

import torch


@dataclasses.dataclass
class Rays3D:
    """Rays in 3D space. `origins`, `directions` (*, 3) fp32; `camera_indices` (*,) uint32
    (stored as int32 with the same bits), used for per-camera appearance embeddings."""

    origins: torch.Tensor
    directions: torch.Tensor
    camera_indices: torch.Tensor

    def get_batch_axes(self) -> Tuple[int, ...]:
        return tuple(self.origins.shape[:-1])

    def reshape(self, *batch) -> "Rays3D":
        return Rays3D(self.origins.reshape(*batch, 3), self.directions.reshape(*batch, 3), self.camera_indices.reshape(*batch))

    def slice(self, a: int, b: int) -> "Rays3D":
        return Rays3D(self.origins[a:b], self.directions[a:b], self.camera_indices[a:b])

    def to(self, device) -> "Rays3D":
        return Rays3D(self.origins.to(device), self.directions.to(device), self.camera_indices.to(device))

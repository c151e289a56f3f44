"""Mirror of tensorf/cameras.py: `Rays3D` (:10-20, the input type of the hot path) and `Camera`
(:23-143), whose per-pixel ray generation runs on the device (`tensorf_pixel_rays`) instead of the
reference's CPU-pinned jit (:124)."""
from __future__ import annotations

import dataclasses
import math
from typing import Optional, Tuple

import numpy as np
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


@dataclasses.dataclass
class Camera:
    """cameras.py:23-143.  `K` (3,3) intrinsics; `T_camera_world` a 4x4 homogeneous matrix (the reference
    holds a jaxlie.SE3); host numpy values."""

    K: np.ndarray
    T_camera_world: np.ndarray
    image_width: int
    image_height: int

    @staticmethod
    def from_fov(T_camera_world, image_width: int, image_height: int, fov_x_radians: Optional[float] = None,
                 fov_y_radians: Optional[float] = None) -> "Camera":
        """cameras.py:35-77."""
        cx, cy = image_width / 2.0 - 0.5, image_height / 2.0 - 0.5
        fx = (image_width / 2.0) / math.tan(fov_x_radians / 2.0) if fov_x_radians is not None else None
        fy = (image_height / 2.0) / math.tan(fov_y_radians / 2.0) if fov_y_radians is not None else None
        assert fx is not None or fy is not None
        fx = fy if fx is None else fx
        fy = fx if fy is None else fy
        K = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=np.float32)
        return Camera(K, np.asarray(T_camera_world, dtype=np.float32), image_width, image_height)

    def compute_fov_x_radians(self) -> float:  # :79-82
        return 2.0 * math.atan((self.image_width / 2.0) / float(self.K[0, 0]))

    def compute_fov_y_radians(self) -> float:  # :84-87
        return 2.0 * math.atan((self.image_height / 2.0) / float(self.K[1, 1]))

    def resize_with_fixed_fov(self, image_width: int, image_height: int) -> "Camera":  # :89-99
        return Camera.from_fov(self.T_camera_world, image_width, image_height, self.compute_fov_x_radians(),
                               self.compute_fov_y_radians())

    def ray_matrices(self) -> Tuple[np.ndarray, np.ndarray]:
        """(M, origin) with M = R_world_camera @ K^-1 in fp32 (cameras.py:107-113) and origin =
        T_world_camera.translation() = -R^T t (:118)."""
        T = np.asarray(self.T_camera_world, dtype=np.float32)
        R_wc = T[:3, :3].T
        origin = -(R_wc @ T[:3, 3])
        M = (R_wc @ np.linalg.inv(np.asarray(self.K, dtype=np.float32))).astype(np.float32)
        return M, origin.astype(np.float32)

    def pixel_rays_wrt_world(self, camera_index: int, device="cuda", rows: Optional[Tuple[int, int]] = None) -> Rays3D:
        """cameras.py:124-143: rays of every pixel, batch axes (H, W) — or of the row band `rows` (image tiles
        for multi-GPU rendering) — generated on the device."""
        from . import ops
        M, origin = self.ray_matrices()
        r0, r1 = rows if rows is not None else (0, self.image_height)
        o, d, c = ops.pixel_rays(M.reshape(-1), origin, self.image_width, (r0, r1), int(camera_index), device)
        return Rays3D(o, d, c).reshape(r1 - r0, self.image_width)

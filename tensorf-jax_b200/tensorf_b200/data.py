"""Mirror of the training data path (tensorf/data.py:284-337, training.py:318-343) with the ray table
resident on the device.

The reference builds one `RenderedRays` of every pixel of every training view on the host
(`rendered_rays_from_views`, data.py:301-337), keeps it in a `fifteen.data.InMemoryDataLoader` and ships
a shuffled minibatch (40 B/ray) to the device every step.  Here the table is built on the device
(`tensorf_pixel_rays`, `tensorf_rgba_over_white`) and a minibatch is `tensorf_gather_rays` by a shuffled index
— a step moves no ray data across PCIe.

`fifteen` (editable `../fifteen`, unpinned) is not under /root/reference; its loader is restated as: every epoch
a fresh permutation of the table (seeded `shuffle_seed + epoch`), consecutive minibatches of `minibatch_size`,
the incomplete tail dropped, cycled forever.
"""
from __future__ import annotations

import dataclasses
from typing import Iterator, List, Optional

import numpy as np
import torch

from . import _lib, cameras, ops
from ._lib import check
from .training import RenderedRays


@dataclasses.dataclass
class RegisteredRgbaView:  # data.py:279-287
    image_rgba: torch.Tensor  # (H, W, 4) fp32 in [0,1]
    camera: cameras.Camera


def rgba_over_white(rgba: torch.Tensor) -> torch.Tensor:
    """data.py:318-320."""
    rgba = rgba.to(torch.float32).contiguous()
    n = rgba.numel() // 4
    out = torch.empty((n, 3), dtype=torch.float32, device=rgba.device)
    check(_lib.load().tensorf_rgba_over_white(ops._stream(), ops._ptr(rgba, name="rgba"), n, out.data_ptr()))
    return out


def rendered_rays_from_views(views: List[RegisteredRgbaView], device="cuda") -> RenderedRays:
    """data.py:301-337: one flat RenderedRays over all pixels of all views; camera_index = view position."""
    o, d, c, col = [], [], [], []
    for i, view in enumerate(views):
        h, w = view.camera.image_height, view.camera.image_width
        assert tuple(view.image_rgba.shape) == (h, w, 4)
        rays = view.camera.pixel_rays_wrt_world(camera_index=i, device=device)
        assert rays.get_batch_axes() == (h, w)
        o.append(rays.origins.reshape(-1, 3))
        d.append(rays.directions.reshape(-1, 3))
        c.append(rays.camera_indices.reshape(-1))
        col.append(rgba_over_white(view.image_rgba.to(device)))
    return RenderedRays(colors=torch.cat(col), rays_wrt_world=cameras.Rays3D(torch.cat(o), torch.cat(d), torch.cat(c)))


class DeviceRayLoader:
    """In-memory (device-memory) minibatch loader: `for minibatch in loader.cycled(shuffle_seed=0)`."""

    def __init__(self, dataset: RenderedRays, minibatch_size: int):
        (self.n,) = dataset.get_batch_axes()
        if minibatch_size < 1 or minibatch_size > self.n:
            raise ValueError(f"minibatch_size {minibatch_size} outside [1, {self.n}]")
        r = dataset.rays_wrt_world
        self.origins = r.origins.to(torch.float32).contiguous()
        self.directions = r.directions.to(torch.float32).contiguous()
        self.cams = r.camera_indices.to(torch.int32).contiguous()
        self.colors = dataset.colors.to(torch.float32).contiguous()
        self.minibatch_size = int(minibatch_size)
        self.device = self.origins.device
        self._bad = torch.zeros((), dtype=torch.int32, device=self.device)

    def minibatch_count(self) -> int:
        return self.n // self.minibatch_size

    def gather(self, idx: torch.Tensor) -> RenderedRays:
        """Rows `idx` (int64, on the device) of the table."""
        idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
        R = idx.numel()
        o = torch.empty((R, 3), dtype=torch.float32, device=self.device)
        d = torch.empty((R, 3), dtype=torch.float32, device=self.device)
        c = torch.empty((R,), dtype=torch.int32, device=self.device)
        col = torch.empty((R, 3), dtype=torch.float32, device=self.device)
        check(_lib.load().tensorf_gather_rays(ops._stream(), self.origins.data_ptr(), self.directions.data_ptr(),
                                              self.cams.data_ptr(), self.colors.data_ptr(), self.n, idx.data_ptr(), R,
                                              o.data_ptr(), d.data_ptr(), c.data_ptr(), col.data_ptr(), self._bad.data_ptr()))
        return RenderedRays(colors=col, rays_wrt_world=cameras.Rays3D(o, d, c))

    def bad_index_count(self) -> int:
        """Out-of-table indices seen by the last gather (blocking read)."""
        return int(self._bad.item())

    def epoch_permutation(self, shuffle_seed: int, epoch: int) -> torch.Tensor:
        perm = np.random.default_rng(shuffle_seed + epoch).permutation(self.n)
        return torch.from_numpy(perm).to(self.device)       # 8 B/ray once per epoch

    def cycled(self, shuffle_seed: Optional[int] = 0) -> Iterator[RenderedRays]:
        """training.py:320-323 `cycled_minibatches(dataloader, shuffle_seed=0)`."""
        epoch = 0
        while True:
            perm = self.epoch_permutation(shuffle_seed, epoch) if shuffle_seed is not None else torch.arange(self.n, device=self.device)
            for b in range(self.minibatch_count()):
                yield self.gather(perm[b * self.minibatch_size:(b + 1) * self.minibatch_size])
            epoch += 1


class HostStage:
    """Host-resident minibatches (the reference's loader hands NumPy batches to `training_step`,
    training.py:318-343): every per-step input lives in ONE pinned host buffer mirrored by ONE device buffer, so a
    step's inputs cross PCIe as a single copy instead of one per array (each small copy costs a fixed ~2-3 us of
    stream time).  `host[name]` / `device[name]` are views; all arrays must be 4-byte typed (fp32, int32, uint32)."""

    def __init__(self, arrays, device="cuda"):
        import numpy as np
        metas, off = [], 0
        for k, a in arrays.items():
            t = torch.from_numpy(np.ascontiguousarray(a)) if not isinstance(a, torch.Tensor) else a.contiguous().cpu()
            if t.dtype == torch.uint32:
                t = t.view(torch.int32)
            if t.element_size() != 4:
                raise TypeError(f"HostStage: '{k}' has dtype {t.dtype}; only 4-byte element types are staged")
            metas.append((k, t, off))
            off += (t.numel() + 3) // 4 * 4        # every view starts on a 16-byte boundary
        self._host = torch.empty(max(off, 4), dtype=torch.float32)
        if torch.device(device).type == "cuda":
            self._host = self._host.pin_memory()
        self._dev = torch.empty(max(off, 4), dtype=torch.float32, device=device)
        self.host, self.device = {}, {}
        for k, t, o in metas:
            n = t.numel()
            self.host[k] = self._host[o:o + n].view(t.dtype).view(t.shape)
            self.device[k] = self._dev[o:o + n].view(t.dtype).view(t.shape)
            self.host[k].copy_(t)
        self.nbytes = off * 4

    def upload(self) -> None:
        """One asynchronous host-to-device copy of every staged array, on the current stream."""
        self._dev.copy_(self._host, non_blocking=True)

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

from typing import Dict, List, Sequence, Tuple

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


def balanced_stripe(height: int, world: int, largest: int = 16) -> int:
    """Largest stripe height <= `largest` (halving) for which every rank gets the same number of stripes of a frame:
    800 rows over 8 ranks -> 4 rows (25 stripes each); 16-row stripes would give 6 or 7 per rank, i.e. the slowest rank
    renders 12 % more than the mean.  Falls back to `largest` when no power-of-two fraction divides evenly."""
    s = largest
    while s >= 1:
        if height % s == 0 and (height // s) % world == 0:
            return s
        s //= 2
    return largest


def stripe_rows(height: int, rank: int, world: int, stripe: int = 16) -> List[Tuple[int, int]]:
    """Rows of a frame rendered by `rank` as interleaved stripes: stripe s = rows [s*stripe, (s+1)*stripe) belongs to
    rank s % world.  Rays through the middle of the frame cross more of the scene than those near its border, so
    contiguous bands (`tile_rows`) leave the central ranks with the longest job; 16-row stripes dealt round-robin give
    every rank the same mix.  No collective: every rank writes its own rows of the frame."""
    out = []
    for s in range((height + stripe - 1) // stripe):
        if s % world == rank:
            out.append((s * stripe, min(height, (s + 1) * stripe)))
    return out


class FlatGrads:
    """All gradient leaves as views into one contiguous fp32 buffer, plus (optionally) one slot for the loss, so
    the exchange step is one or two collectives instead of one per leaf.

    Two buckets follow the order in which the reverse pass finishes its leaves (`RenderCall.backward(phase=1/2)`):
    `early` = everything but the density factors (+ the loss slot), final after phase 1, exchanged while phase 2
    (the density scatter) runs; `late` = the density factors.  Needs the density leaves first in `shapes`
    (`ops.param_shapes` order); otherwise there is a single bucket."""

    def __init__(self, shapes: Dict[str, Tuple[int, ...]], device, loss_slot: bool = False):
        self.shapes = dict(shapes)
        self.total = int(sum(int(np.prod(s)) for s in shapes.values()))
        self.buffer = torch.zeros(self.total + (1 if loss_slot else 0), dtype=torch.float32, device=device)
        self.flat = self.buffer[:self.total]
        self.loss = self.buffer[self.total:self.total + 1] if loss_slot else None
        self.leaves: Dict[str, torch.Tensor] = {}
        off, split, prefix = 0, 0, True
        for k, s in shapes.items():
            n = int(np.prod(s))
            self.leaves[k] = self.flat[off:off + n].view(s)
            off += n
            if prefix and k.startswith("density_"):
                split = off
            else:
                prefix = False
        self.late = self.buffer[:split]
        self.early = self.buffer[split:]
        self._pending = []

    def allreduce(self, group=None) -> None:
        """Sum over ranks (training.py:140 is a mean over the whole batch: every rank already
        scaled its cotangent by 1/(3*R_global), so the reduction is a plain sum)."""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group)

    def start_allreduce(self, which: str, group=None) -> None:
        """Asynchronous sum of one bucket ("early" / "late"): the collective is ordered after the work already
        enqueued on the current stream and runs beside whatever is enqueued next; `finish()` joins."""
        import torch.distributed as dist

        t = {"early": self.early, "late": self.late}[which]
        if t.numel() and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            self._pending.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=True))

    def finish(self) -> None:
        for w in self._pending:
            w.wait()
        self._pending = []


def global_loss_scale(local_rays: int, world: int) -> float:
    """1/(3*R_global) for equal shards."""
    return 1.0 / (3.0 * local_rays * world)


class PeerAdam:
    """Gradient reduce-scatter + Adam + parameter all-gather as ONE kernel over peer memory
    (`tensorf_adam_step_peer`, SURVEY §8e "fused follow-up"): replaces `FlatGrads.allreduce()` +
    `ops.AdamCall.step()` (training.py:153-156 -> :158-243 with the ray batch sharded over ranks).

    Every leaf of LearnableParams and of its gradient lives in one symmetric allocation per rank
    (`torch.distributed._symmetric_memory`: cuMem handles exchanged at rendezvous, so every rank holds a
    mapping of every other rank's buffer, plus the NVSwitch multicast address when the fabric has one):
    `[params: T | grads: T | aux: 16 (aux[0] = loss) | norm slots: 16 | signal pad: 32 x u32]` floats, T = leaves (each 16-byte aligned)
    rounded up to a multiple of 4.  `allreduce()` is the exchange alone (`tensorf_peer_allreduce` over grads + aux, on
    the launch stream) for callers that keep their own optimiser.
    Rank r owns elements `[shard_begin, shard_end)`; its shard of the Adam moments is local (memory and
    traffic / world).  `params` / `grads` are views for the render calls (`RenderCall.backward(None, grads)`
    writes straight into the symmetric buffer).  world == 1 needs no process group and runs the same kernel.
    """

    def __init__(self, shapes: Dict[str, Tuple[int, ...]], neg_lrs: Dict[str, float], device, group=None,
                 b1: float = 0.9, b2: float = 0.99, eps: float = 1e-8, eps_root: float = 0.0, multicast=None):
        import ctypes as C
        import os

        import torch.distributed as dist

        from . import _lib

        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("PeerAdam needs a CUDA device (there is no CPU path)")
        self.names = list(shapes.keys())
        self.shapes = {k: tuple(int(x) for x in shapes[k]) for k in self.names}
        sizes = [int(np.prod(self.shapes[k])) for k in self.names]
        # every leaf starts on a 16-byte boundary (the render kernels take vector paths on aligned leaves); a gap
        # is a pseudo-leaf with learning rate 0 whose gradient is never written (stays 0)
        self.leaf_offsets: Dict[str, int] = {}
        table_offs, table_lrs, off = [0], [], 0
        for k, n in zip(self.names, sizes):
            self.leaf_offsets[k] = off
            if n > 0:
                table_offs.append(off + n)
                table_lrs.append(float(neg_lrs[k]))
            off += n
            pad = -off % 4
            if pad and k != self.names[-1]:
                off += pad
                table_offs.append(off)
                table_lrs.append(0.0)
        if not table_lrs:
            raise ValueError("PeerAdam: no parameters")
        if len(table_lrs) > _lib.PEER_MAX_LEAVES:
            raise ValueError(f"PeerAdam: {len(table_lrs)} leaves and alignment gaps (max {_lib.PEER_MAX_LEAVES})")
        self.leaf_total = int(table_offs[-1])
        self.total = (self.leaf_total + 3) // 4 * 4
        self.b1, self.b2, self.eps, self.eps_root = b1, b2, eps, eps_root
        use_dist = dist.is_available() and dist.is_initialized()
        self.group = group if group is not None else (dist.group.WORLD if use_dist else None)
        self.world = dist.get_world_size(self.group) if use_dist else 1
        self.rank = dist.get_rank(self.group) if use_dist else 0
        if self.world > _lib.PEER_MAX_WORLD:
            raise ValueError(f"PeerAdam: world {self.world} > {_lib.PEER_MAX_WORLD}")
        n_all = 2 * self.total + 64
        self._hdl = None
        mc_base = 0
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm_mem

            self.buf = symm_mem.empty(n_all, dtype=torch.float32, device=self.device)
            self._hdl = symm_mem.rendezvous(self.buf, self.group)
            bases = [int(x) for x in self._hdl.buffer_ptrs]
            own_off = self.buf.data_ptr() - bases[self.rank]  # 0 unless the tensor sits inside a pooled block
            if own_off < 0:
                raise RuntimeError("PeerAdam: symmetric allocation does not contain its own tensor")
            peer_ptrs = [b + own_off for b in bases]
            # transport: P2P loads / stores by default (measured faster than the switch reduction at this payload,
            # profiles/r01_v17_peer_exchange_2gpu.json); multicast=True or TENSORF_PEER_MULTICAST=1 selects NVLS
            env = os.environ.get("TENSORF_PEER_MULTICAST")
            want_mc = multicast if multicast is not None else (env is not None and env != "0")
            mc = int(self._hdl.multicast_ptr or 0)
            if want_mc and mc == 0:
                raise RuntimeError("PeerAdam: multicast requested but the symmetric allocation has no multicast address")
            if mc:
                mc_base = mc + own_off
            self._want_mc = bool(want_mc)
        else:
            self.buf = torch.empty(n_all, dtype=torch.float32, device=self.device)
            peer_ptrs = [self.buf.data_ptr()]
        self.buf.zero_()
        self.has_multicast = mc_base != 0
        self.multicast = self.has_multicast and getattr(self, "_want_mc", False)  # default transport of allreduce() / step()
        self.params_flat = self.buf[:self.total]
        self.grads_flat = self.buf[self.total:2 * self.total]
        self.aux = self.buf[2 * self.total:2 * self.total + 16]
        self.loss = self.aux[0:1]
        self.slots = self.buf[2 * self.total + 16:2 * self.total + 32]
        self.params: Dict[str, torch.Tensor] = {}
        self.grads: Dict[str, torch.Tensor] = {}
        for k, n in zip(self.names, sizes):
            o = self.leaf_offsets[k]
            self.params[k] = self.params_flat[o:o + n].view(self.shapes[k])
            self.grads[k] = self.grads_flat[o:o + n].view(self.shapes[k])
        # end of the leading density_* leaves (a multiple of 4: every leaf is 16-byte aligned): the "late" bucket
        self._density_end = 0
        for k, n in zip(self.names, sizes):
            if not k.startswith("density_"):
                break
            self._density_end = (self.leaf_offsets[k] + n + 3) // 4 * 4
        b, e = C.c_int64(), C.c_int64()
        self.lib.tensorf_peer_shard(self.total, self.rank, self.world, C.byref(b), C.byref(e))
        self.shard = (int(b.value), int(e.value))
        n_shard = self.shard[1] - self.shard[0]
        self.mu = torch.zeros(max(n_shard, 4), dtype=torch.float32, device=self.device)
        self.nu = torch.zeros(max(n_shard, 4), dtype=torch.float32, device=self.device)
        nbytes = int(self.lib.tensorf_peer_adam_scratch_bytes(n_shard))
        self.scratch = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.device)
        self.grad_norm = torch.zeros((), dtype=torch.float32, device=self.device)
        W = self.world
        self._n_table = len(table_lrs)
        self._offs = (C.c_int64 * (self._n_table + 1))(*table_offs)
        self._neg_lrs = (C.c_float * self._n_table)(*table_lrs)
        self._g = (C.c_void_p * W)(*[p + 4 * self.total for p in peer_ptrs])
        self._p = (C.c_void_p * W)(*peer_ptrs)
        self._s = (C.c_void_p * W)(*[p + 8 * self.total + 64 for p in peer_ptrs])
        self._x_mc = C.c_void_p(mc_base + 4 * self.total) if mc_base else None
        # in-kernel ordering of tensorf_peer_allreduce_sync: signal pads (zeroed with the buffer), local gate, call count
        self._sig = (C.c_void_p * W)(*[p + 8 * self.total + 128 for p in peer_ptrs])
        self._local_flags = torch.zeros(4, dtype=torch.int32, device=self.device)
        self._epoch = 0
        # "barrier" (two torch symmetric-memory barrier launches around the kernel; measured on 2 and 4 GPUs) or
        # "kernel" (both barriers inside tensorf_peer_allreduce_sync: 37 vs 40 us at 2 GPUs, measured on 2 GPUs only)
        self.sync = os.environ.get("TENSORF_PEER_SYNC", "barrier")
        if self.sync not in ("barrier", "kernel"):
            raise ValueError(f"TENSORF_PEER_SYNC={self.sync!r}: expected 'barrier' or 'kernel'")
        self._g_mc = C.c_void_p(mc_base + 4 * self.total) if mc_base else None
        self._p_mc = C.c_void_p(mc_base) if mc_base else None
        self.barrier()  # every rank's buffer is zeroed before anyone may store into it

    def barrier(self) -> None:
        """Cross-rank barrier ordered on the current stream (device-side signal pads, no host wait)."""
        if self._hdl is not None:
            self._hdl.barrier(channel=0)

    def autotune_transport(self, reps: int = 5) -> Dict[str, float]:
        """Times the exchange of the whole buffer over both transports (P2P loads / stores vs NVSwitch multicast
        `multimem.ld_reduce` / `multimem.st`) and makes the faster one (max over ranks) the default.  Which one wins depends
        on the world size and the payload: P2P at 2-4 GPUs and 12.8 MB, the in-switch reduction for larger worlds / payloads.
        The buffer's contents are summed `2 * (reps + 1)` times: call before the first step.  Returns the timings (us)."""
        import torch.distributed as dist
        out = {}
        if self.world == 1 or not self.has_multicast:
            return out
        for name, mc in (("p2p", False), ("multicast", True)):
            self.multicast = mc
            self.grads_flat.zero_()
            self.allreduce()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            dist.barrier(group=self.group)
            e0.record()
            for _ in range(reps):
                self.allreduce()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], dtype=torch.float64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            out[name] = float(t.item())
        self.multicast = out["multicast"] < out["p2p"]
        self.grads_flat.zero_()
        self.aux.zero_()
        return out

    def allreduce(self, which: str = "all", channel: int = 0, max_ctas: int = 0) -> None:
        """Sum `grads` and `aux` (the loss slot) over the ranks in place, on the current stream: barrier (all
        gradients written), one kernel (each rank reduces 1/world of the range from every peer and stores the sums
        to every peer), barrier (all stores landed).  `which` = "late" (the leading density leaves) / "early"
        (everything else + aux) reduces one bucket of the two-half reverse pass (`RenderCall.backward(phase=1/2)`);
        buckets running on different streams must use different barrier channels."""
        import ctypes as C

        from . import _lib
        from .ops import _stream

        if self.world == 1:
            return
        if which == "all" and self.sync == "kernel":  # both barriers inside the kernel (signal pads in the symmetric buffer)
            # epoch 0: the kernel keeps the call number in its own device memory, so the launch is CUDA-graph capturable
            _lib.check(self.lib.tensorf_peer_allreduce_sync(_stream(), self.rank, self.world, self.total + 16, self._g,
                                                            self._x_mc if self.multicast else None, self._sig, self._local_flags.data_ptr(), 0))
            return
        lo, hi = {"all": (0, self.total + 16), "late": (0, self._density_end), "early": (self._density_end, self.total + 16)}[which]
        if hi == lo:
            return
        ptrs = self._g if lo == 0 else (C.c_void_p * self.world)(*[int(p) + 4 * lo for p in self._g])
        mc = None if not self.multicast else (self._x_mc if lo == 0 else C.c_void_p(self._x_mc.value + 4 * lo))
        if self._hdl is not None:
            self._hdl.barrier(channel=channel)
        if max_ctas:  # an exchange that overlaps a compute kernel: a few CTAs only
            _lib.check(self.lib.tensorf_peer_set_max_ctas(int(max_ctas)))
        try:
            _lib.check(self.lib.tensorf_peer_allreduce(_stream(), self.rank, self.world, hi - lo, ptrs, mc))
        finally:
            if max_ctas:
                _lib.check(self.lib.tensorf_peer_set_max_ctas(0))
        if self._hdl is not None:
            self._hdl.barrier(channel=channel)

    def load_params(self, flat: Dict[str, torch.Tensor]) -> None:
        """Copy replicated leaves into the symmetric buffer (every rank calls this with the same values)."""
        with torch.no_grad():
            for k in self.names:
                self.params[k].copy_(flat[k].detach())

    def step(self, count: int, lr_decay: float = 1.0) -> torch.Tensor:
        """One fused exchange + optimiser step over `self.grads` (each rank's LOCAL gradient, already scaled by
        the global 1/(3R)); afterwards `self.params` holds the same new parameters on every rank.  Returns the
        device scalar optax.global_norm(sum of the ranks' gradients)."""
        import ctypes as C

        from . import _lib
        from .ops import _stream

        t = np.float32(count + 1)
        bc1 = np.float32(1) - np.power(np.float32(self.b1), t)
        bc2 = np.float32(1) - np.power(np.float32(self.b2), t)
        d = _lib.PeerAdamDesc(
            adam=_lib.AdamDesc(n_leaves=self._n_table, reserved=0, b1=self.b1, b2=self.b2, eps=self.eps,
                               eps_root=self.eps_root, bias_correction1=float(bc1), bias_correction2=float(bc2),
                               lr_decay=float(lr_decay), reserved2=0.0),
            rank=self.rank, world=self.world, total=self.total, shard_begin=self.shard[0], shard_end=self.shard[1])
        self.barrier()  # all ranks' gradients are complete
        _lib.check(self.lib.tensorf_adam_step_peer(_stream(), C.byref(d), self._offs, self._neg_lrs, self._g, self._p,
                                                   self._g_mc if self.multicast else None, self._p_mc if self.multicast else None, self.mu.data_ptr(), self.nu.data_ptr(), self._s,
                                                   self.scratch.data_ptr(), self.scratch.numel()))
        self.barrier()  # all parameter and slot stores have landed everywhere
        _lib.check(self.lib.tensorf_peer_grad_norm(_stream(), self.slots.data_ptr(), self.world, self.grad_norm.data_ptr()))
        return self.grad_norm

    def gather_moments(self, shard: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Full moment leaves from the per-rank shards (grid resampling needs them whole, training.py:245-276)."""
        full = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        b, e = self.shard
        full[b:e] = shard[:e - b]
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(full, group=self.group)  # shards are disjoint: the sum is the concatenation
        out = {}
        for k in self.names:
            o, n = self.leaf_offsets[k], int(np.prod(self.shapes[k]))
            out[k] = full[o:o + n].view(self.shapes[k])
        return out

    def scatter_moments(self, shard: torch.Tensor, leaves: Dict[str, torch.Tensor]) -> None:
        """Inverse of gather_moments: keep this rank's range of replicated full leaves."""
        full = torch.zeros(self.total, dtype=torch.float32, device=self.device)
        for k in self.names:
            o, n = self.leaf_offsets[k], int(np.prod(self.shapes[k]))
            full[o:o + n] = leaves[k].reshape(-1)
        b, e = self.shard
        shard[:e - b] = full[b:e]

"""Mirror of tensorf/training.py:19-243 (TrainState.initialize / training_step) over the CUDA
library.  The forward + reverse (training.py:108-156) is the hot path and runs entirely in
libtensorf_b200.so (fused MSE, no autograd graph); Adam + LR decay + apply + grad_norm
(training.py:158-204, optax scale_by_adam b1=.9 b2=.99 eps=1e-8, masked group LRs) is ONE launch of
`tensorf_adam_step` over all leaves; `resize_grid` (training.py:245-276) resamples parameters and
both Adam moments with `tensorf_vm_resize`.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import cameras, networks, ops, prng, render, tensor_vm, train_config


@dataclasses.dataclass
class RenderedRays:  # data.py:290-298
    colors: torch.Tensor  # (R,3)
    rays_wrt_world: cameras.Rays3D

    def get_batch_axes(self) -> Tuple[int, ...]:
        return self.rays_wrt_world.get_batch_axes()


def psnr_from_mse(mse: float) -> float:  # utils.py:1-9
    return -10.0 * math.log10(mse) if mse > 0 else float("inf")


@dataclasses.dataclass
class TrainState:
    config: train_config.TensorfConfig
    appearance_mlp: networks.FeatureMlp
    learnable_params: render.LearnableParams
    optimizer_state: Dict[str, Dict[str, torch.Tensor]]  # {"mu": leaves, "nu": leaves}
    aabb: torch.Tensor
    prng_key: prng.Key
    step: int = 0
    world_size: int = 1  # rays sharded over ranks; gradients sum-allreduced
    _adam: Optional[ops.AdamCall] = dataclasses.field(default=None, repr=False, compare=False)
    _adam_names: Optional[list] = dataclasses.field(default=None, repr=False, compare=False)
    _peer: Optional[object] = dataclasses.field(default=None, repr=False, compare=False)  # dist.PeerAdam
    _fg: Optional[object] = dataclasses.field(default=None, repr=False, compare=False)    # dist.FlatGrads (world_size > 1)

    @property
    def rank(self) -> int:
        """This process's position in the sharded ray batch (0 when not distributed)."""
        import torch.distributed as dist
        return dist.get_rank() if (self.world_size > 1 and dist.is_available() and dist.is_initialized()) else 0

    @staticmethod
    def initialize(config: train_config.TensorfConfig, grid_dim: int, prng_key, num_cameras: int,
                   device="cuda", world_size: int = 1) -> "TrainState":
        """training.py:35-99.  The initial values come from a torch generator seeded from the key (every rank must pass
        the same key): the distributions are the reference's (networks.py:9-10 variance-scaling truncated normals,
        N(0, 0.1^2) factors), the random streams are torch's, not jax.random's."""
        if isinstance(prng_key, prng.Key):  # split(key, 5)[0] as training.py:44-60 does for the first sub-key; its words seed torch
            k = prng.split(prng_key, 5)[0]
            seed = (int(k.k0) << 32) | int(k.k1)
        else:
            seed = int(prng_key)
        gen = torch.Generator(device=device).manual_seed(seed & ((1 << 63) - 1))
        mlp = networks.FeatureMlp(feature_n_freqs=config.feature_n_freqs, viewdir_n_freqs=config.viewdir_n_freqs,
                                  num_cameras=num_cameras if config.camera_embeddings else None)
        dummy = torch.zeros((1, config.appearance_feat_dim * 3), device=device)
        mlp_params = mlp.init(gen, dummy)
        lp = render.LearnableParams(
            appearance_mlp_params=mlp_params,
            appearance_tensor=tensor_vm.TensorVM.initialize(grid_dim, config.appearance_feat_dim, prng_key=gen, device=device),
            density_tensor=tensor_vm.TensorVM.initialize(grid_dim, config.density_feat_dim, prng_key=gen, device=device),
            scene_contraction=config.scene_contraction)
        flat = lp.flat()
        opt = {"mu": {k: torch.zeros_like(v) for k, v in flat.items()}, "nu": {k: torch.zeros_like(v) for k, v in flat.items()}}
        aabb = torch.tensor([config.initial_aabb_min, config.initial_aabb_max], dtype=torch.float32, device=device)
        key = prng_key if isinstance(prng_key, prng.Key) else prng.Key.from_seed(seed)
        return TrainState(config, mlp, lp, opt, aabb, prng.split(key, 5)[4], 0, world_size)

    # ---- hot path -------------------------------------------------------------------------------
    def sample_counts(self) -> Tuple[int, int]:
        """training.py:113-118."""
        g = self.learnable_params.appearance_tensor.grid_dim()
        assert g == self.learnable_params.density_tensor.grid_dim()
        n = int(math.sqrt(3 * g**2) * self.config.train_ray_sample_multiplier)
        return n, int(0.15 * n)

    def loss_and_grads(self, minibatch: RenderedRays, render_prng_key, grads: Optional[Dict[str, torch.Tensor]] = None,
                       flat_grads=None):
        """training.py:108-156: (mse, grads w.r.t. every LearnableParams leaf). The mean is over the
        GLOBAL batch (`world_size` x local rays); the caller allreduces the gradients - or passes
        `flat_grads` (dist.FlatGrads with a loss slot): then the reverse pass runs in its two halves, the exchange
        of everything but the density factors overlaps with the density scatter, and loss and gradients come back
        already summed over the ranks."""
        (R,) = minibatch.get_batch_axes()
        N, K = self.sample_counts()
        flat = self.learnable_params.flat()
        mlp = self.appearance_mlp
        desc = ops.make_desc(R=R, N=N, K=K, G=self.learnable_params.density_tensor.grid_dim(),
                             cd=flat["density_vector"].shape[1], ca=flat["appearance_vector"].shape[1],
                             contracted=self.config.scene_contraction, feat_freqs=mlp.feature_n_freqs,
                             view_freqs=mlp.viewdir_n_freqs, num_cameras=mlp.num_cameras, squash=mlp.feature_squash_dim,
                             units=mlp.units, loss_scale=1.0 / (3 * R * self.world_size))
        dev = self.aabb.device
        rays = minibatch.rays_wrt_world
        inputs = {"origins": rays.origins.contiguous(), "directions": rays.directions.contiguous(),
                  "camera_indices": rays.camera_indices.to(torch.int32).contiguous(), "aabb": self.aabb,
                  "colors": minibatch.colors.contiguous()}
        # sharded batches: rank r owns rows [r*R, (r+1)*R) of the global batch and draws those rows of the global jitter
        inputs.update(render._noise_inputs(render_prng_key, R, N, self.config.scene_contraction, True, dev, first_ray=self.rank * R))
        if self.config.scene_contraction:
            base, delta = render.contracted_schedule(self.config.render_near, self.config.render_far, N)
            inputs["base_ts"], inputs["deltas"] = torch.from_numpy(base).to(dev), torch.from_numpy(delta).to(dev)
        call = render._acquire(desc, dev)
        if flat_grads is not None:
            fg = flat_grads
            _, loss = call.forward({k: v.contiguous() for k, v in flat.items()}, inputs, loss_out=fg.loss)
            call.backward(None, fg.leaves, phase=1)
            fg.start_allreduce("early")
            call.backward(None, fg.leaves, phase=2)
            fg.start_allreduce("late")
            fg.finish()
            render._release(call)
            return loss, fg.leaves
        _, loss = call.forward({k: v.contiguous() for k, v in flat.items()}, inputs)
        grads = call.backward(None, grads)
        render._release(call)
        return loss, grads

    def enable_peer_optimizer(self, group=None, multicast=None) -> "TrainState":
        """SURVEY §8e fused follow-up: move the parameters into a symmetric peer-mapped buffer and replace
        `all_reduce(grads)` + Adam by `tensorf_adam_step_peer` (one kernel: reduce-scatter over NVLink, Adam on the
        owner's shard, parameter all-gather).  The Adam moments become per-rank shards (`optimizer_state` keeps
        {"mu": {"shard": ...}, "nu": {"shard": ...}}); call before the first step (moments must still be zero)
        and again after `resize_grid`."""
        from . import dist as tdist
        if self.step != 0 and self._peer is None:
            raise ValueError("enable_peer_optimizer: call before the first training step (moments are re-created)")
        oc = self.config.optimizer
        flat = self.learnable_params.flat()
        neg = {k: -(oc.lr_init_tensor if k.startswith(("density_", "appearance_")) else oc.lr_init_mlp) for k in flat}
        peer = tdist.PeerAdam({k: tuple(v.shape) for k, v in flat.items()}, neg, self.aabb.device, group=group,
                              b1=0.9, b2=0.99, eps=1e-8, multicast=multicast)
        if peer.world != self.world_size:
            raise ValueError(f"enable_peer_optimizer: process group has {peer.world} ranks, world_size={self.world_size}")
        peer.load_params(flat)
        peer.barrier()
        self.learnable_params = render.LearnableParams.from_flat(peer.params, self.config.scene_contraction)
        self.optimizer_state = {"mu": {"shard": peer.mu}, "nu": {"shard": peer.nu}}
        self._peer, self._adam = peer, None
        return self

    def training_step(self, minibatch: RenderedRays) -> Tuple["TrainState", Dict[str, float]]:
        """training.py:101-205."""
        keys = prng.split(self.prng_key)
        render_key, new_key = keys[0], keys[1]
        if self._peer is not None:
            return self._training_step_peer(minibatch, render_key, new_key)
        if self.world_size > 1:
            from . import dist as tdist
            shapes = {k: tuple(v.shape) for k, v in self.learnable_params.flat().items()}
            if self._fg is None or self._fg.shapes != shapes:
                self._fg = tdist.FlatGrads(shapes, self.aabb.device, loss_slot=True)
            loss, grads = self.loss_and_grads(minibatch, render_key, flat_grads=self._fg)
        else:
            loss, grads = self.loss_and_grads(minibatch, render_key)
        coeff = self.lr_decay_coeff()
        oc = self.config.optimizer
        # optax.scale_by_adam(b1=0.9, b2=0.99, eps=1e-8) + masked group learning rates (:213-243), one launch
        flat = self.learnable_params.flat()
        names = list(flat.keys())
        if self._adam is None or self._adam_names != names or any(
                a.data_ptr() != flat[k].data_ptr() for a, k in zip(self._adam.params, names)):
            neg_lrs = [-(oc.lr_init_tensor if k.startswith(("density_", "appearance_")) else oc.lr_init_mlp) for k in names]
            self._adam = ops.AdamCall([flat[k] for k in names], [self.optimizer_state["mu"][k] for k in names],
                                      [self.optimizer_state["nu"][k] for k in names], neg_lrs, b1=0.9, b2=0.99, eps=1e-8)
            self._adam_names = names
        gnorm = self._adam.step([grads[k] for k in names], count=self.step, lr_decay=coeff)
        step = self.step
        self.prng_key = new_key
        self.step = step + 1
        mse = float(loss.item())  # the blocking read the reference does at training.py:342
        log = {"train/mse": mse, "train/psnr": psnr_from_mse(mse), "train/lr_tensor": coeff * oc.lr_init_tensor,
               "train/lr_mlp": coeff * oc.lr_init_mlp, "train/grad_norm": float(gnorm.item())}
        return self, log

    def _training_step_peer(self, minibatch: RenderedRays, render_key, new_key):
        """training_step with the exchange and the optimiser fused (`enable_peer_optimizer`)."""
        peer = self._peer
        loss, _ = self.loss_and_grads(minibatch, render_key, grads=peer.grads)  # reverse pass writes the symmetric buffer
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(loss)
        coeff = self.lr_decay_coeff()
        oc = self.config.optimizer
        gnorm = peer.step(count=self.step, lr_decay=coeff)
        self.prng_key = new_key
        self.step += 1
        mse = float(loss.item())
        log = {"train/mse": mse, "train/psnr": psnr_from_mse(mse), "train/lr_tensor": coeff * oc.lr_init_tensor,
               "train/lr_mlp": coeff * oc.lr_init_mlp, "train/grad_norm": float(gnorm.item())}
        return self, log

    def lr_decay_coeff(self) -> float:
        """training.py:158-181: optax.exponential_decay(1.0, decay_iters, ratio, end_value=ratio) at the step
        count since the last upsampling (when `lr_upsample_reset`)."""
        cfg, oc = self.config, self.config.optimizer
        step = self.step
        if oc.lr_upsample_reset:
            resetted = min(step - u for u in (0,) + tuple(cfg.upsamp_iters) if step - u >= 0)
        else:
            resetted = step
        decay_iters = oc.lr_decay_iters if oc.lr_decay_iters is not None else cfg.n_iters
        return max(oc.lr_decay_target_ratio ** (resetted / decay_iters), oc.lr_decay_target_ratio)

    def resize_grid(self, new_grid_dim: int) -> "TrainState":
        """training.py:245-276: resample the factor grids and their Adam moments (mu, nu)."""
        if self._peer is not None:
            return self._resize_grid_peer(new_grid_dim)
        lp = self.learnable_params
        # one scratch buffer for the six resamplings (parameters and both moments of both factor sets): the outputs are
        # the new leaves and must be fresh tensors, the intermediate of the separable resampling need not be
        g_old = lp.density_tensor.grid_dim()
        nbytes = max(ops.vm_resize_scratch_bytes(t.channel_dim(), g_old, int(new_grid_dim))
                     for t in (lp.density_tensor.stacked_single_vm, lp.appearance_tensor.stacked_single_vm))
        scratch = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.aabb.device)
        lp.density_tensor = lp.density_tensor.resize(new_grid_dim, scratch=scratch)
        lp.appearance_tensor = lp.appearance_tensor.resize(new_grid_dim, scratch=scratch)
        for mom in ("mu", "nu"):
            st = self.optimizer_state[mom]
            for which in ("density", "appearance"):
                v, m = ops.vm_resize(st[f"{which}_vector"], st[f"{which}_matrix"], int(new_grid_dim), scratch=scratch)
                st[f"{which}_vector"], st[f"{which}_matrix"] = v, m
        self._adam = None  # leaf buffers changed: rebuild the pointer tables
        return self

    def _resize_grid_peer(self, new_grid_dim: int) -> "TrainState":
        """resize_grid with sharded moments: gather the moment shards (5 times per training run), resample
        parameters and moments on every rank, re-create the symmetric buffers at the new size."""
        old = self._peer
        full = {"mu": old.gather_moments(old.mu), "nu": old.gather_moments(old.nu)}
        lp = self.learnable_params
        dens, app = lp.density_tensor.resize(new_grid_dim), lp.appearance_tensor.resize(new_grid_dim)
        flat = dict(lp.flat())
        flat["density_vector"], flat["density_matrix"] = dens.stacked_single_vm.vector, dens.stacked_single_vm.matrix
        flat["appearance_vector"], flat["appearance_matrix"] = app.stacked_single_vm.vector, app.stacked_single_vm.matrix
        flat = {k: v.clone() for k, v in flat.items()}  # the old symmetric buffer is about to be released
        for st in full.values():
            for which in ("density", "appearance"):
                st[f"{which}_vector"], st[f"{which}_matrix"] = ops.vm_resize(st[f"{which}_vector"].contiguous(),
                                                                             st[f"{which}_matrix"].contiguous(), int(new_grid_dim))
        self.learnable_params = render.LearnableParams.from_flat(flat, self.config.scene_contraction)
        self._peer = None
        step, self.step = self.step, 0
        self.enable_peer_optimizer(group=old.group, multicast=old.multicast if old.world > 1 else None)
        self.step = step
        self._peer.scatter_moments(self._peer.mu, full["mu"])
        self._peer.scatter_moments(self._peer.nu, full["nu"])
        return self

"""Mirror of tensorf/training.py:19-243 (TrainState.initialize / training_step) over the CUDA
library.  The forward + reverse (training.py:108-156) is the hot path and runs entirely in
libtensorf_b200.so (fused MSE, no autograd graph); Adam + LR decay + apply
(training.py:158-204, optax scale_by_adam b1=.9 b2=.99 eps=1e-8, masked group LRs) is a "next"
row and is plain torch elementwise math here.
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

    @staticmethod
    def initialize(config: train_config.TensorfConfig, grid_dim: int, prng_key, num_cameras: int,
                   device="cuda", world_size: int = 1) -> "TrainState":
        """training.py:35-99."""
        seed = prng_key if isinstance(prng_key, int) else 0
        gen = torch.Generator(device=device).manual_seed(seed)
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

    def loss_and_grads(self, minibatch: RenderedRays, render_prng_key, grads: Optional[Dict[str, torch.Tensor]] = None):
        """training.py:108-156: (mse, grads w.r.t. every LearnableParams leaf). The mean is over the
        GLOBAL batch (`world_size` x local rays); the caller allreduces the gradients."""
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
        noise = prng.render_noise(render_prng_key, R, N, self.config.scene_contraction)
        rays = minibatch.rays_wrt_world
        inputs = {"origins": rays.origins.contiguous(), "directions": rays.directions.contiguous(),
                  "camera_indices": rays.camera_indices.to(torch.int32).contiguous(), "aabb": self.aabb,
                  "colors": minibatch.colors.contiguous()}
        inputs.update(render._device_noise(noise, dev))
        if self.config.scene_contraction:
            base, delta = render.contracted_schedule(self.config.render_near, self.config.render_far, N)
            inputs["base_ts"], inputs["deltas"] = torch.from_numpy(base).to(dev), torch.from_numpy(delta).to(dev)
        call = render._acquire(desc, dev)
        _, loss = call.forward({k: v.contiguous() for k, v in flat.items()}, inputs)
        grads = call.backward(None, grads)
        render._release(call)
        return loss, grads

    def training_step(self, minibatch: RenderedRays) -> Tuple["TrainState", Dict[str, float]]:
        """training.py:101-205."""
        keys = prng.split(self.prng_key)
        render_key, new_key = keys[0], keys[1]
        loss, grads = self.loss_and_grads(minibatch, render_key)
        if self.world_size > 1:
            import torch.distributed as dist
            for g in grads.values():
                dist.all_reduce(g)
            dist.all_reduce(loss)
        cfg, oc = self.config, self.config.optimizer
        # LR decay with reset after upsampling (training.py:161-181)
        step = self.step
        if oc.lr_upsample_reset:
            deltas = [step - u for u in (0,) + tuple(cfg.upsamp_iters) if step - u >= 0]
            resetted = min(deltas)
        else:
            resetted = step
        decay_iters = oc.lr_decay_iters if oc.lr_decay_iters is not None else cfg.n_iters
        coeff = max(oc.lr_decay_target_ratio ** (resetted / decay_iters), oc.lr_decay_target_ratio)
        # optax.scale_by_adam(b1=0.9, b2=0.99, eps=1e-8) + masked group learning rates (:213-243)
        b1, b2, eps, t = 0.9, 0.99, 1e-8, step + 1
        flat = self.learnable_params.flat()
        gnorm_sq = torch.zeros((), device=loss.device)
        for k, p in flat.items():
            g = grads[k]
            gnorm_sq += (g * g).sum()
            mu, nu = self.optimizer_state["mu"][k], self.optimizer_state["nu"][k]
            mu.mul_(b1).add_(g, alpha=1 - b1)
            nu.mul_(b2).addcmul_(g, g, value=1 - b2)
            lr = oc.lr_init_tensor if k.startswith(("density_", "appearance_")) else oc.lr_init_mlp
            upd = (mu / (1 - b1**t)) / (torch.sqrt(nu / (1 - b2**t)) + eps)
            p.add_(upd, alpha=-lr * coeff)
        self.prng_key = new_key
        self.step = step + 1
        mse = float(loss.item())  # the blocking read the reference does at training.py:342
        log = {"train/mse": mse, "train/psnr": psnr_from_mse(mse), "train/lr_tensor": coeff * oc.lr_init_tensor,
               "train/lr_mlp": coeff * oc.lr_init_mlp, "train/grad_norm": float(torch.sqrt(gnorm_sq).item())}
        return self, log

"""Mirror of tensorf/train_config.py:6-71 (values feed the hot-path shapes)."""
from __future__ import annotations

import dataclasses
import pathlib
from typing import Literal, Optional, Tuple


@dataclasses.dataclass(frozen=True)
class OptimizerConfig:
    lr_init_tensor: float = 0.02
    lr_init_mlp: float = 1e-3
    lr_decay_iters: Optional[int] = None
    lr_decay_target_ratio: float = 0.1
    lr_upsample_reset: bool = True


@dataclasses.dataclass(frozen=True)
class TensorfConfig:
    run_dir: pathlib.Path = pathlib.Path("./runs/unused")
    dataset_path: pathlib.Path = pathlib.Path("./data/unused")
    dataset_type: Literal["blender", "nerfstudio"] = "blender"
    minibatch_size: int = 4096
    n_iters: int = 30000
    optimizer: OptimizerConfig = dataclasses.field(default_factory=OptimizerConfig)
    initial_aabb_min: Tuple[float, float, float] = (-1.0, -1.0, -1.0)
    initial_aabb_max: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    appearance_feat_dim: int = 24
    density_feat_dim: int = 8
    feature_n_freqs: int = 6
    viewdir_n_freqs: int = 6
    grid_dim_init: int = 128
    grid_dim_final: int = 300
    upsamp_iters: Tuple[int, ...] = (2000, 3000, 4000, 5500, 7000)
    scene_contraction: bool = False
    scene_scale: float = 1.0
    camera_embeddings: bool = False
    render_near: float = 0.05
    render_far: float = 200.0
    train_ray_sample_multiplier: float = 1.0


def lego_config(**kw) -> TensorfConfig:
    """train_lego.py:23-37."""
    base = dict(initial_aabb_min=(-0.6585, -1.1833, -0.4651), initial_aabb_max=(0.6636, 1.1929, 1.0512),
                appearance_feat_dim=48, density_feat_dim=16, feature_n_freqs=2, viewdir_n_freqs=2, grid_dim_init=128,
                grid_dim_final=300, upsamp_iters=(2000, 3000, 4000, 5500, 7000))
    base.update(kw)
    return TensorfConfig(**base)


def nerfstudio_config(**kw) -> TensorfConfig:
    """train_nerfstudio.py:23-44."""
    base = dict(dataset_type="nerfstudio", initial_aabb_min=(-2.0, -2.0, -2.0), initial_aabb_max=(2.0, 2.0, 2.0),
                appearance_feat_dim=48, density_feat_dim=32, feature_n_freqs=6, viewdir_n_freqs=6, grid_dim_init=128,
                grid_dim_final=300, upsamp_iters=(2_500, 5_000, 10_000), scene_contraction=True, camera_embeddings=True,
                render_near=0.05, render_far=200.0, train_ray_sample_multiplier=3.0, minibatch_size=2048)
    base.update(kw)
    return TensorfConfig(**base)

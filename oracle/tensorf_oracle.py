"""CPU oracle for the tensorf-jax per-ray hot path.  TEST INFRASTRUCTURE ONLY.

This file is a PyTorch-CPU restatement of the arithmetic of the reference
(brentyi/tensorf-jax).  It is the *checker* for the CUDA path: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import it.  Nothing under `tensorf-jax_b200/` (the product) imports it, and the product has no
CPU fallback.

PARITY STATUS: **parity unpinned against the reference itself.**  The reference ships no
tests, golden vectors or fixtures, and JAX/flax are not installable in this
image, so the reference cannot be executed here.  What the oracle *is* pinned against:
  * `scipy.ndimage.map_coordinates(order=1, mode="nearest")` — the SciPy routine that
    `jax.scipy.ndimage.map_coordinates` (jax 0.9.0.1, `uv.lock:504-531`) re-implements
    (tests/test_oracle_interp.py);
  * hand-derived micro cases (G=2 grids, rays that miss the box, contraction inside/outside
    the unit cube, median-depth corner cases) (tests/test_oracle_render.py);
  * its own fp64 instance (the arbiter) and torch.autograd for gradients;
  * the analytical reverse-mode statement of SURVEY.md Appendix A.6.
Third-party arithmetic restated here from its published algorithm: jax 0.9.0.1
(`map_coordinates`, `random.choice(replace=False, p)` = top_k(gumbel + log p), `nn.softplus` =
logaddexp(x, 0), `cumsum`), flax 0.12.4 (`nn.Dense` = x @ kernel + bias, `nn.Embed`).

Every function cites the reference file:line (relative to /root/reference) it follows.
All randomness (jitter, gumbel) enters as arrays, exactly as at the C-ABI boundary.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

RGB, DIST_MEDIAN, DIST_MEAN = 0, 1, 2  # render.py:18-23 (RenderMode)


# --------------------------------------------------------------------------------------
# tensor_vm.py
# --------------------------------------------------------------------------------------
def linear_interpolation_with_channel_axis(grid: torch.Tensor, coordinates: torch.Tensor) -> torch.Tensor:
    """tensor_vm.py:226-250 → jax.scipy.ndimage.map_coordinates(order=1, mode="nearest"),
    vmapped over the leading channel axis.

    grid (C, G) or (C, G, G); coordinates (d, *b) → (C, *b).
    Algorithm (jax/_src/scipy/ndimage.py, 0.9.0.1): per axis lower=floor(x),
    upper_weight=x-lower, lower_weight=1-upper_weight, indices clip(lower), clip(lower+1)
    to [0, size-1]; contributions in itertools.product order, weight product first, then
    multiplied with the gathered value, summed left to right."""
    d = coordinates.shape[0]
    assert grid.dim() - 1 == d
    nodes = []
    for ax in range(d):
        x = coordinates[ax]
        size = grid.shape[1 + ax]
        lower = torch.floor(x)
        upper_w = x - lower
        lower_w = 1.0 - upper_w
        idx = lower.to(torch.int64)
        nodes.append(
            [
                (torch.clamp(idx, 0, size - 1), lower_w),
                (torch.clamp(idx + 1, 0, size - 1), upper_w),
            ]
        )
    out = None
    if d == 1:
        for i0, w0 in nodes[0]:
            contrib = w0 * grid[:, i0]
            out = contrib if out is None else out + contrib
    elif d == 2:
        for i0, w0 in nodes[0]:
            for i1, w1 in nodes[1]:
                contrib = (w0 * w1) * grid[:, i0, i1]
                out = contrib if out is None else out + contrib
    else:
        raise NotImplementedError
    return out


def vm_single_interpolate(vector: torch.Tensor, matrix: torch.Tensor, ijk: torch.Tensor) -> torch.Tensor:
    """tensor_vm.py:140-167 (TensorVMSingle.interpolate). vector (C,G), matrix (C,G,G),
    ijk (3,*b) in [-1,1] → (C,*b)."""
    G = matrix.shape[-1]
    assert matrix.shape[-2] == G and vector.shape[-1] == G  # :169-174
    x = (ijk + 1.0) / 2.0  # :150
    x = x * (G - 1.0)  # :153
    vec = linear_interpolation_with_channel_axis(vector, x[0:1])  # :155-157
    mat = linear_interpolation_with_channel_axis(matrix, x[1:3])  # :158-160
    return vec * mat  # :167


def vm_interpolate(vector: torch.Tensor, matrix: torch.Tensor, ijk: torch.Tensor) -> torch.Tensor:
    """tensor_vm.py:42-89 (TensorVM.interpolate). vector (3,C,G), matrix (3,C,G,G),
    ijk (3,*b) → (3C,*b).  The "magic vmap" (:56-72) is numerically a no-op."""
    kij = ijk[[2, 0, 1]]  # :50
    jki = ijk[[1, 2, 0]]  # :51
    feats = [vm_single_interpolate(vector[p], matrix[p], c) for p, c in enumerate((ijk, kij, jki))]
    feat = torch.stack(feats, dim=0)  # (3, C, *b)   :81-82
    return feat.reshape((3 * vector.shape[1],) + tuple(ijk.shape[1:]))  # :86-88


# --------------------------------------------------------------------------------------
# networks.py
# --------------------------------------------------------------------------------------
def fourier_encode(coords: torch.Tensor, n_freqs: int) -> torch.Tensor:
    """networks.py:13-35. (*, D) → (*, D*2F): per dim [sin(2^j x) j<F, sin(2^j x + pi/2) j<F]."""
    coeffs = 2.0 ** torch.arange(n_freqs, dtype=coords.dtype)
    inputs = coords[..., None] * coeffs  # (*, D, F)
    # 0.5*jnp.pi is a Python double, weak-typed: it is rounded to the array dtype.
    half_pi = torch.tensor(0.5 * math.pi, dtype=coords.dtype)
    out = torch.sin(torch.cat([inputs, inputs + half_pi], dim=-1))
    return out.reshape(coords.shape[:-1] + (coords.shape[-1] * 2 * n_freqs,))


@dataclasses.dataclass(frozen=True)
class MlpConfig:
    """networks.py:38-43 (FeatureMlp fields)."""

    feature_squash_dim: int = 27
    units: int = 128
    feature_n_freqs: int = 6
    viewdir_n_freqs: int = 6
    num_cameras: Optional[int] = None

    def encoded_dim(self) -> int:  # networks.py:77-82
        return (
            self.feature_squash_dim + 3 + 2 * self.feature_n_freqs * self.feature_squash_dim + 2 * self.viewdir_n_freqs * 3
        )


def feature_mlp(cfg: MlpConfig, mlp: Dict[str, torch.Tensor], features, viewdirs, camera_indices, aux: Optional[dict] = None):
    """networks.py:46-121 (FeatureMlp.__call__). `mlp` holds flax leaves:
    w0 (Ca,27); w1 (enc,128), b1; w2 (128,128), b2; w3 (128,3), b3; embed (ncam,128)."""
    f = features @ mlp["w0"]  # :57-61 (no bias)
    x = torch.cat(
        [f, viewdirs, fourier_encode(f, cfg.feature_n_freqs), fourier_encode(viewdirs, cfg.viewdir_n_freqs)], dim=-1
    )  # :68-76
    assert x.shape[-1] == cfg.encoded_dim()
    z1 = x @ mlp["w1"] + mlp["b1"]
    x = torch.relu(z1)  # :86-90
    z2 = x @ mlp["w2"] + mlp["b2"]
    x = torch.relu(z2)  # :94-98
    if aux is not None:  # pre-activations: lets tests find rows that sit on a ReLU kink
        aux["z1"], aux["z2"] = z1.detach(), z2.detach()
    if cfg.num_cameras is not None:  # :103-111
        h = cfg.units // 2
        cond = mlp["embed"][camera_indices.to(torch.int64)]
        x = torch.cat([x[..., :h], cond[..., :h] * x[..., h:] + cond[..., h:]], dim=-1)
    x = x @ mlp["w3"] + mlp["b3"]  # :114-117
    return torch.sigmoid(x)  # :120


# --------------------------------------------------------------------------------------
# render.py
# --------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class RenderConfig:
    """render.py:26-36."""

    near: float
    far: float
    mode: int
    density_samples_per_ray: int
    appearance_samples_per_ray: int


def ray_segment_from_bounding_box(origins, directions, aabb, min_segment_length=1e-3):
    """render.py:399-434, batched over rays. origins/directions (R,3), aabb (2,3)."""
    offsets = aabb[:, None, :] - origins[None, :, :]  # (2,R,3)   :412
    t_int = offsets / (directions + 1e-8)[None]  # :413-415
    t_min_axis = t_int.min(dim=0).values  # :418
    t_max_axis = t_int.max(dim=0).values  # :419
    zero = torch.zeros((), dtype=origins.dtype)
    t_min = torch.maximum(zero, t_min_axis.max(dim=-1).values)  # :423
    t_max = t_max_axis.min(dim=-1).values  # :424
    t_max_clipped = torch.maximum(t_max, t_min + min_segment_length)  # :425
    valid = t_min < t_max  # :429
    msl = torch.tensor(min_segment_length, dtype=origins.dtype)
    return torch.where(valid, t_min, zero), torch.where(valid, t_max_clipped, msl)  # :431-434


def contracted_schedule(near: float, far: float, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """render.py:127-155: the constant `ts` schedule and step sizes of the contracted branch.
    close_ts is an fp32 jnp.linspace; far_ts is float64 numpy, cast to fp32 on concatenation."""
    nc = n // 2  # :127
    nf = n - nc  # :128
    # jnp.linspace(start, stop, num) in fp32 [upstream, jax/_src/numpy/array_creation.py]:
    # step_i = i * (1/(num-1)); out_i = start*(1-step_i) + stop*step_i for i < num-1, and the
    # endpoint `stop` is appended exactly.
    f32 = np.float32
    if nc > 1:
        stepv = np.arange(nc - 1, dtype=np.float32) * f32(1.0 / (nc - 1))
        close = (f32(near) * (f32(1.0) - stepv) + f32(near + 1.0) * stepv).astype(np.float32)
        close = np.concatenate([close, np.array([near + 1.0], dtype=np.float32)])
    else:
        close = np.full((nc,), near, dtype=np.float32)
    far_start = near + 1.0 + 1.0 / nc  # :135
    k = 10.0  # :136
    far_deltas = (1.0 / (1.0 - np.linspace(0.0, 1.0 - 1 / ((far - far_start) / k + 1), nf)) - 1.0) * np.linspace(
        1.0, k, nf
    )  # :137-148
    far_ts = far_start + far_deltas  # :149
    base = np.concatenate([close, far_ts.astype(np.float32)]).astype(np.float32)  # :151
    delta = np.roll(base, -1) - base  # :154
    delta[-1] = delta[-2]  # :155
    return base, delta.astype(np.float32)


def sample_points(cfg: RenderConfig, scene_contraction: bool, aabb, origins, directions, jitter):
    """render.py:122-195. Returns normalised points (3,R,N), ts (R,N), step_sizes (R,N).
    jitter: (N,) shared by all rays (bounded; :177-183,:373-379) or (R,N) (contracted; :158-161)."""
    R = origins.shape[0]
    N = cfg.density_samples_per_ray
    dt = origins.dtype
    if scene_contraction:
        base, delta = contracted_schedule(cfg.near, cfg.far, N)
        ts = torch.from_numpy(base).to(dt)[None, :].expand(R, N)
        step_sizes = torch.from_numpy(delta).to(dt)[None, :].expand(R, N)
        assert jitter.shape == (R, N)
        ts = ts + step_sizes * jitter  # :161
        points = origins[:, None, :] + ts[:, :, None] * directions[:, None, :]  # :164-167
        norm = points.abs().max(dim=-1, keepdim=True).values  # :170 (ord=inf)
        points = torch.where(norm <= 1.0, points, (2.0 - 1.0 / norm) * points / norm)  # :171
        points = points.permute(2, 0, 1)  # :173
    else:
        t0, t1 = ray_segment_from_bounding_box(origins, directions, aabb)
        step = (t1 - t0) / N  # :369
        assert jitter.shape == (N,)
        s = torch.arange(N, dtype=dt) + jitter  # :372-379
        ts = s[None, :] * step[:, None]  # :380
        ts = t0[:, None] + ts  # :381
        points = origins.T[:, :, None] + directions.T[:, :, None] * ts[None, :, :]  # :384-386
        step_sizes = step[:, None].expand(R, N)  # :186
    points = ((points - aabb[0][:, None, None]) / (aabb[1] - aabb[0])[:, None, None] - 0.5) * 2.0  # :193-195
    return points, ts, step_sizes


def compute_segment_probabilities(sigmas, step_sizes):
    """render.py:300-347 → (p_exits, p_terminates)."""
    neg = -sigmas * step_sizes  # :322
    p_exits = torch.exp(torch.cumsum(neg, dim=-1))  # :323
    p_term_given = 1.0 - torch.exp(neg)  # :327
    ones = torch.ones(neg.shape[:-1] + (1,), dtype=neg.dtype)
    p_terminates = p_term_given * torch.cat([ones, p_exits[..., :-1]], dim=-1)  # :331-341
    return p_exits, p_terminates


def gumbel_topk(g: torch.Tensor, k: int) -> torch.Tensor:
    """`lax.top_k(g, k)[1]`: the k largest, ties → lower index first (XLA CPU TopK compares a
    sign-flipped integer key and breaks ties by index [upstream]).  A stable descending sort
    has exactly that order.  g (R,N) → (R,k) int64, in top-k (descending) order."""
    order = torch.sort(g, dim=-1, descending=True, stable=True).indices
    return order[..., :k]


def softplus(z):
    """jax.nn.softplus = logaddexp(z, 0) = max(z,0) + log1p(exp(-|z|))  (render.py:207)."""
    return torch.clamp(z, min=0.0) + torch.log1p(torch.exp(-z.abs()))


def render_rays(
    cfg: RenderConfig,
    mlp_cfg: MlpConfig,
    params: Dict[str, torch.Tensor],
    scene_contraction: bool,
    aabb: torch.Tensor,
    origins: torch.Tensor,
    directions: torch.Tensor,
    camera_indices: torch.Tensor,
    jitter: torch.Tensor,
    gumbel: Optional[torch.Tensor],
    return_aux: bool = False,
    forced_indices: Optional[torch.Tensor] = None,
):
    """render.py:105-279 (render_rays) with `_rgb_from_points` (:437-549) inlined.

    params: density_vector (3,cd,G), density_matrix (3,cd,G,G), appearance_vector (3,ca,G),
    appearance_matrix (3,ca,G,G) and the MLP leaves (see `feature_mlp`).
    `forced_indices` replaces the Gumbel top-k result (used to compare downstream values
    when a near-tie makes the kernel and the oracle select different sets)."""
    R = origins.shape[0]
    N, K = cfg.density_samples_per_ray, cfg.appearance_samples_per_ray
    points, ts, step_sizes = sample_points(cfg, scene_contraction, aabb, origins, directions, jitter)
    density_feat = vm_interpolate(params["density_vector"], params["density_matrix"], points)  # :199
    z = density_feat.sum(dim=0) + 10.0  # :207
    sigmas = softplus(z)
    p_exits, p_terminates = compute_segment_probabilities(sigmas, step_sizes)  # :211
    aux = dict(points=points, ts=ts, step_sizes=step_sizes, z=z, sigmas=sigmas, p_exits=p_exits, p_terminates=p_terminates)

    if cfg.mode == RGB:
        if forced_indices is not None:
            idx = forced_indices.to(torch.int64)
        else:
            g = gumbel[None, :].to(p_terminates.dtype) + torch.log(p_terminates)  # :461-469
            idx = gumbel_topk(g.detach(), K)
            aux["g"] = g.detach()
        rows = torch.arange(R)[:, None]
        visible_points = points[:, rows, idx]  # (3,R,K)   :472
        app_feat = vm_interpolate(params["appearance_vector"], params["appearance_matrix"], visible_points)  # :476
        Ca = app_feat.shape[0]
        app_feat = app_feat.permute(1, 2, 0).reshape(R * K, Ca)  # :484-486
        viewdirs = directions[:, None, :].expand(R, K, 3).reshape(-1, 3)  # :487-490
        cams = camera_indices[:, None].expand(R, K).reshape(-1)  # :494-496
        visible_rgb = feature_mlp(mlp_cfg, params, app_feat, viewdirs, cams, aux=aux).reshape(R, K, 3)  # :499-509
        rgb = torch.zeros(R, N, 3, dtype=visible_rgb.dtype)
        rgb = rgb.index_put((rows, idx), visible_rgb)  # :511-515
        sampled_pt = p_terminates[rows, idx]  # :529-531
        unbias = (1.0 - p_exits[:, -1] + 1e-8) / (sampled_pt.sum(dim=1) + 1e-8)  # :537-546
        expected = (rgb * p_terminates[:, :, None]).sum(dim=-2) * unbias[:, None]  # :233-236
        out = expected + p_exits[:, -1:] * torch.ones(3, dtype=expected.dtype)  # :241-244
        aux.update(indices=idx, visible_rgb=visible_rgb, unbias=unbias, app_feat=app_feat)
    elif cfg.mode == DIST_MEDIAN:
        inf = torch.full((R, 1), float("inf"), dtype=ts.dtype)
        dist = torch.cat([ts, inf], dim=-1)  # :250-252
        pna = torch.cat([1.0 - p_exits, torch.ones(R, 1, dtype=ts.dtype)], dim=-1)  # :253-255
        mask = pna > 0.5  # :258
        mm = torch.zeros_like(mask)
        mm[..., 1:] = torch.logical_xor(mask[..., :-1], mask[..., 1:])  # :259-263
        # :266 is `sum(mask * dist)`; under jit XLA turns convert(pred)*x into select(pred,x,0)
        # [upstream], so the observed semantics are where(mask, dist, 0) (no 0*inf NaN).
        out = torch.where(mm, dist, torch.zeros_like(dist)).sum(dim=-1)
    elif cfg.mode == DIST_MEAN:
        dist = torch.cat([ts, ts[:, -1:]], dim=-1)  # :271
        ptp = torch.cat([p_terminates, p_exits[:, -1:]], dim=-1)  # :272-274
        out = (ptp * dist).sum(dim=-1)  # :276
    else:
        raise AssertionError
    return (out, aux) if return_aux else out


def training_sample_counts(grid_dim: int, multiplier: float = 1.0) -> Tuple[int, int]:
    """training.py:115-118."""
    n = int(math.sqrt(3 * grid_dim**2) * multiplier)
    return n, int(0.15 * n)


def training_loss(cfg, mlp_cfg, params, scene_contraction, aabb, origins, directions, camera_indices, colors, jitter, gumbel,
                  forced_indices=None):
    """training.py:108-141 (compute_loss): mse = mean((rendered - colors)^2) over (R,3)."""
    rendered = render_rays(cfg, mlp_cfg, params, scene_contraction, aabb, origins, directions, camera_indices, jitter, gumbel,
                           forced_indices=forced_indices)
    return torch.mean((rendered - colors) ** 2), rendered


def loss_and_grads(cfg, mlp_cfg, params, scene_contraction, aabb, origins, directions, camera_indices, colors, jitter, gumbel,
                   forced_indices=None):
    """training.py:153-156 (jax.value_and_grad over LearnableParams) via torch.autograd."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    loss, rendered = training_loss(cfg, mlp_cfg, leaves, scene_contraction, aabb, origins, directions, camera_indices, colors,
                                   jitter, gumbel, forced_indices)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return loss.detach(), rendered.detach(), grads


# --------------------------------------------------------------------------------------
# training.py:158-243 — optimiser (optax 0.2.6, restated from its published algorithm)
# --------------------------------------------------------------------------------------
def lr_decay_coeff(step: int, upsamp_iters, n_iters: int, lr_decay_iters, target_ratio: float, upsample_reset: bool) -> float:
    """training.py:158-181.  optax.exponential_decay(init_value=1, transition_steps=T, decay_rate=r,
    end_value=r)(count) = clip(r ** (count / T), min=r) (decay_rate < 1 → end_value is a lower bound)."""
    if upsample_reset:
        deltas = [step - u for u in (0,) + tuple(upsamp_iters)]
        resetted = min(d for d in deltas if d >= 0)  # training.py:161-168
    else:
        resetted = step
    T = lr_decay_iters if lr_decay_iters is not None else n_iters  # :171-173
    return max(target_ratio ** (resetted / T), target_ratio)


def global_norm(grads) -> float:
    """optax.global_norm (training.py:194): sqrt(sum over leaves of sum(g^2)); fp64 arbiter."""
    return math.sqrt(sum(float((np.asarray(g, dtype=np.float64) ** 2).sum()) for g in grads))


def adam_step(params, grads, mu, nu, count: int, neg_lrs, lr_decay: float, b1=0.9, b2=0.99, eps=1e-8, eps_root=0.0,
              dtype=np.float32):
    """training.py:183-201 with the optimiser of :213-243, per leaf, every operation rounded in `dtype`
    in optax's order:
      scale_by_adam:  mu' = (1-b1)*g + b1*mu ; nu' = (1-b2)*g**2 + b2*nu          (update_moment[_per_elem_norm])
                      mu_hat = mu' / (1 - b1**t) ; nu_hat = nu' / (1 - b2**t), t = count+1   (bias_correction)
                      u = mu_hat / (sqrt(nu_hat + eps_root) + eps)
      masked scale:   u = neg_lr * u      (neg_lr = -lr_init_mlp | -lr_init_tensor, :225-243)
      decay:          u = lr_decay * u    (:186)
      apply_updates:  p' = p + u          (:199-201)
    Returns (new_params, new_mu, new_nu) as lists of numpy arrays."""
    f = dtype
    t = f(count + 1)
    bc1 = f(1) - np.power(f(b1), t)
    bc2 = f(1) - np.power(f(b2), t)
    omb1, omb2 = f(1) - f(b1), f(1) - f(b2)
    out_p, out_m, out_v = [], [], []
    for p, g, m, v, nlr in zip(params, grads, mu, nu, neg_lrs):
        p, g, m, v = (np.asarray(a, dtype=f) for a in (p, g, m, v))
        m2 = omb1 * g + f(b1) * m
        v2 = omb2 * (g * g) + f(b2) * v
        u = (m2 / bc1) / (np.sqrt(v2 / bc2 + f(eps_root)) + f(eps))
        out_p.append(p + f(lr_decay) * (f(nlr) * u))
        out_m.append(m2)
        out_v.append(v2)
    return out_p, out_m, out_v


# --------------------------------------------------------------------------------------
# tensor_vm.py:183-223 — resize (jax.image.scale_and_translate, jax 0.9.0.1, restated)
# --------------------------------------------------------------------------------------
def resize_weight_matrix(input_size: int, output_size: int, dtype=np.float32) -> np.ndarray:
    """jax/_src/image/scale.py compute_weight_mat for the triangle ("linear") kernel with antialias=True and
    the align-corners scale/translation of tensor_vm.py:204-213:
      scale = (out-1)/(in-1) ; translation = -(scale/2 - 0.5)
      inv_scale = 1/scale ; kernel_scale = max(inv_scale, 1)
      sample_f = (arange(out)+0.5)*inv_scale - translation*inv_scale - 0.5
      x = |sample_f[None,:] - arange(in)[:,None]| / kernel_scale ; w = max(0, 1-x)
      w /= sum_in(w) where |sum| > 1000*eps32 (else 0) ; w = 0 where sample_f outside [-0.5, in-0.5]
    Returns (in, out)."""
    f = dtype
    scale = f(f(output_size - 1.0) / f(input_size - 1.0))
    translation = f(-(scale / f(2.0) - f(0.5)))
    inv_scale = f(f(1.0) / scale)
    kernel_scale = max(inv_scale, f(1.0))
    sample_f = (np.arange(output_size, dtype=f) + f(0.5)) * inv_scale - translation * inv_scale - f(0.5)
    x = np.abs(sample_f[None, :] - np.arange(input_size, dtype=f)[:, None]) / kernel_scale
    w = np.maximum(f(0), f(1) - np.abs(x)).astype(f)
    total = w.sum(axis=0, keepdims=True, dtype=f)
    w = np.where(np.abs(total) > 1000.0 * float(np.finfo(np.float32).eps), w / np.where(total != 0, total, 1), 0).astype(f)
    inside = np.logical_and(sample_f >= -0.5, sample_f <= input_size - 0.5)
    return np.where(inside[None, :], w, 0).astype(f)


def vm_resize(vector: np.ndarray, matrix: np.ndarray, grid_dim: int, dtype=np.float32):
    """TensorVM.resize (tensor_vm.py:91-100) → TensorVMSingle.resize (:183-199) →
    resize_with_aligned_corners (:202-223): vector (3,C,G) → (3,C,g'), matrix (3,C,G,G) → (3,C,g',g').
    Spatial dims are those whose size changes (:208-210); none → identity."""
    G = vector.shape[-1]
    v, m = np.asarray(vector, dtype=dtype), np.asarray(matrix, dtype=dtype)
    if G == grid_dim:
        return v.copy(), m.copy()
    w = resize_weight_matrix(G, grid_dim, dtype)
    vo = np.einsum("pci,io->pco", v, w).astype(dtype)
    mo = np.einsum("pcij,ia,jb->pcab", m, w, w, optimize=True).astype(dtype)
    return vo, mo


# --------------------------------------------------------------------------------------
# cameras.py:100-143 — per-pixel rays
# --------------------------------------------------------------------------------------
def pixel_rays(K: np.ndarray, T_camera_world: np.ndarray, width: int, height: int, camera_index: int, dtype=np.float32):
    """Camera.pixel_rays_wrt_world (cameras.py:124-143) over ray_wrt_world_from_uv (:100-121):
    direction = R_world_camera @ K^-1 @ [u, v, 1] (left to right), / (norm + 1e-8); origin =
    T_world_camera.translation() = -R^T t; u = column, v = row (onp.mgrid[:H, :W], :134).
    T_camera_world is a 4x4 homogeneous matrix. Returns origins (H,W,3), directions (H,W,3), camera_indices (H,W)."""
    f = dtype
    T = np.asarray(T_camera_world, dtype=f)
    R_wc = T[:3, :3].T
    origin = (-(R_wc @ T[:3, 3])).astype(f)
    M = (R_wc @ np.linalg.inv(np.asarray(K, dtype=f))).astype(f)
    v, u = np.mgrid[:height, :width]
    uv1 = np.stack([u, v, np.ones_like(u)], axis=-1).astype(f)
    d = np.einsum("ij,hwj->hwi", M, uv1).astype(f)
    d = (d / (np.linalg.norm(d, axis=-1, keepdims=True).astype(f) + f(1e-8))).astype(f)
    return np.broadcast_to(origin, d.shape).copy(), d, np.full((height, width), camera_index, dtype=np.uint32)

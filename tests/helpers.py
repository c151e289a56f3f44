"""Shared helpers for the parity tests: the same NumPy arrays go to the CUDA path and to the
CPU oracle."""
import numpy as np
import torch

import tensorf_oracle as O
from tensorf_b200 import synthetic as S

# Tolerances (SURVEY.md §8c / BASELINE.json north_star: 1e-4 relative in fp32).
RTOL_OUT = 1e-4      # |x - ref| <= RTOL_OUT * max(1, |ref|)  for rgb / depth / loss
RTOL_GRAD = 1e-4     # ||g - ref||_inf <= RTOL_GRAD * ||ref||_inf and relative L2 <= RTOL_GRAD


def T(x, dtype=torch.float32, device="cpu"):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if t.dtype.is_floating_point:
        t = t.to(dtype)
    elif t.dtype == torch.uint32:
        t = t.view(torch.int32)
    return t.to(device)


def oracle_cfgs(w, mode=O.RGB):
    cfg = O.RenderConfig(w.near, w.far, mode, w.N, w.K)
    mc = O.MlpConfig(27, 128, w.feat_freqs, w.view_freqs, w.num_cameras)
    return cfg, mc


def oracle_inputs(inp, dtype=torch.float32):
    P = {k: T(v, dtype) for k, v in inp["params"].items()}
    return dict(
        params=P, aabb=T(inp["aabb"], dtype), origins=T(inp["origins"], dtype), directions=T(inp["directions"], dtype),
        camera_indices=torch.from_numpy(inp["camera_indices"].astype(np.int64)), colors=T(inp["colors"], dtype),
        jitter=T(inp["jitter"], dtype), gumbel=T(inp["gumbel"], dtype),
    )


def device_inputs(w, inp, device, with_colors=True):
    params = {k: T(v, device=device) for k, v in inp["params"].items()}
    d = dict(origins=T(inp["origins"], device=device), directions=T(inp["directions"], device=device),
             camera_indices=T(inp["camera_indices"], device=device), aabb=T(inp["aabb"], device=device),
             jitter=T(inp["jitter"], device=device), gumbel=T(inp["gumbel"], device=device))
    if w.contracted:
        base, delta = O.contracted_schedule(w.near, w.far, w.N)
        d["base_ts"] = T(base, device=device)
        d["deltas"] = T(delta, device=device)
    if with_colors:
        d["colors"] = T(inp["colors"], device=device)
    return params, d


def assert_close_out(x, ref, rtol=RTOL_OUT, what=""):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert x.shape == ref.shape, (what, x.shape, ref.shape)
    both_inf = np.isinf(x) & np.isinf(ref) & (np.sign(x) == np.sign(ref))
    err = np.where(both_inf, 0.0, np.abs(x - ref))
    bound = rtol * np.maximum(1.0, np.abs(np.where(both_inf, 0.0, ref)))
    bad = ~(err <= bound)
    assert not bad.any(), f"{what}: {bad.sum()} / {bad.size} out of tolerance, max err {np.nanmax(err):.3e}"


def grad_errors(g, ref):
    g = np.asarray(g, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    ninf = np.abs(ref).max()
    l2 = np.sqrt((ref**2).sum())
    if ninf == 0:
        return float(np.abs(g).max()), float(np.sqrt((g**2).sum()))
    return float(np.abs(g - ref).max() / ninf), float(np.sqrt(((g - ref) ** 2).sum()) / l2)


def assert_close_grad(g, ref, rtol=RTOL_GRAD, what=""):
    einf, el2 = grad_errors(g, ref)
    assert einf <= rtol and el2 <= rtol, f"{what}: rel-inf {einf:.3e}, rel-L2 {el2:.3e} > {rtol}"


KINK_TAU = 5e-6


def kink_rows(aux, tau=KINK_TAU):
    """Rows of the MLP batch with a hidden pre-activation within rounding of zero. d relu/dz is
    discontinuous there: two correct fp32 implementations may legitimately pick different
    sides, so such rows get a zero cotangent in BOTH the kernel and the oracle (the same
    treatment the selection stage gets for near-ties)."""
    bad = (aux["z1"].abs() < tau).any(dim=-1) | (aux["z2"].abs() < tau).any(dim=-1)
    return torch.where(bad)[0].numpy()


def audit_median_mismatches(depth, ref, aux64, N):
    """DIST_MEDIAN picks the first sample whose 1 - p_exit exceeds 0.5 (render.py:250-266): a discontinuous output.
    Every ray on which the kernel and the fp32 oracle disagree must be a rounding case of that threshold: in the
    fp64 oracle (`aux64`), all samples between the two answers have |1 - E - 0.5| within fp32 rounding of the
    cumulative sum (4 N eps).  Returns the number of (audited) mismatching rays."""
    same = np.isclose(depth, ref, rtol=1e-4, atol=1e-6) | (np.isinf(depth) & np.isinf(ref))
    ts = aux64["ts"].numpy()
    pna = 1.0 - aux64["p_exits"].numpy()
    tol = 4 * N * np.finfo(np.float32).eps
    for r in np.where(~same)[0]:
        def index_of(v):  # sample whose distance the value is (N = "never crossed": inf)
            return N if np.isinf(v) else int(np.argmin(np.abs(ts[r] - v)))
        a, b = sorted((index_of(depth[r]), index_of(ref[r])))
        assert a != b, f"ray {r}: depth {depth[r]} vs {ref[r]} but the same sample"
        if not np.isinf(depth[r]):
            assert abs(ts[r][index_of(depth[r])] - depth[r]) <= 1e-4 * max(1.0, abs(depth[r])), f"ray {r}: {depth[r]} is not a sample distance"
        between = pna[r, a:b]
        assert (np.abs(between - 0.5) <= tol).all(), f"ray {r}: samples {a}..{b} are not at the 0.5 threshold: {between}"
    return int((~same).sum())

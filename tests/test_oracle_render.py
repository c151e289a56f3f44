"""Hand-derived micro cases and internal consistency of the render oracle (render.py)."""
import math

import numpy as np
import torch

import tensorf_oracle as O
from helpers import T, grad_errors, oracle_cfgs, oracle_inputs
from tensorf_b200 import synthetic as S


def test_ray_segment_hit_and_miss():
    aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]], dtype=torch.float64)
    o = torch.tensor([[-3.0, 0.0, 0.0], [-3.0, 5.0, 0.0], [0.0, 0.0, 0.0]], dtype=torch.float64)
    d = torch.tensor([[1.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    t0, t1 = O.ray_segment_from_bounding_box(o, d, aabb)
    assert abs(t0[0] - 2.0) < 1e-6 and abs(t1[0] - 4.0) < 1e-6          # through the box
    assert t0[1] == 0.0 and t1[1] == 1e-3                                # misses: (0, min_segment_length)
    assert t0[2] == 0.0 and abs(t1[2] - 1.0) < 1e-6                      # origin inside: t_min clipped to 0


def test_contraction_inside_outside():
    cfg = O.RenderConfig(0.05, 200.0, O.RGB, 8, 2)
    aabb = torch.tensor([[-2.0] * 3, [2.0] * 3], dtype=torch.float64)
    o = torch.zeros(1, 3, dtype=torch.float64)
    d = torch.tensor([[1.0, 0.0, 0.0]], dtype=torch.float64)
    pts, ts, steps = O.sample_points(cfg, True, aabb, o, d, torch.zeros(1, 8, dtype=torch.float64))
    base, delta = O.contracted_schedule(0.05, 200.0, 8)
    world = ts[0].numpy()
    expect = np.where(world <= 1.0, world, (2 - 1 / world))           # L-inf contraction along x
    np.testing.assert_allclose(pts[0, 0].numpy(), expect / 2.0, rtol=1e-12)   # aabb +-2 -> /2
    assert (np.abs(pts.numpy()) <= 1.0).all()
    np.testing.assert_allclose(steps[0].numpy(), delta.astype(np.float64))
    assert base[0] == np.float32(0.05) and abs(base[3] - 1.05) < 1e-6 and delta[-1] == delta[-2]
    assert abs(float(base[-1]) - 200.0) / 200.0 < 1e-3               # far end reaches config.far


def test_segment_probabilities_sum_to_one():
    sig = torch.rand(5, 40, dtype=torch.float64) * 3
    steps = torch.full((5, 40), 0.1, dtype=torch.float64)
    pe, pt = O.compute_segment_probabilities(sig, steps)
    np.testing.assert_allclose((pt.sum(-1) + pe[:, -1]).numpy(), 1.0, rtol=1e-12)
    np.testing.assert_allclose(pe[:, 0].numpy(), np.exp(-sig[:, 0].numpy() * 0.1))


def test_median_depth_corner_cases():
    w = S.Workload("m", 3, 4, 1, 1, 6, 1, 1, 1)
    inp = S.make_inputs(w)
    cfg = O.RenderConfig(0.05, 200.0, O.DIST_MEDIAN, 6, 1)
    mc = O.MlpConfig(27, 128, 1, 1, None)
    oi = oracle_inputs(inp, torch.float64)
    aabb = torch.tensor([[-1.0] * 3, [1.0] * 3], dtype=torch.float64)
    o = torch.tensor([[-3.0, 0.0, 0.0]] * 3, dtype=torch.float64)
    d = torch.tensor([[1.0, 0.0, 0.0]] * 3, dtype=torch.float64)
    P = dict(oi["params"])
    jit = torch.zeros(6, dtype=torch.float64)
    # (a) opaque at the first sample: softplus(10)*step=10/3 -> 1-E_0 > 0.5 -> median 0 (no transition)
    P["density_vector"] = torch.zeros_like(P["density_vector"])
    out = O.render_rays(cfg, mc, P, False, aabb, o, d, oi["camera_indices"][:3], jit, None)
    assert out.tolist() == [0.0, 0.0, 0.0]
    # (b) transparent: sigma = softplus(-40) ~ 0 -> never crosses -> inf
    P["density_vector"] = torch.full_like(P["density_vector"], 1.0)
    P["density_matrix"] = torch.full_like(P["density_matrix"], -50.0 / 3)
    out = O.render_rays(cfg, mc, P, False, aabb, o, d, oi["camera_indices"][:3], jit, None)
    assert torch.isinf(out).all()
    # (c) sigma*step = softplus(z)*(1/3): choose z so the crossing happens at the 3rd sample
    z = math.log(math.expm1(0.9))            # softplus(z) = 0.9 -> a = -0.3/sample; 1-exp(-0.3 k) > 0.5 at k=3
    P["density_matrix"] = torch.full_like(P["density_matrix"], (z - 10.0) / 3)
    out, aux = O.render_rays(cfg, mc, P, False, aabb, o, d, oi["camera_indices"][:3], jit, None, return_aux=True)
    np.testing.assert_allclose(out.numpy(), aux["ts"][:, 2].numpy())


def test_gumbel_topk_ties_and_c_oracle():
    import ctypes
    import pathlib
    g = torch.tensor([[1.0, 3.0, 3.0, -float("inf"), 2.0, 3.0, -float("inf")]])
    assert O.gumbel_topk(g, 4).tolist() == [[1, 2, 5, 4]]
    assert O.gumbel_topk(g, 7).tolist() == [[1, 2, 5, 4, 0, 3, 6]]
    so = pathlib.Path(__file__).resolve().parents[1] / "oracle" / "_build" / "liboracle_topk.so"
    if so.exists():
        lib = ctypes.CDLL(str(so))
        rng = np.random.default_rng(0)
        gg = np.round(rng.normal(size=(30, 97)) * 3).astype(np.float32)
        gg[:, 50:60] = -np.inf
        out = np.empty((30, 20), np.int32)
        lib.oracle_topk_select(gg.ctypes.data_as(ctypes.c_void_p), 30, 97, 20, out.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(out, O.gumbel_topk(torch.from_numpy(gg), 20).numpy())


def test_reverse_mode_statement_matches_autograd():
    """SURVEY.md Appendix A.6 (what the CUDA reverse kernels implement) vs torch.autograd."""
    torch.manual_seed(0)
    N, K = 30, 6
    z = (torch.randn(N, dtype=torch.float64) * 2).requires_grad_(True)
    delta = torch.rand(N, dtype=torch.float64) * 0.2 + 0.01
    c = torch.rand(K, 3, dtype=torch.float64, requires_grad=True)
    idx = torch.sort(torch.randperm(N)[:K]).values
    col = torch.rand(3, dtype=torch.float64)
    sigma = O.softplus(z)
    pe, pt = O.compute_segment_probabilities(sigma[None], delta[None])
    pe, pt = pe[0], pt[0]
    S_ = pt[idx].sum()
    A, B = 1 - pe[-1] + 1e-8, S_ + 1e-8
    W = (c * pt[idx, None]).sum(0)
    rgb = W * (A / B) + pe[-1]
    loss = ((rgb - col) ** 2).sum() / 3
    loss.backward()
    with torch.no_grad():
        go = 2 * (rgb - col) / 3
        u = A / B
        dc = go[None, :] * u * pt[idx, None]
        gpt = torch.zeros(N, dtype=torch.float64)
        gpt[idx] = (go[None, :] * (u * c - W[None, :] * A / B**2)).sum(-1)
        gE = (go * (1 - W / B)).sum()
        q = gpt * pt
        suffix = torch.flip(torch.cumsum(torch.flip(q, [0]), 0), [0]) - q
        da = -gpt * pe + suffix + gE * pe[-1]
        dz = da * (-delta) * torch.exp(z - sigma)
    np.testing.assert_allclose(dc.numpy(), c.grad.numpy(), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(dz.numpy(), z.grad.numpy(), rtol=1e-9, atol=1e-14)


def test_fp32_oracle_within_tolerance_of_fp64():
    w = S.Workload("small", 64, 9, 2, 3, 37, 5, 2, 2)
    inp = S.make_inputs(w, bias_std=0.05)
    cfg, mc = oracle_cfgs(w)
    res = {}
    for dt in (torch.float32, torch.float64):
        oi = oracle_inputs(inp, dt)
        res[dt] = O.loss_and_grads(cfg, mc, oi["params"], False, oi["aabb"], oi["origins"], oi["directions"],
                                   oi["camera_indices"], oi["colors"], oi["jitter"], oi["gumbel"])
    assert abs(float(res[torch.float32][0]) - float(res[torch.float64][0])) < 1e-5
    np.testing.assert_allclose(res[torch.float32][1].numpy(), res[torch.float64][1].numpy(), atol=1e-4)
    for k in res[torch.float64][2]:
        einf, el2 = grad_errors(res[torch.float32][2][k].numpy(), res[torch.float64][2][k].numpy())
        assert einf < 1e-4 and el2 < 1e-4, (k, einf, el2)


def test_sample_counts():
    assert S.sample_counts(128) == (221, 33) and S.sample_counts(300) == (519, 77)
    assert S.sample_counts(128, 3.0) == (665, 99) and S.sample_counts(300, 3.0) == (1558, 233)
    assert O.training_sample_counts(162) == (280, 42)

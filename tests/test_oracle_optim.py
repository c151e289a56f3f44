"""Pins of the optimiser / resize oracle against independent implementations
available in this image: torch.optim.Adam and torch.nn.functional.interpolate(align_corners=True)."""
import numpy as np
import torch

from oracle import tensorf_oracle as O


def test_adam_oracle_matches_torch_adam():
    rng = np.random.default_rng(0)
    shapes = [(7, 5), (33,), (3, 4, 9)]
    p0 = [rng.normal(size=s).astype(np.float32) for s in shapes]
    lrs = [1e-3, 0.02, 0.02]
    tp = [torch.tensor(p, dtype=torch.float64, requires_grad=True) for p in p0]
    opt = torch.optim.Adam([{"params": [t], "lr": lr} for t, lr in zip(tp, lrs)], betas=(0.9, 0.99), eps=1e-8)
    p = [x.copy() for x in p0]
    mu = [np.zeros_like(x) for x in p0]
    nu = [np.zeros_like(x) for x in p0]
    for step in range(6):
        g = [rng.normal(size=s).astype(np.float32) * 10.0 ** rng.integers(-4, 2) for s in shapes]
        for t, gi in zip(tp, g):
            t.grad = torch.tensor(gi, dtype=torch.float64)
        opt.step()
        p, mu, nu = O.adam_step(p, g, mu, nu, count=step, neg_lrs=[-x for x in lrs], lr_decay=1.0)
        for a, t in zip(p, tp):
            # torch divides sqrt(nu) by sqrt(bc2) instead of nu by bc2: same value up to rounding / eps placement
            np.testing.assert_allclose(a, t.detach().numpy(), rtol=2e-5, atol=1e-7)


def test_adam_oracle_fp32_vs_fp64_and_decay():
    rng = np.random.default_rng(1)
    p = [rng.normal(size=(50,)).astype(np.float32)]
    g = [rng.normal(size=(50,)).astype(np.float32)]
    z = [np.zeros(50, np.float32)]
    a32 = O.adam_step(p, g, z, z, 0, [-0.02], 0.5)
    a64 = O.adam_step(p, g, z, z, 0, [-0.02], 0.5, dtype=np.float64)
    for x, y in zip(a32, a64):
        np.testing.assert_allclose(x[0], y[0], rtol=1e-5, atol=1e-9)
    # first step of Adam moves every coordinate by ~lr * lr_decay * sign(g)
    np.testing.assert_allclose(a64[0][0] - p[0], -0.02 * 0.5 * np.sign(g[0]), rtol=1e-4)


def test_global_norm():
    g = [np.full((3, 4), 2.0, np.float32), np.full((4,), 3.0, np.float32)]
    assert abs(O.global_norm(g) - np.sqrt(12 * 4 + 4 * 9)) < 1e-12


def test_lr_decay_coeff_schedule():
    ups = (2000, 3000, 4000, 5500, 7000)
    kw = dict(upsamp_iters=ups, n_iters=30000, lr_decay_iters=None, target_ratio=0.1)
    assert O.lr_decay_coeff(0, upsample_reset=True, **kw) == 1.0
    assert abs(O.lr_decay_coeff(1999, upsample_reset=True, **kw) - 0.1 ** (1999 / 30000)) < 1e-15
    assert O.lr_decay_coeff(2000, upsample_reset=True, **kw) == 1.0          # reset at an upsampling iteration
    assert abs(O.lr_decay_coeff(7100, upsample_reset=True, **kw) - 0.1 ** (100 / 30000)) < 1e-15
    assert abs(O.lr_decay_coeff(7100, upsample_reset=False, **kw) - 0.1 ** (7100 / 30000)) < 1e-15
    assert O.lr_decay_coeff(10 ** 6, upsample_reset=False, **kw) == 0.1      # end_value floor


def test_resize_oracle_matches_torch_align_corners_when_upsampling():
    rng = np.random.default_rng(2)
    for gi, go in ((5, 9), (16, 23), (128, 162)):
        v = rng.normal(size=(3, 2, gi)).astype(np.float32)
        m = rng.normal(size=(3, 2, gi, gi)).astype(np.float32)
        vo, mo = O.vm_resize(v, m, go)
        tv = torch.nn.functional.interpolate(torch.from_numpy(v), size=go, mode="linear", align_corners=True).numpy()
        tm = torch.nn.functional.interpolate(torch.from_numpy(m), size=(go, go), mode="bilinear", align_corners=True).numpy()
        # fp32 sample positions carry ~G*eps of rounding (the reference computes them in fp32 too) ...
        np.testing.assert_allclose(vo, tv, rtol=0, atol=3e-4)
        np.testing.assert_allclose(mo, tm, rtol=0, atol=3e-4)
        # ... the fp64 instance agrees with torch's fp64 interpolation to rounding
        vo64, mo64 = O.vm_resize(v, m, go, dtype=np.float64)
        tm64 = torch.nn.functional.interpolate(torch.from_numpy(m).double(), size=(go, go), mode="bilinear", align_corners=True)
        np.testing.assert_allclose(mo64, tm64.numpy(), rtol=0, atol=1e-11)
        np.testing.assert_allclose(mo, mo64, rtol=0, atol=3e-4)
        # corners are kept exactly (align_corners)
        np.testing.assert_allclose(mo[..., 0, 0], m[..., 0, 0], rtol=0, atol=1e-6)
        np.testing.assert_allclose(mo[..., -1, -1], m[..., -1, -1], rtol=0, atol=1e-5)


def test_resize_weights_properties():
    for gi, go in ((9, 5), (300, 128), (128, 300), (7, 7 * 3)):
        w = O.resize_weight_matrix(gi, go)
        assert w.shape == (gi, go)
        np.testing.assert_allclose(w.sum(axis=0), 1.0, rtol=0, atol=1e-6)        # normalised columns
        assert (w >= 0).all()
        support = (w > 0).sum(axis=0)
        if go >= gi:
            assert support.max() <= 2                                             # plain linear interpolation
        else:
            assert support.max() > 2                                              # antialiased: widened triangle
        w64 = O.resize_weight_matrix(gi, go, dtype=np.float64)
        np.testing.assert_allclose(w, w64, rtol=0, atol=5e-5)
    # constants are preserved, identity when the size does not change
    v, m = np.full((3, 2, 6), 1.5, np.float32), np.full((3, 2, 6, 6), -2.0, np.float32)
    vo, mo = O.vm_resize(v, m, 11)
    np.testing.assert_allclose(vo, 1.5, atol=1e-6)
    np.testing.assert_allclose(mo, -2.0, atol=1e-6)
    vo, mo = O.vm_resize(v, m, 6)
    assert np.array_equal(vo, v) and np.array_equal(mo, m)

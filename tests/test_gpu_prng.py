"""Device-side jax.random restatement and pixel-ray generation against the host
restatement (tensorf_b200/prng.py, Random123 KAT-checked) and the oracle's cameras.py:100-143."""
import math

import numpy as np
import pytest
import torch

from oracle import tensorf_oracle as O
from tensorf_b200 import cameras, ops, prng, synthetic as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


@pytest.mark.parametrize("shape", [(1,), (33,), (665,), (7, 41), (2048, 665), (0,)])
def test_uniform_bits_exact(cuda, shape):
    key = prng.split(prng.Key.from_seed(11))[0]
    got = ops.prng_uniform(key.k0, key.k1, shape, cuda).cpu().numpy()
    want = prng.uniform(key, shape)
    assert got.shape == want.shape and np.array_equal(got, want)             # integer cipher + exact fp32 ops
    if got.size:
        assert got.min() >= 0.0 and got.max() < 1.0


def test_uniform_range_arguments(cuda):
    key = prng.Key(123, 456)
    got = ops.prng_uniform(key.k0, key.k1, (1000,), cuda, minval=-2.0, maxval=3.0).cpu().numpy()
    assert np.array_equal(got, prng.uniform(key, (1000,), -2.0, 3.0))


def test_gumbel_matches_host(cuda):
    key = prng.split(prng.Key.from_seed(5))[1]
    got = ops.prng_gumbel(key.k0, key.k1, (4096,), cuda).cpu().numpy()
    want = prng.gumbel(key, (4096,))
    # -log(-log u): logf vs numpy log differ by an ulp; the outer log amplifies that by 1/|log u'| at most ~1e1
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)
    assert np.isfinite(got).all()


def test_render_noise_device_matches_host(cuda):
    key = prng.Key.from_seed(9)
    for contracted in (False, True):
        host = prng.render_noise(key, 16, 37, contracted)
        dev = prng.render_noise_device(key, 16, 37, contracted, cuda)
        assert np.array_equal(dev["jitter"].cpu().numpy(), host.jitter)
        np.testing.assert_allclose(dev["gumbel"].cpu().numpy(), host.gumbel, rtol=2e-6, atol=2e-6)
    assert prng.render_noise_device(key, 4, 8, False, cuda, need_gumbel=False)["gumbel"] is None


def _look_at(origin):
    z = -origin / np.linalg.norm(origin)
    x = np.cross(z, np.array([0.0, 0.0, 1.0])); x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R_wc = np.stack([x, y, z], axis=1)           # camera axes as columns
    T = np.eye(4)
    T[:3, :3] = R_wc.T
    T[:3, 3] = -R_wc.T @ origin
    return T.astype(np.float32)


@pytest.mark.parametrize("W,H", [(40, 30), (800, 800), (1, 1)])
def test_pixel_rays_match_oracle(cuda, W, H):
    T_cw = _look_at(np.array([2.5, -1.5, 2.0]))
    cam = cameras.Camera.from_fov(T_cw, W, H, fov_x_radians=0.6911)
    rays = cam.pixel_rays_wrt_world(camera_index=7, device=cuda)
    o, d, c = O.pixel_rays(cam.K, T_cw, W, H, 7)
    o64, d64, _ = O.pixel_rays(cam.K, T_cw, W, H, 7, dtype=np.float64)
    assert rays.get_batch_axes() == (H, W)
    np.testing.assert_allclose(rays.origins.cpu().numpy(), o, rtol=0, atol=1e-6)
    np.testing.assert_allclose(rays.directions.cpu().numpy(), d, rtol=0, atol=2e-6)
    np.testing.assert_allclose(rays.directions.cpu().numpy(), d64, rtol=0, atol=1e-4)   # fp32 K^-1 and 800-px coordinates
    assert (rays.camera_indices.cpu().numpy().view(np.uint32) == c).all()
    n = torch.linalg.norm(rays.directions, dim=-1)
    assert float((n - 1).abs().max()) < 1e-6
    # a row band equals the same rows of the full frame (image tiles for multi-GPU rendering)
    if H >= 4:
        band = cam.pixel_rays_wrt_world(7, device=cuda, rows=(H // 4, H // 2))
        assert torch.equal(band.directions, rays.directions[H // 4:H // 2])
    assert abs(cam.compute_fov_x_radians() - 0.6911) < 1e-6
    small = cam.resize_with_fixed_fov(max(W // 2, 1), max(H // 2, 1))
    assert abs(small.compute_fov_x_radians() - 0.6911) < 1e-5


def test_uniform_slice_is_the_rows_of_the_whole_draw(cuda):
    """A rank of a sharded contracted-scene batch draws ITS rows of the global (R,N) jitter (render.py:158-161):
    `tensorf_prng_uniform_slice` must give exactly the corresponding elements of the single-device draw."""
    from tensorf_b200 import ops, prng
    key = prng.Key.from_seed(11)
    R, N = 37, 53
    whole = ops.prng_uniform(key.k0, key.k1, (R, N), cuda)
    for a, b in ((0, 9), (9, 30), (30, 37)):
        part = ops.prng_uniform(key.k0, key.k1, (b - a, N), cuda, first=a * N)
        assert torch.equal(part, whole[a:b])
    d0 = prng.render_noise_device(key, 10, N, True, cuda, need_gumbel=False, first_ray=0)["jitter"]
    d1 = prng.render_noise_device(key, 10, N, True, cuda, need_gumbel=False, first_ray=10)["jitter"]
    both = prng.render_noise_device(key, 20, N, True, cuda, need_gumbel=False)["jitter"]
    assert torch.equal(torch.cat([d0, d1]), both)

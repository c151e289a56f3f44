"""Device-resident training data path: ray table built on the device, minibatches by gather."""
import numpy as np
import pytest
import torch

from oracle import tensorf_oracle as O
from tensorf_b200 import cameras, data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _views(n, W, H, rng):
    views = []
    for i in range(n):
        a = 0.7 * i
        o = 3.0 * np.array([np.cos(a), np.sin(a), 0.6])
        z = -o / np.linalg.norm(o)
        x = np.cross(z, [0.0, 0.0, 1.0]); x /= np.linalg.norm(x)
        R_wc = np.stack([x, np.cross(z, x), z], axis=1)
        T = np.eye(4); T[:3, :3] = R_wc.T; T[:3, 3] = -R_wc.T @ o
        cam = cameras.Camera.from_fov(T.astype(np.float32), W, H, fov_x_radians=0.69)
        views.append(data.RegisteredRgbaView(torch.from_numpy(rng.uniform(size=(H, W, 4)).astype(np.float32)), cam))
    return views


def test_ray_table_and_minibatches(cuda):
    rng = np.random.default_rng(0)
    W, H, nv = 12, 9, 3
    views = _views(nv, W, H, rng)
    table = data.rendered_rays_from_views(views, device=cuda)
    assert table.get_batch_axes() == (nv * W * H,)
    # table == oracle restatement of data.py:301-337 / cameras.py:100-143
    for i, v in enumerate(views):
        o, d, c = O.pixel_rays(v.camera.K, v.camera.T_camera_world, W, H, i)
        sl = slice(i * W * H, (i + 1) * W * H)
        np.testing.assert_allclose(table.rays_wrt_world.directions[sl].cpu().numpy(), d.reshape(-1, 3), atol=2e-6, rtol=0)
        np.testing.assert_allclose(table.rays_wrt_world.origins[sl].cpu().numpy(), o.reshape(-1, 3), atol=1e-6, rtol=0)
        assert (table.rays_wrt_world.camera_indices[sl].cpu().numpy() == i).all()
        rgba = v.image_rgba.numpy()
        want = rgba[..., :3] * rgba[..., 3:4] + (np.float32(1.0) - rgba[..., 3:4])
        assert np.array_equal(table.colors[sl].cpu().numpy(), want.reshape(-1, 3))        # exact: same fp32 ops
    loader = data.DeviceRayLoader(table, minibatch_size=50)
    assert loader.minibatch_count() == (nv * W * H) // 50
    it = loader.cycled(shuffle_seed=0)
    seen = []
    for _ in range(loader.minibatch_count()):
        mb = next(it)
        assert mb.get_batch_axes() == (50,) and tuple(mb.colors.shape) == (50, 3)
        seen.append(mb)
    perm = np.random.default_rng(0).permutation(nv * W * H)
    got_o = torch.cat([m.rays_wrt_world.origins for m in seen]).cpu().numpy()
    got_c = torch.cat([m.colors for m in seen]).cpu().numpy()
    used = perm[:loader.minibatch_count() * 50]
    assert np.array_equal(got_o, table.rays_wrt_world.origins.cpu().numpy()[used])          # gather is exact
    assert np.array_equal(got_c, table.colors.cpu().numpy()[used])
    assert len(set(used.tolist())) == used.size                                             # an epoch never repeats a ray
    nxt = next(it)                                                                          # next epoch reshuffles
    perm1 = np.random.default_rng(1).permutation(nv * W * H)
    assert np.array_equal(nxt.colors.cpu().numpy(), table.colors.cpu().numpy()[perm1[:50]])
    assert loader.bad_index_count() == 0


def test_gather_edge_cases(cuda):
    rng = np.random.default_rng(1)
    table = data.rendered_rays_from_views(_views(1, 5, 4, rng), device=cuda)
    loader = data.DeviceRayLoader(table, minibatch_size=20)
    empty = loader.gather(torch.zeros(0, dtype=torch.int64, device=cuda))
    assert empty.get_batch_axes() == (0,)
    mb = loader.gather(torch.tensor([0, 19, 19, 3, 25, -1], device=cuda))                   # repeats and out-of-table rows
    assert loader.bad_index_count() == 2
    assert torch.equal(mb.colors[1], table.colors[19]) and torch.equal(mb.colors[4], table.colors[0])
    with pytest.raises(ValueError):
        data.DeviceRayLoader(table, minibatch_size=21)


def test_host_stage_single_copy(cuda):
    """data.HostStage: all staged arrays reach the device with one copy; re-upload after a host-side edit."""
    from tensorf_b200.data import HostStage
    rng = np.random.default_rng(12)
    arrays = {"origins": rng.normal(size=(33, 3)).astype(np.float32), "camera_indices": rng.integers(0, 2**32, 33, dtype=np.uint32),
              "jitter": rng.uniform(size=(33, 5)).astype(np.float32)}
    st = HostStage(arrays, cuda)
    assert st._host.is_pinned() and st.nbytes == (100 + 36 + 168) * 4
    st.upload()
    torch.cuda.synchronize()
    for k, a in arrays.items():
        got = st.device[k].cpu().numpy()
        assert np.array_equal(got.view(a.dtype) if a.dtype == np.uint32 else got, a), k
        assert st.device[k].data_ptr() % 16 == 0
    st.host["jitter"].zero_()
    st.upload()
    torch.cuda.synchronize()
    assert float(st.device["jitter"].abs().sum()) == 0.0 and np.array_equal(st.device["origins"].cpu().numpy(), arrays["origins"])

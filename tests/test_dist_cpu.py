"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group (the N>1 path of bench.py /
TrainState.training_step) and the sharding arithmetic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tensorf_b200 import dist as tdist
from tensorf_b200 import synthetic as S


def test_shard_ranges_cover_and_balance():
    for total in (0, 1, 7, 4096, 16385):
        for world in (1, 2, 3, 8):
            spans = [tdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        tdist.shard_range(10, 2, 2)
    assert tdist.tile_rows(800, 3, 8) == (300, 400)


def test_stripe_rows_cover_every_row_once():
    for H in (1, 16, 21, 800):
        for world in (1, 2, 3, 8):
            seen = np.zeros(H, int)
            for r in range(world):
                for a, b in tdist.stripe_rows(H, r, world):
                    assert 0 <= a < b <= H and b - a <= 16
                    seen[a:b] += 1
            assert (seen == 1).all()
    # 800 rows over 8 ranks: 50 stripes, 6 or 7 per rank -> at most one stripe of imbalance
    sizes = [sum(b - a for a, b in tdist.stripe_rows(800, r, 8)) for r in range(8)]
    assert max(sizes) - min(sizes) <= 16 and sum(sizes) == 800
    # the balanced stripe height deals every rank the same number of rows whenever a power-of-two stripe can
    assert [tdist.balanced_stripe(800, w) for w in (1, 2, 4, 8)] == [16, 16, 8, 4]
    for w in (2, 4, 8):
        s = tdist.balanced_stripe(800, w)
        assert len({sum(b - a for a, b in tdist.stripe_rows(800, r, w, s)) for r in range(w)}) == 1
    assert tdist.balanced_stripe(21, 2) == 16  # nothing divides: keep the default


def test_shard_rays_slices_only_per_ray_arrays():
    w = S.dozer_workload(R=10, G=8)
    inp = S.make_inputs(w)
    flat = {k: inp[k] for k in ("origins", "directions", "camera_indices", "colors", "jitter", "gumbel", "aabb")}
    parts = [tdist.shard_rays(flat, r, 3, per_ray=("origins", "directions", "camera_indices", "colors", "jitter")) for r in range(3)]
    assert [p["origins"].shape[0] for p in parts] == [4, 3, 3]
    assert np.array_equal(np.concatenate([p["jitter"] for p in parts]), inp["jitter"])   # contracted: (R,N) rows
    assert all(p["gumbel"] is inp["gumbel"] and p["aabb"] is inp["aabb"] for p in parts)  # shared by all ranks


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shapes = {"density_vector": (3, 2, 5), "w1": (7, 4), "b3": (3,)}
        fg = tdist.FlatGrads(shapes, "cpu")
        assert fg.total == 30 + 28 + 3 and fg.leaves["w1"].data_ptr() == fg.flat[30:].data_ptr()
        # rank-local "gradients": a deterministic function of the rank's ray shard
        a, b = tdist.shard_range(10, rank, world)
        for i, (k, t) in enumerate(fg.leaves.items()):
            t.copy_(torch.full(t.shape, float(sum(range(a, b)) + i)))
        fg.allreduce()
        expect = {k: float(sum(range(10)) + world * i) for i, k in enumerate(shapes)}
        ok = all(torch.all(fg.leaves[k] == expect[k]).item() for k in shapes)
        scale = tdist.global_loss_scale(b - a, world)
        # two buckets + loss slot (the overlapped exchange of bench.py / TrainState.training_step)
        fb = tdist.FlatGrads(shapes, "cpu", loss_slot=True)
        assert fb.buffer.numel() == fb.total + 1 and fb.late.numel() == 30 and fb.early.numel() == 28 + 3 + 1
        assert fb.late.data_ptr() == fb.leaves["density_vector"].data_ptr() and fb.early.data_ptr() == fb.leaves["w1"].data_ptr()
        for i, (k, t) in enumerate(fb.leaves.items()):
            t.copy_(torch.full(t.shape, float(rank + 1 + i)))
        fb.loss.fill_(0.25 * (rank + 1))
        fb.start_allreduce("early")
        late_before = fb.late.clone()
        fb.start_allreduce("late")
        fb.finish()
        tri = world * (world + 1) / 2
        ok = ok and all(torch.all(fb.leaves[k] == tri + world * i).item() for i, k in enumerate(shapes))
        ok = ok and float(fb.loss.item()) == 0.25 * tri and torch.all(late_before == rank + 1).item() and not fb._pending
        single = tdist.FlatGrads({"w1": (2,), "density_vector": (3,)}, "cpu")   # density not first: one bucket
        ok = ok and single.late.numel() == 0 and single.early.numel() == 5
        out[rank] = (ok, scale)
    finally:
        dist.destroy_process_group()


def test_flat_grads_allreduce_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0][0] and out[1][0]
    assert out[0][1] == pytest.approx(1.0 / 30.0)   # 1/(3*R_global) with R_global = 10


def test_host_stage_layout_cpu():
    """data.HostStage: one buffer, 16-byte aligned views, dtypes and values preserved (layout logic only; the pinned
    H2D path runs in tests/test_gpu_data.py)."""
    from tensorf_b200.data import HostStage
    arrays = {"origins": np.arange(15, dtype=np.float32).reshape(5, 3), "camera_indices": np.array([7, 2**32 - 1, 3], np.uint32),
              "gumbel": np.linspace(0, 1, 7, dtype=np.float32)}
    st = HostStage(arrays, "cpu")
    assert st.nbytes == (16 + 4 + 8) * 4
    st.host["gumbel"][0] = 5.0
    st.upload()
    assert np.array_equal(st.device["origins"].numpy(), arrays["origins"])
    assert st.device["camera_indices"].dtype == torch.int32
    assert np.array_equal(st.device["camera_indices"].numpy().view(np.uint32), arrays["camera_indices"])
    assert float(st.device["gumbel"][0]) == 5.0 and st.device["gumbel"].shape == (7,)
    assert all(v.data_ptr() % 16 == 0 for v in st.device.values())
    with pytest.raises(TypeError):
        HostStage({"x": np.zeros(3, np.float64)}, "cpu")

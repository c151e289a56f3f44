"""`tensorf_adam_step_peer` (SURVEY §8e fused follow-up: gradient reduce-scatter + Adam + parameter all-gather in
one kernel over peer memory) against `tensorf_adam_step` and the oracle restatement of training.py:158-243.

The transport is the only thing a single GPU cannot exercise, so the kernel's multi-rank logic (shard ranges,
rank-ordered sums, float4s that straddle leaves, stores into every rank's buffer, norm slots) is checked here with
`world` buffers living on ONE device; the two-process test (NVLink P2P and NVSwitch multicast) runs when the box has
two GPUs (tests/peer_worker.py under torchrun)."""
import ctypes as C
import os
import pathlib
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import tensorf_oracle as O
from helpers import T
from test_gpu_optim import _ulp_diff

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parents[1]

SHAPES = [
    {"density_vector": (3, 4, 9), "density_matrix": (3, 4, 9, 9), "w0": (144, 27), "w1": (150, 128), "b1": (128,),
     "w3": (128, 3), "b3": (3,)},                                                  # ragged, unaligned leaf boundaries
    {"a": (4096,), "b": (4097,), "c": (1,), "d": (0,), "e": (3, 16, 128), "f": (8192 + 5,)},   # tile boundaries, empty leaf
    {"tiny": (2,)},                                                                # shards smaller than the world
]


def _neg_lrs(shapes):
    return {k: -(0.02 if i % 2 else 1e-3) for i, k in enumerate(shapes)}


def _grads(rng, shapes, step):
    g = {k: (rng.normal(size=s) * 10.0 ** rng.integers(-6, 3)).astype(np.float32) for k, s in shapes.items()}
    if step == 2:
        first = next(iter(shapes))
        g[first][...] = 0.0
    return g


@pytest.mark.parametrize("shapes", SHAPES)
def test_world1_bit_identical_to_adam_step(cuda, shapes):
    from tensorf_b200 import dist as tdist, ops
    rng = np.random.default_rng(5)
    neg = _neg_lrs(shapes)
    p0 = {k: rng.normal(size=s).astype(np.float32) for k, s in shapes.items()}
    peer = tdist.PeerAdam(shapes, neg, cuda)
    assert peer.world == 1 and peer.shard == (0, peer.total) and peer.total % 4 == 0 and not peer.multicast
    peer.load_params({k: T(v, device=cuda) for k, v in p0.items()})
    names = list(shapes)
    rp = [T(p0[k], device=cuda) for k in names]
    rm, rv = [torch.zeros_like(t) for t in rp], [torch.zeros_like(t) for t in rp]
    ref = ops.AdamCall(rp, rm, rv, [neg[k] for k in names])
    op, om, ov = [p0[k] for k in names], [np.zeros(shapes[k], np.float32) for k in names], [np.zeros(shapes[k], np.float32) for k in names]
    for step in range(4):
        g = _grads(rng, shapes, step)
        decay = 0.1 ** (step / 7.0)
        for k in names:
            peer.grads[k].copy_(T(g[k], device=cuda))
        gn = peer.step(count=step, lr_decay=decay)
        gn_ref = ref.step([T(g[k], device=cuda) for k in names], count=step, lr_decay=decay)
        op, om, ov = O.adam_step(op, [g[k] for k in names], om, ov, step, [neg[k] for k in names], decay)
        want = O.global_norm([g[k] for k in names])
        assert abs(float(gn.item()) - want) <= 2e-6 * max(1.0, want)
        assert abs(float(gn.item()) - float(gn_ref.item())) <= 2e-6 * max(1.0, want)
        mu_full, nu_full = peer.gather_moments(peer.mu), peer.gather_moments(peer.nu)
        for i, k in enumerate(names):
            assert torch.equal(peer.params[k], rp[i]), (step, k)       # same adam_one, same bits
            assert torch.equal(mu_full[k], rm[i]) and torch.equal(nu_full[k], rv[i]), (step, k)
            assert _ulp_diff(peer.params[k].cpu().numpy(), op[i]) <= 2, (step, k)
    assert float(peer.buf[peer.leaf_total:peer.total].abs().sum()) == 0.0      # padding untouched


def _simulated_step(lib, world, total, offs, neg, bufs, mus, nus, scratch, count, decay, b1=0.9, b2=0.99):
    """Run the kernel once per simulated rank; all `world` flat buffers live on one device."""
    from tensorf_b200 import _lib
    t = np.float32(count + 1)
    bc1 = np.float32(1) - np.power(np.float32(b1), t)
    bc2 = np.float32(1) - np.power(np.float32(b2), t)
    n = len(neg)
    c_offs = (C.c_int64 * (n + 1))(*[int(x) for x in offs])
    c_neg = (C.c_float * n)(*neg)
    gp = (C.c_void_p * world)(*[b.data_ptr() + 4 * total for b in bufs])
    pp = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
    sp = (C.c_void_p * world)(*[b.data_ptr() + 8 * total + 64 for b in bufs])
    for r in range(world):
        sb, se = C.c_int64(), C.c_int64()
        lib.tensorf_peer_shard(total, r, world, C.byref(sb), C.byref(se))
        d = _lib.PeerAdamDesc(adam=_lib.AdamDesc(n_leaves=n, reserved=0, b1=b1, b2=b2, eps=1e-8, eps_root=0.0,
                                                 bias_correction1=float(bc1), bias_correction2=float(bc2),
                                                 lr_decay=float(decay), reserved2=0.0),
                              rank=r, world=world, total=total, shard_begin=sb.value, shard_end=se.value)
        _lib.check(lib.tensorf_adam_step_peer(None, C.byref(d), c_offs, c_neg, gp, pp, None, None, mus[r].data_ptr(),
                                              nus[r].data_ptr(), sp, scratch.data_ptr(), scratch.numel()))


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("shapes", SHAPES[:2])
def test_simulated_ranks_on_one_device(cuda, world, shapes):
    """`world` rank buffers on one GPU: after every rank's launch all parameter copies are identical and equal to
    Adam on the rank-ordered fp32 sum of the gradients; the norm slots hold each shard's share."""
    from tensorf_b200 import _lib, ops
    lib = _lib.load()
    rng = np.random.default_rng(6)
    names = list(shapes)
    sizes = [int(np.prod(shapes[k])) for k in names]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    leaf_total = int(offs[-1])
    total = (leaf_total + 3) // 4 * 4
    neg = [_neg_lrs(shapes)[k] for k in names]
    p0 = rng.normal(size=leaf_total).astype(np.float32)
    bufs = [torch.zeros(2 * total + 32, device=cuda) for _ in range(world)]
    for b in bufs:
        b[:leaf_total] = T(p0, device=cuda)
    shard_sizes = []
    for r in range(world):
        sb, se = C.c_int64(), C.c_int64()
        lib.tensorf_peer_shard(total, r, world, C.byref(sb), C.byref(se))
        shard_sizes.append((sb.value, se.value))
    mus = [torch.zeros(max(e - b, 4), device=cuda) for b, e in shard_sizes]
    nus = [torch.zeros(max(e - b, 4), device=cuda) for b, e in shard_sizes]
    scratch = torch.empty(int(lib.tensorf_peer_adam_scratch_bytes(total)), dtype=torch.uint8, device=cuda)
    ref_p = T(p0, device=cuda)
    ref_m, ref_v = torch.zeros_like(ref_p), torch.zeros_like(ref_p)
    views = lambda t: [t[int(offs[i]):int(offs[i + 1])] for i in range(len(names))]
    ref = ops.AdamCall(views(ref_p), views(ref_m), views(ref_v), neg)
    for step in range(3):
        gs = [np.concatenate([x.ravel() for x in _grads(rng, shapes, step + r).values()]).astype(np.float32) for r in range(world)]
        for r in range(world):
            bufs[r][total:total + leaf_total] = T(gs[r], device=cuda)
        gsum = gs[0].copy()
        for r in range(1, world):
            gsum = (gsum + gs[r]).astype(np.float32)        # rank order, fp32, every rounding kept
        decay = 0.5 ** step
        _simulated_step(lib, world, total, offs, neg, bufs, mus, nus, scratch, step, decay)
        gsum_t = T(gsum, device=cuda)
        ref.step(views(gsum_t), count=step, lr_decay=decay)
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(bufs[r][:leaf_total], ref_p), (step, r)
            assert torch.equal(bufs[r][2 * total + 16:2 * total + 16 + world], bufs[0][2 * total + 16:2 * total + 16 + world])
        slots = bufs[0][2 * total + 16:2 * total + 16 + world].double().cpu().numpy()
        for r, (b, e) in enumerate(shard_sizes):
            want = float(np.sum(gsum[b:min(e, leaf_total)].astype(np.float64) ** 2))
            assert abs(slots[r] - want) <= 2e-6 * max(1e-30, want), (step, r)
        gn = torch.zeros((), device=cuda)
        _lib.check(lib.tensorf_peer_grad_norm(None, bufs[1].data_ptr() + 8 * total + 64, world, gn.data_ptr()))
        want = float(np.sqrt(np.sum(gsum.astype(np.float64) ** 2)))
        assert abs(float(gn.item()) - want) <= 2e-6 * max(1.0, want)
        mu_cat = torch.cat([m[:e - b] for m, (b, e) in zip(mus, shard_sizes)])[:leaf_total]
        nu_cat = torch.cat([v[:e - b] for v, (b, e) in zip(nus, shard_sizes)])[:leaf_total]
        assert torch.equal(mu_cat, ref_m) and torch.equal(nu_cat, ref_v)


@pytest.mark.parametrize("max_ctas", [0, 3], ids=["full_grid", "3_ctas"])
@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_simulated_allreduce_on_one_device(cuda, world, max_ctas):
    """tensorf_peer_allreduce with `world` buffers on one GPU: after every rank's launch all buffers hold the
    rank-ordered fp32 sum.  `max_ctas`: the grid cap an exchange overlapped with a compute kernel uses
    (tensorf_peer_set_max_ctas) must not change the result."""
    from tensorf_b200 import _lib
    lib = _lib.load()
    _lib.check(lib.tensorf_peer_set_max_ctas(max_ctas))
    assert lib.tensorf_peer_set_max_ctas(-1) == -1
    rng = np.random.default_rng(8)
    for total in (4, 8 * 4096 + 20, 1_000_004):
        xs = [(rng.normal(size=total) * 10.0 ** rng.integers(-3, 3)).astype(np.float32) for _ in range(world)]
        bufs = [T(x, device=cuda) for x in xs]
        want = xs[0].copy()
        for r in range(1, world):
            want = (want + xs[r]).astype(np.float32)
        ptrs = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
        for r in range(world):
            _lib.check(lib.tensorf_peer_allreduce(None, r, world, total, ptrs, None))
        torch.cuda.synchronize()
        for r in range(world):
            assert np.array_equal(bufs[r].cpu().numpy(), want), (total, r)
    _lib.check(lib.tensorf_peer_set_max_ctas(0))
    assert lib.tensorf_peer_allreduce(None, 0, world, 6, ptrs, None) == -1
    assert lib.tensorf_peer_allreduce(None, world, world, 8, ptrs, None) == -1


def test_peer_rejects_bad_arguments(cuda):
    from tensorf_b200 import _lib
    lib = _lib.load()
    buf = torch.zeros(2 * 8 + 16, device=cuda)
    mu, nu = torch.zeros(8, device=cuda), torch.zeros(8, device=cuda)
    scratch = torch.empty(64, dtype=torch.uint8, device=cuda)
    offs = (C.c_int64 * 2)(0, 7)
    neg = (C.c_float * 1)(-1.0)
    ptrs = lambda o: (C.c_void_p * 1)(buf.data_ptr() + o)

    def call(**kw):
        f = dict(rank=0, world=1, total=8, shard_begin=0, shard_end=8, bc1=0.1)
        f.update(kw)
        d = _lib.PeerAdamDesc(adam=_lib.AdamDesc(n_leaves=1, reserved=0, b1=0.9, b2=0.99, eps=1e-8, eps_root=0.0,
                                                 bias_correction1=f["bc1"], bias_correction2=0.01, lr_decay=1.0, reserved2=0.0),
                              rank=f["rank"], world=f["world"], total=f["total"], shard_begin=f["shard_begin"],
                              shard_end=f["shard_end"])
        return lib.tensorf_adam_step_peer(None, C.byref(d), offs, neg, ptrs(32), ptrs(0), f.get("g_mc"), None, mu.data_ptr(),
                                          nu.data_ptr(), ptrs(64), scratch.data_ptr(), f.get("scratch_bytes", 64))
    assert call() == 0
    assert call(total=6) == -1 and b"multiple of 4" in lib.tensorf_last_error()
    assert call(shard_begin=2) == -1
    assert call(shard_end=12) == -1
    assert call(rank=1) == -1
    assert call(world=17) == -1
    assert call(bc1=0.0) == -1
    assert call(g_mc=buf.data_ptr()) == -1 and b"both" in lib.tensorf_last_error()
    assert call(scratch_bytes=8) == -1
    assert call(total=16, shard_end=16) == -1 and b"padding" in lib.tensorf_last_error()   # leaves cover 7 of 16
    torch.cuda.synchronize()


def test_train_state_peer_optimizer_world1(cuda):
    """TrainState.training_step with the fused exchange + optimiser (world 1) follows the plain path: same loss
    and gradient norm per step (the scatter's atomics reorder sums, so not bit-for-bit), and resize_grid keeps
    parameters and moments."""
    from tensorf_b200 import cameras, synthetic as S, train_config, training
    cfg = train_config.lego_config(grid_dim_init=16, grid_dim_final=24, upsamp_iters=(2,), n_iters=10, minibatch_size=128,
                                   appearance_feat_dim=8, density_feat_dim=4)
    o, d, c = S.lego_rays(128, seed=5)
    mb = training.RenderedRays(colors=T(S.make_colors(128), device=cuda),
                               rays_wrt_world=cameras.Rays3D(T(o, device=cuda), T(d, device=cuda), T(c, device=cuda)))
    a = training.TrainState.initialize(cfg, grid_dim=16, prng_key=0, num_cameras=10, device=cuda)
    b = training.TrainState.initialize(cfg, grid_dim=16, prng_key=0, num_cameras=10, device=cuda).enable_peer_optimizer()
    assert b._peer is not None and b.learnable_params.flat()["w1"].data_ptr() == b._peer.params["w1"].data_ptr()
    for step in range(4):
        if step == 2:
            a, b = a.resize_grid(24), b.resize_grid(24)
            assert b.learnable_params.density_tensor.grid_dim() == 24 and b._peer.shapes["density_matrix"][-1] == 24
            mu_a = a.optimizer_state["mu"]["density_matrix"]
            mu_b = b._peer.gather_moments(b._peer.mu)["density_matrix"]
            np.testing.assert_allclose(mu_b.cpu().numpy(), mu_a.cpu().numpy(), rtol=0,
                                       atol=2e-3 * float(mu_a.abs().max()) + 1e-12)
        _, la = a.training_step(mb)
        _, lb = b.training_step(mb)
        assert abs(la["train/mse"] - lb["train/mse"]) <= 1e-4 * max(1.0, la["train/mse"]), (step, la, lb)
        assert abs(la["train/grad_norm"] - lb["train/grad_norm"]) <= 1e-3 * max(1e-6, la["train/grad_norm"]), (step, la, lb)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NVLink peer mapping)")
def test_two_ranks_over_nvlink(tmp_path):
    """Two processes, one per GPU: symmetric-memory rendezvous, P2P and (when the fabric offers it) multicast
    transport, against NCCL all-reduce + tensorf_adam_step."""
    out = tmp_path / "peer.json"
    env = dict(os.environ, PEER_WORKER_OUT=str(out), PEER_WORKER_QUICK="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29741", str(ROOT / "tests" / "peer_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=420)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert out.exists()


def test_backward_phases_match_whole_pass(cuda):
    """tensorf_render_rgb_bwd_phase: phase 1 (appearance half) then phase 2 (density half) give the gradients of the
    whole pass (atomics reorder sums: 1e-5 of the leaf's largest entry); phase 1 leaves the density leaves alone and
    phase 2 touches nothing else - what the overlapped exchange relies on."""
    from helpers import device_inputs
    from tensorf_b200 import ops, synthetic as S
    w = S.Workload("phases", 256, 32, 16, 48, 55, 8, 2, 2)
    inp = S.make_inputs(w, bias_std=0.05)
    desc = ops.make_desc(R=w.R, N=w.N, K=w.K, G=w.G, cd=w.cd, ca=w.ca, feat_freqs=2, view_freqs=2, loss_scale=1.0 / (3 * w.R))
    call = ops.RenderCall(desc, cuda)
    params, dins = device_inputs(w, inp, cuda)
    slot = torch.zeros(1, device=cuda)
    _, loss = call.forward(params, dins, loss_out=slot)
    assert loss.data_ptr() == slot.data_ptr() and float(slot.item()) > 0.0
    whole = {k: v.clone() for k, v in call.backward(None).items()}
    g = {k: torch.full_like(v, float("nan")) for k, v in whole.items()}
    call.forward(params, dins, loss_out=slot)      # fresh residuals (the reverse pass may reuse activation buffers)
    call.backward(None, g, phase=1)
    torch.cuda.synchronize()
    dens = ("density_vector", "density_matrix")
    assert all(torch.isnan(g[k]).all() for k in dens)
    assert all(not torch.isnan(g[k]).any() for k in g if k not in dens)
    after1 = {k: v.clone() for k, v in g.items()}
    call.backward(None, g, phase=2)
    torch.cuda.synchronize()
    for k in g:
        if k not in dens:
            assert torch.equal(g[k], after1[k]), k
        ref = whole[k]
        tol = 1e-5 * float(ref.abs().max()) + 1e-12
        assert float((g[k] - ref).abs().max()) <= tol, (k, float((g[k] - ref).abs().max()), tol)
    with pytest.raises(Exception):
        call.backward(None, g, phase=3)


@pytest.mark.skipif(os.environ.get("CUDA_LAUNCH_BLOCKING") == "1",
                    reason="the simulated ranks must run concurrently (they meet inside their kernels)")
@pytest.mark.parametrize("device_epoch", [False, True], ids=["epoch_arg", "epoch_on_device"])
@pytest.mark.parametrize("world", [2, 4])
def test_inkernel_sync_ranks_on_streams(cuda, world, device_epoch):
    """tensorf_peer_allreduce_sync: the two cross-rank barriers inside the kernel (signal pads, epochs, grid gate).
    `world` simulated ranks launch on separate streams of one GPU and have to meet inside their kernels; three calls
    in a row exercise the epoch logic and the arrival-counter reset.  `device_epoch`: epoch argument 0, the kernel keeps
    the call number in local_flags[2] (the CUDA-graph capturable form)."""
    from tensorf_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(9)
    total = 8 * 4096 + 20                                    # a few blocks per rank: all co-resident
    sig = [torch.zeros(32, dtype=torch.int32, device=cuda) for _ in range(world)]
    local = [torch.zeros(4, dtype=torch.int32, device=cuda) for _ in range(world)]
    streams = [torch.cuda.Stream(device=cuda) for _ in range(world)]
    sig_ptrs = (C.c_void_p * world)(*[s.data_ptr() for s in sig])
    for epoch in (1, 2, 3):
        xs = [rng.normal(size=total).astype(np.float32) for _ in range(world)]
        bufs = [T(x, device=cuda) for x in xs]
        want = xs[0].copy()
        for r in range(1, world):
            want = (want + xs[r]).astype(np.float32)
        ptrs = (C.c_void_p * world)(*[b.data_ptr() for b in bufs])
        torch.cuda.synchronize()
        for r in range(world):
            _lib.check(lib.tensorf_peer_allreduce_sync(C.c_void_p(streams[r].cuda_stream), r, world, total, ptrs, None, sig_ptrs,
                                                       local[r].data_ptr(), 0 if device_epoch else epoch))
        torch.cuda.synchronize()
        for r in range(world):
            assert np.array_equal(bufs[r].cpu().numpy(), want), (epoch, r)
            assert sig[r][:world].tolist() == [epoch] * world and sig[r][16:16 + world].tolist() == [epoch] * world
            assert local[r][:3].tolist() == [epoch, 0, epoch if device_epoch else 0]
    assert lib.tensorf_peer_allreduce_sync(None, 0, world, total, ptrs, None, None, None, 4) == -1

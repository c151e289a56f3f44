"""CUDA optimiser step (`tensorf_adam_step`) and grid resampling (`tensorf_vm_resize`) against the oracle
restatement of training.py:158-276 / tensor_vm.py:183-223."""
import numpy as np
import pytest
import torch

from oracle import tensorf_oracle as O
from helpers import T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _ulp_diff(a, b):
    a, b = np.asarray(a, np.float32).ravel(), np.asarray(b, np.float32).ravel()
    ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, np.int64(-(2 ** 31)) - ia, ia)
    ib = np.where(ib < 0, np.int64(-(2 ** 31)) - ib, ib)
    return int(np.abs(ia - ib).max()) if a.size else 0


@pytest.mark.parametrize("shapes", [
    [(3, 4, 9), (3, 4, 9, 9), (144, 27), (150, 128), (128,), (128, 3), (3,)],      # ragged, unaligned tails
    [(4096,), (4097,), (1,), (0,), (3, 16, 128), (8192 + 5,)],                      # chunk boundaries, empty leaf
])
def test_adam_step_matches_oracle(cuda, shapes):
    from tensorf_b200 import ops
    rng = np.random.default_rng(3)
    p = [rng.normal(size=s).astype(np.float32) for s in shapes]
    mu = [np.zeros(s, np.float32) for s in shapes]
    nu = [np.zeros(s, np.float32) for s in shapes]
    neg_lrs = [-(0.02 if i % 2 else 1e-3) for i in range(len(shapes))]
    dp, dm, dv = ([T(x, device=cuda) for x in xs] for xs in (p, mu, nu))
    call = ops.AdamCall(dp, dm, dv, neg_lrs)
    for step in range(4):
        g = [(rng.normal(size=s) * 10.0 ** rng.integers(-6, 3)).astype(np.float32) for s in shapes]
        if step == 2:
            g[0][...] = 0.0                       # all-zero gradient leaf: update = 0/(0+eps) path
        decay = 0.1 ** (step / 7.0)
        gnorm = call.step([T(x, device=cuda) for x in g], count=step, lr_decay=decay)
        p, mu, nu = O.adam_step(p, g, mu, nu, step, neg_lrs, decay)
        assert abs(float(gnorm.item()) - O.global_norm(g)) <= 2e-6 * max(1.0, O.global_norm(g))
        for i in range(len(shapes)):
            for got, want, what in ((dp[i], p[i], "param"), (dm[i], mu[i], "mu"), (dv[i], nu[i], "nu")):
                got = got.cpu().numpy()
                # same operation order with every rounding kept: at most an ulp or two (sqrt/div are IEEE on both sides)
                assert _ulp_diff(got, want) <= 2, (step, i, what, _ulp_diff(got, want))


def test_adam_step_rejects_bad_arguments(cuda):
    from tensorf_b200 import ops
    p = [torch.zeros(8, device=cuda)]
    with pytest.raises(ValueError):
        ops.AdamCall(p, [torch.zeros(9, device=cuda)], [torch.zeros(8, device=cuda)], [-1.0])
    call = ops.AdamCall(p, [torch.zeros(8, device=cuda)], [torch.zeros(8, device=cuda)], [-1.0])
    with pytest.raises(ValueError):
        call.step([torch.zeros(7, device=cuda)], count=0)
    with pytest.raises(ValueError):
        ops.AdamCall(p * 17, p * 17, p * 17, [-1.0] * 17)


@pytest.mark.parametrize("C_,gi,go", [(4, 16, 23), (3, 9, 5), (16, 128, 162), (2, 40, 16), (5, 12, 12), (1, 2, 7)])
def test_vm_resize_matches_oracle(cuda, C_, gi, go):
    from tensorf_b200 import ops
    rng = np.random.default_rng(4)
    v = rng.normal(size=(3, C_, gi)).astype(np.float32)
    m = rng.normal(size=(3, C_, gi, gi)).astype(np.float32)
    vo, mo = ops.vm_resize(T(v, device=cuda), T(m, device=cuda), go)
    rv, rm = O.vm_resize(v, m, go)
    rv64, rm64 = O.vm_resize(v, m, go, dtype=np.float64)
    assert tuple(vo.shape) == (3, C_, go) and tuple(mo.shape) == (3, C_, go, go)
    # fp32 sums of <= 6x6 taps in a different order: 1e-5 absolute on O(1) data; both within that of fp64
    np.testing.assert_allclose(vo.cpu().numpy(), rv, rtol=0, atol=1e-5)
    np.testing.assert_allclose(mo.cpu().numpy(), rm, rtol=0, atol=1e-5)
    np.testing.assert_allclose(vo.cpu().numpy(), rv64, rtol=0, atol=5e-4)   # fp32 sample positions: ~G*eps*slope
    np.testing.assert_allclose(mo.cpu().numpy(), rm64, rtol=0, atol=5e-4)


def test_train_state_adam_and_resize_grid(cuda):
    """training_step's optimiser half and resize_grid against the oracle, leaf by leaf."""
    from tensorf_b200 import cameras, synthetic as S, train_config, training
    cfg = train_config.lego_config(grid_dim_init=16, minibatch_size=128, appearance_feat_dim=8, density_feat_dim=4)
    state = training.TrainState.initialize(cfg, grid_dim=16, prng_key=0, num_cameras=10, device=cuda)
    o, d, c = S.lego_rays(128, seed=5)
    mb = training.RenderedRays(colors=T(S.make_colors(128), device=cuda),
                               rays_wrt_world=cameras.Rays3D(T(o, device=cuda), T(d, device=cuda), T(c, device=cuda)))
    names = list(state.learnable_params.flat().keys())
    oc = cfg.optimizer
    neg_lrs = [-(oc.lr_init_tensor if k.startswith(("density_", "appearance_")) else oc.lr_init_mlp) for k in names]
    for it in range(3):
        flat = state.learnable_params.flat()
        p = [flat[k].detach().cpu().numpy().copy() for k in names]
        mu = [state.optimizer_state["mu"][k].cpu().numpy().copy() for k in names]
        nu = [state.optimizer_state["nu"][k].cpu().numpy().copy() for k in names]
        keys = __import__("tensorf_b200.prng", fromlist=["split"]).split(state.prng_key)
        _, grads = state.loss_and_grads(mb, keys[0])
        g = [grads[k].cpu().numpy().copy() for k in names]
        coeff = O.lr_decay_coeff(state.step, cfg.upsamp_iters, cfg.n_iters, oc.lr_decay_iters, oc.lr_decay_target_ratio,
                                 oc.lr_upsample_reset)
        state, log = state.training_step(mb)
        want_p, want_mu, want_nu = O.adam_step(p, g, mu, nu, it, neg_lrs, coeff)
        flat = state.learnable_params.flat()
        for i, k in enumerate(names):
            # gradients of the two evaluations differ in atomic order only; Adam's first steps normalise them,
            # so compare updates at 1e-4 of the learning rate
            tol = 1e-4 * abs(neg_lrs[i]) + 1e-7
            np.testing.assert_allclose(flat[k].detach().cpu().numpy(), want_p[i], rtol=0, atol=50 * tol, err_msg=k)
            np.testing.assert_allclose(state.optimizer_state["mu"][k].cpu().numpy(), want_mu[i], rtol=1e-3,
                                       atol=1e-4 * np.abs(want_mu[i]).max() + 1e-12, err_msg=k)
        assert abs(log["train/grad_norm"] - O.global_norm(g)) <= 1e-3 * O.global_norm(g)
    # resize_grid: parameters and both moments
    flat = state.learnable_params.flat()
    before = {k: flat[k].detach().cpu().numpy().copy() for k in names}
    mom = {m_: {k: state.optimizer_state[m_][k].cpu().numpy().copy() for k in names} for m_ in ("mu", "nu")}
    state = state.resize_grid(21)
    flat = state.learnable_params.flat()
    assert state.learnable_params.density_tensor.grid_dim() == 21 and state.sample_counts() == (36, 5)
    for which in ("density", "appearance"):
        rv, rm = O.vm_resize(before[f"{which}_vector"], before[f"{which}_matrix"], 21)
        np.testing.assert_allclose(flat[f"{which}_vector"].detach().cpu().numpy(), rv, rtol=0, atol=1e-5)
        np.testing.assert_allclose(flat[f"{which}_matrix"].detach().cpu().numpy(), rm, rtol=0, atol=1e-5)
        for m_ in ("mu", "nu"):
            rv, rm = O.vm_resize(mom[m_][f"{which}_vector"], mom[m_][f"{which}_matrix"], 21)
            scale = max(np.abs(rm).max(), 1e-30)
            np.testing.assert_allclose(state.optimizer_state[m_][f"{which}_matrix"].cpu().numpy(), rm, rtol=0, atol=1e-5 * scale)
            np.testing.assert_allclose(state.optimizer_state[m_][f"{which}_vector"].cpu().numpy(), rv, rtol=0,
                                       atol=1e-5 * max(np.abs(rv).max(), 1e-30))
    for k in ("w0", "w1", "b1", "w2"):
        assert np.array_equal(flat[k].detach().cpu().numpy(), before[k])      # MLP leaves untouched
    state, log = state.training_step(mb)                                       # and training continues on the new grid
    assert np.isfinite(log["train/mse"])

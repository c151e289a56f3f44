"""End-to-end parity of the CUDA hot path (render.py:105-279 + training.py:108-156 forward and
reverse) against the CPU oracle, through the C ABI, on the same seeded inputs.

Tolerances (BASELINE.json north_star / SURVEY.md §8c): selection index sets exact up to audited
near-ties; rgb / depth / loss |x-ref| <= 1e-4*max(1,|ref|); gradients per leaf
||g-ref||_inf <= 1e-4*||ref||_inf and relative L2 <= 1e-4, against the fp64 oracle."""
import numpy as np
import pytest
import torch

import tensorf_oracle as O
from helpers import (T, assert_close_grad, assert_close_out, audit_median_mismatches, device_inputs, kink_rows, oracle_cfgs,
                     oracle_inputs)
from tensorf_b200 import synthetic as S

pytestmark = pytest.mark.gpu

SMALL = S.Workload("small", 64, 9, 2, 3, 37, 5, 2, 2)
SMALL6 = S.Workload("small6", 64, 9, 2, 3, 37, 5, 6, 6)
MID = S.Workload("mid", 512, 48, 16, 48, 83, 12, 2, 2)
ODD = S.Workload("odd", 33, 11, 5, 6, 50, 50, 1, 3)  # cd, ca not multiples of 4; K == N
DOZER_S = S.Workload("dozer_s", 48, 16, 32, 48, 90, 13, 6, 6, contracted=True, num_cameras=7)
DOZER_T = S.Workload("dozer_t", 40, 9, 4, 3, 41, 6, 2, 2, contracted=True, num_cameras=None)
FUSED_S = S.Workload("fused_s", 96, 12, 4, 16, 45, 7, 2, 2)  # 3*ca = 48: smallest shape family of the fused MLP kernels
FUSED_C = S.Workload("fused_c", 64, 10, 4, 32, 43, 6, 2, 2, contracted=True, num_cameras=None)


# BASELINE.json shapes at grid_dim_final = 300 (training.py:115-118, render_360.py:43-51) on a ray slice the fp64 oracle
# finishes in seconds: B = configs[2] (N=519, K=77), C = configs[3] at 300^3 (N=1558, K=233, contraction + embeddings),
# D = configs[4] (N=512, K=128)
LEGO300 = S.lego_workload(R=48, G=300)
DOZER300 = S.dozer_workload(R=32, G=300)
RENDER300 = S.Workload("render300", 40, 300, 16, 48, 512, 128, 2, 2)
# gradient slices at the bench shapes: gradients are sums over rays, so a 64-ray call has a well-defined oracle answer
LEGO_A_SLICE = S.lego_workload(R=64, G=128, N=256, K=38, name="legoA_slice")
DOZER128_SLICE = S.dozer_workload(R=40, G=128)


def fused_ok(w):
    """Shapes TENSORF_MLP_FUSED accepts (csrc/mlp_fused.cu: mlp_fused_supported)."""
    return (3 * w.ca) % 16 == 0 and 3 * w.ca <= 160 and w.feat_freqs == 2 and w.view_freqs == 2 and not w.num_cameras


def run_cuda(w, inp, cuda, with_colors=True, mlp_impl=0):
    from tensorf_b200 import ops
    desc = ops.make_desc(R=w.R, N=w.N, K=w.K, G=w.G, cd=w.cd, ca=w.ca, contracted=w.contracted, feat_freqs=w.feat_freqs,
                         view_freqs=w.view_freqs, num_cameras=w.num_cameras, loss_scale=1.0 / (3 * w.R), mlp_impl=mlp_impl)
    call = ops.RenderCall(desc, cuda)
    params, dins = device_inputs(w, inp, cuda, with_colors)
    rgb, loss = call.forward(params, dins)
    return call, params, dins, rgb, loss


def audit_selection(idx_cuda, aux, K):
    """Index sets must agree; rows that differ must be near-ties of g (few ulp)."""
    ref = np.sort(aux["indices"].numpy(), axis=-1)
    bad_rows = np.where((idx_cuda != ref).any(axis=-1))[0]
    g = aux["g"].numpy()
    for r in bad_rows:
        only_c = np.setdiff1d(idx_cuda[r], ref[r])
        only_r = np.setdiff1d(ref[r], idx_cuda[r])
        assert len(only_c) == len(only_r)
        gap = np.abs(np.sort(g[r, only_c]) - np.sort(g[r, only_r]))
        scale = np.maximum(np.abs(g[r, only_r]).max(), 1.0)
        assert (gap <= 64 * np.finfo(np.float32).eps * scale).all(), f"row {r}: selection differs beyond a near-tie: {gap}"
    return len(bad_rows)


@pytest.mark.parametrize("mlp_impl", [2, 1, 3], ids=["tcgen05", "simt_fp32", "fused"])
@pytest.mark.parametrize("w", [SMALL, SMALL6, MID, ODD, DOZER_S, DOZER_T, FUSED_S, FUSED_C, LEGO300, DOZER300, RENDER300, LEGO_A_SLICE,
                               DOZER128_SLICE], ids=lambda w: w.name)
def test_render_rgb_forward_and_grads(cuda, w, mlp_impl):
    if w.G >= 128 and mlp_impl == 1:
        pytest.skip("full-size shapes: the exact-fp32 SIMT MLP is covered on the small shapes")
    if mlp_impl == 3 and not fused_ok(w):
        from tensorf_b200 import _lib
        with pytest.raises(_lib.TensorfError):  # the fused kernels refuse other networks instead of falling back silently
            run_cuda(w, S.make_inputs(w, bias_std=0.05), cuda, mlp_impl=3)
        return
    inp = S.make_inputs(w, bias_std=0.05)
    call, params, dins, rgb, loss = run_cuda(w, inp, cuda, mlp_impl=mlp_impl)
    idx = call.view("idx").cpu().numpy().reshape(w.R, w.K)
    assert (np.diff(idx, axis=-1) > 0).all() or w.K == 1

    cfg, mc = oracle_cfgs(w)
    oi = oracle_inputs(inp, torch.float32)
    out32, aux = O.render_rays(cfg, mc, oi["params"], w.contracted, oi["aabb"], oi["origins"], oi["directions"],
                               oi["camera_indices"], oi["jitter"], oi["gumbel"], return_aux=True)
    n_bad = audit_selection(idx, aux, w.K)
    assert n_bad <= max(1, w.R // 100), f"{n_bad} rays with a different (near-tie) selection"

    # residuals
    assert_close_out(call.view("z").cpu().numpy().reshape(w.R, w.N), aux["z"].numpy(), what="z")
    # downstream values with the kernel's own selection forced into the oracle (fp64 arbiter)
    forced = torch.from_numpy(idx.astype(np.int64))
    o64 = oracle_inputs(inp, torch.float64)
    leaves = {k: v.clone().requires_grad_(True) for k, v in o64["params"].items()}
    rgb64, aux64 = O.render_rays(cfg, mc, leaves, w.contracted, o64["aabb"], o64["origins"], o64["directions"],
                                 o64["camera_indices"], o64["jitter"], o64["gumbel"], return_aux=True, forced_indices=forced)
    loss64 = torch.mean((rgb64 - o64["colors"]) ** 2)
    assert_close_out(rgb.cpu().numpy(), rgb64.detach().numpy(), what="rgb")
    assert_close_out(loss.cpu().numpy(), loss64.detach().numpy(), what="loss")
    good = np.where((idx == np.sort(aux["indices"].numpy(), axis=-1)).all(axis=-1))[0]
    assert_close_out(rgb.cpu().numpy()[good], out32.numpy()[good], what="rgb vs fp32 oracle (own selection)")

    # reverse mode with an explicit cotangent: d loss / d rgb of the MSE (training.py:140), with
    # the rays that own a ReLU-kink row zeroed for kernel and oracle alike
    d_rgb64 = (2.0 / (3 * w.R)) * (rgb64.detach() - o64["colors"])
    amb_rays = np.unique(kink_rows(aux64) // w.K)
    assert len(amb_rays) <= max(3, w.R // 3)  # (K = 233 rows x 256 units per ray at the 300^3 dozer shape: a few rays own a kink row)
    d_rgb64[amb_rays] = 0.0
    (rgb64 * d_rgb64).sum().backward()
    grads = call.backward(d_rgb64.to(torch.float32).to(cuda).contiguous())
    for k, leaf in leaves.items():
        assert_close_grad(grads[k].cpu().numpy(), leaf.grad.numpy(), what=f"grad {k}")

    # fused-loss cotangent path (d_rgb = NULL) == explicit cotangent 2(rgb - c)/(3R) (same kernels)
    g_fused = {k: v.clone() for k, v in call.backward().items()}
    g_expl = call.backward(((2.0 / (3 * w.R)) * (rgb - dins["colors"])).contiguous())
    for k in g_fused:
        assert_close_grad(g_fused[k].cpu().numpy(), g_expl[k].cpu().numpy(), rtol=2e-5, what=f"fused vs explicit {k}")


@pytest.mark.parametrize("w", [SMALL, MID, DOZER_S, LEGO300, DOZER300, RENDER300], ids=lambda w: w.name)
@pytest.mark.parametrize("mode", [O.DIST_MEDIAN, O.DIST_MEAN])
def test_render_depth(cuda, w, mode):
    from tensorf_b200 import ops
    inp = S.make_inputs(w)
    # make the median interesting: a few nearly transparent rays (never cross 0.5 -> inf) by
    # making some rays miss the box is already covered by lego_rays; also scale density down.
    inp["params"]["density_vector"] *= 0.2
    desc = ops.make_desc(R=w.R, N=w.N, K=w.K, G=w.G, cd=w.cd, ca=w.ca, mode=mode, contracted=w.contracted)
    call = ops.RenderCall(desc, cuda)
    params, dins = device_inputs(w, inp, cuda, with_colors=False)
    depth = call.depth(params, dins).cpu().numpy()
    cfg, mc = oracle_cfgs(w, mode)
    oi = oracle_inputs(inp, torch.float32)
    ref = O.render_rays(cfg, mc, oi["params"], w.contracted, oi["aabb"], oi["origins"], oi["directions"],
                        oi["camera_indices"], oi["jitter"], None).numpy()
    if mode == O.DIST_MEDIAN:
        # the crossing sample can legitimately differ when 1-E is within rounding of 0.5: every mismatch is audited
        o64 = oracle_inputs(inp, torch.float64)
        _, aux64 = O.render_rays(cfg, mc, o64["params"], w.contracted, o64["aabb"], o64["origins"], o64["directions"],
                                 o64["camera_indices"], o64["jitter"], None, return_aux=True)
        n_bad = audit_median_mismatches(depth, ref, aux64, w.N)
        assert n_bad <= max(1, w.R // 50), f"median depth: {n_bad} audited threshold cases of {w.R} rays"
    else:
        assert_close_out(depth, ref, what="mean depth")


def test_zero_rays_and_errors(cuda):
    from tensorf_b200 import ops
    from tensorf_b200._lib import TensorfError
    w = SMALL
    inp = S.make_inputs(w)
    params, dins = device_inputs(w, inp, cuda)
    with pytest.raises(TensorfError):  # K > N
        ops.RenderCall(ops.make_desc(R=4, N=5, K=6, G=9, cd=2, ca=3), cuda)
    with pytest.raises(TensorfError):  # units != 128 unsupported
        ops.RenderCall(ops.make_desc(R=4, N=5, K=2, G=9, cd=2, ca=3, units=64), cuda)
    call = ops.RenderCall(ops.make_desc(R=w.R, N=w.N, K=w.K, G=w.G, cd=w.cd, ca=w.ca, feat_freqs=2, view_freqs=2), cuda)
    bad = dict(dins)
    bad["jitter"] = dins["jitter"][:-1]
    with pytest.raises(ValueError):
        call.forward(params, bad)
    bad = dict(dins)
    del bad["gumbel"]
    with pytest.raises(KeyError):
        call.forward(params, bad)

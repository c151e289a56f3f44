"""CUDA FeatureMlp (networks.py:38-121) forward / reverse vs the CPU oracle."""
import numpy as np
import pytest
import torch

import tensorf_oracle as O
from helpers import T, assert_close_grad, assert_close_out, kink_rows
from tensorf_b200 import synthetic as S

pytestmark = pytest.mark.gpu

MLP_LEAVES = ("w0", "w1", "b1", "w2", "b2", "w3", "b3", "embed")


@pytest.mark.parametrize("ca,F,V,ncam,rays,rpr", [(3, 2, 2, None, 70, 5), (48, 2, 2, None, 300, 33), (48, 6, 6, 7, 64, 9),
                                                  (5, 0, 1, 3, 130, 1), (24, 6, 6, None, 257, 3)])
@pytest.mark.parametrize("impl", [1, 2], ids=["simt_fp32", "tcgen05"])
def test_mlp_fwd_bwd(cuda, ca, F, V, ncam, rays, rpr, impl):
    from tensorf_b200 import ops
    M = rays * rpr
    p_np = S.make_params(4, 1, ca, F, V, ncam, seed=11, bias_std=0.1)
    rng = np.random.default_rng(7)
    feat = rng.normal(0, 0.3, (M, 3 * ca)).astype(np.float32)
    vd = rng.normal(size=(rays, 3)).astype(np.float32)
    vd /= np.linalg.norm(vd, axis=-1, keepdims=True)
    cams = rng.integers(0, ncam or 1, size=rays).astype(np.int32)
    d_rgb = rng.normal(size=(M, 3)).astype(np.float32)

    mc = O.MlpConfig(27, 128, F, V, ncam)
    P64 = {k: T(v, torch.float64).requires_grad_(True) for k, v in p_np.items() if k in MLP_LEAVES}
    f64 = T(feat, torch.float64).requires_grad_(True)
    vd64 = T(vd, torch.float64).repeat_interleave(rpr, dim=0)
    cams64 = torch.from_numpy(cams.astype(np.int64)).repeat_interleave(rpr)
    aux = {}
    ref = O.feature_mlp(mc, P64, f64, vd64, cams64, aux=aux)
    amb = kink_rows(aux)
    assert len(amb) <= max(3, M // 200)
    d_rgb[amb] = 0.0  # rows on a ReLU kink: zero cotangent for kernel and oracle alike
    (ref * T(d_rgb, torch.float64)).sum().backward()

    desc = ops.make_desc(R=rays, N=rpr, K=rpr, G=4, cd=1, ca=ca, feat_freqs=F, view_freqs=V, num_cameras=ncam, mlp_impl=impl)
    call = ops.MlpCall(desc, M, cuda)
    params = {k: T(v, device=cuda) for k, v in p_np.items()}
    rgb = call.forward(params, T(feat, device=cuda), T(vd, device=cuda), T(cams, device=cuda), rpr)
    assert_close_out(rgb.cpu().numpy(), ref.detach().numpy(), what="mlp rgb")
    d_feat, grads = call.backward(T(d_rgb, device=cuda))
    assert_close_grad(d_feat.cpu().numpy(), f64.grad.numpy(), what="d_features")
    for k in P64:
        assert_close_grad(grads[k].cpu().numpy(), P64[k].grad.numpy(), what=f"d {k}")

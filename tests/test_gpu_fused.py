"""Fused tcgen05 MLP kernels (csrc/mlp_fused.cu): operand conventions and per-kernel parity."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(x, ref):
    return float(np.abs(x - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("a_mode,b_mode", [(0, 0), (1, 0), (0, 1), (1, 1), (2, 0), (2, 1)])
@pytest.mark.parametrize("a_fmt,b_fmt", [(0, 0), (1, 1)])
@pytest.mark.parametrize("N,K", [(128, 128), (32, 144), (128, 160), (16, 32), (80, 16)])
def test_umma_probe(cuda, a_mode, b_mode, a_fmt, b_fmt, N, K):
    """Slab tiles as K-major / MN-major shared-memory operands and A from tensor memory, fp16 and bf16 terms.
    (Mixed fp16 x bf16 operands in one kind::f16 instruction raise an illegal-instruction fault on sm_100a - measured.)"""
    from tensorf_b200 import _lib, ops
    rng = np.random.default_rng(1000 * a_mode + 100 * b_mode + 10 * a_fmt + b_fmt + N + K)
    A = torch.from_numpy(rng.normal(size=(128, K)).astype(np.float32)).to(cuda)
    B = torch.from_numpy((rng.normal(size=(N, K)) / np.sqrt(K)).astype(np.float32)).to(cuda)
    D = torch.full((128, N), float("nan"), dtype=torch.float32, device=cuda)
    _lib.check(_lib.load().tensorf_tc_umma_probe(ops._stream(), A.data_ptr(), B.data_ptr(), D.data_ptr(), N, K, a_mode, b_mode,
                                                 a_fmt, b_fmt))
    ref = A.double().cpu() @ B.double().cpu().T
    err = _rel(D.cpu().numpy(), ref.numpy())
    tol = 3e-6 if (a_fmt == 0 and b_fmt == 0) else 3e-5
    assert err < tol, f"probe a_mode={a_mode} b_mode={b_mode} fmt={a_fmt}{b_fmt} N={N} K={K}: rel err {err:.3e}"


# ---- fused FeatureMlp forward (k_mlp_fused_fwd) -------------------------------------------------------------------
def _x_perm():
    """Kernel column order of the encoded MLP input (csrc/mlp_fused.cu, fz::perm): kernel column -> reference column."""
    out = []
    for kp in range(160):
        d, v = 8 * (kp // 40) + (kp % 40) // 5, kp % 5
        out.append(-1 if d >= 30 else d if v == 0 else 30 + 4 * d + (v - 1))
    return np.array(out)


def _unslab(buf_u8, tiles, cols):
    """[tile][(c>>3)*2+term][row 128][8 x fp16] -> (tiles*128, cols) float64 = hi + lo."""
    a = buf_u8[: tiles * cols * 512].view(np.float16).reshape(tiles, cols // 8, 2, 128, 8).astype(np.float64)
    v = a[:, :, 0] + a[:, :, 1]  # (tiles, c8, row, 8)
    return v.transpose(0, 2, 1, 3).reshape(tiles * 128, cols)


def _mlp_case(ca, rays, rpr, seed=7):
    import tensorf_oracle as O
    from helpers import T
    from tensorf_b200 import synthetic as S
    M = rays * rpr
    p_np = S.make_params(4, 1, ca, 2, 2, None, seed=11, bias_std=0.1)
    rng = np.random.default_rng(seed)
    feat = rng.normal(0, 0.3, (M, 3 * ca)).astype(np.float32)
    vd = rng.normal(size=(rays, 3)).astype(np.float32)
    vd /= np.linalg.norm(vd, axis=-1, keepdims=True)
    mc = O.MlpConfig(27, 128, 2, 2, None)
    leaves = ("w0", "w1", "b1", "w2", "b2", "w3", "b3")
    P64 = {k: T(v, torch.float64).requires_grad_(True) for k, v in p_np.items() if k in leaves}
    f64 = T(feat, torch.float64).requires_grad_(True)
    vd64 = T(vd, torch.float64).repeat_interleave(rpr, dim=0)
    return M, p_np, feat, vd, mc, P64, f64, vd64


@pytest.mark.parametrize("ca,rays,rpr", [(48, 300, 33), (16, 7, 19), (48, 128, 1), (32, 2000, 38)])
@pytest.mark.parametrize("inference", [True, False])
def test_fused_mlp_forward(cuda, ca, rays, rpr, inference):
    """tensorf_mlp_fwd with TENSORF_MLP_FUSED vs the fp64 oracle (networks.py:46-121); in training mode also the
    residuals the reverse kernels consume: x' / h1 / h2 slab tiles (fp16 hi + lo), ReLU mask words, Dense_0 output."""
    import ctypes as C

    import tensorf_oracle as O
    from helpers import T, assert_close_out
    from tensorf_b200 import _lib, ops
    M, p_np, feat, vd, mc, P64, f64, vd64 = _mlp_case(ca, rays, rpr)
    aux = {}
    with torch.no_grad():
        ref = O.feature_mlp(mc, P64, f64, vd64, None, aux=aux)
    desc = ops.make_desc(R=rays, N=rpr, K=rpr, G=4, cd=1, ca=ca, feat_freqs=2, view_freqs=2, mlp_impl=ops.MLP_FUSED,
                         inference=inference)
    call = ops.MlpCall(desc, M, cuda)
    call.workspace.fill_(float("nan"))
    params = {k: T(v, device=cuda) for k, v in p_np.items()}
    rgb = call.forward(params, T(feat, device=cuda), T(vd, device=cuda), None, rpr)
    torch.cuda.synchronize()
    assert_close_out(rgb.cpu().numpy(), ref.numpy(), what="fused mlp rgb")
    assert np.abs(rgb.cpu().numpy() - ref.numpy()).max() < 5e-6
    if inference:
        return
    off = (C.c_int64 * 10)()
    _lib.check(_lib.load().tensorf_mlp_workspace_layout(C.byref(desc), M, off))
    ws = call.workspace.cpu().numpy()
    tiles = (M + 127) // 128
    with torch.no_grad():
        f_ref = (f64 @ P64["w0"]).numpy()
        x_ref = torch.cat([f64 @ P64["w0"], vd64, O.fourier_encode(f64 @ P64["w0"], 2), O.fourier_encode(vd64, 2)], dim=-1).numpy()
        h1_ref = np.maximum(aux["z1"].numpy(), 0)
        h2_ref = np.maximum(aux["z2"].numpy(), 0)
    fs = ws[off[0]: off[0] + tiles * 4096].reshape(tiles, 8, 128, 4).transpose(0, 2, 1, 3).reshape(tiles * 128, 32)[:M]
    assert np.abs(fs[:, :27] - f_ref).max() < 2e-6 * max(1.0, np.abs(f_ref).max())
    assert np.all(fs[:, 27:] == 0)
    xs = _unslab(ws[off[2]:].view(np.uint8), tiles, 160)[:M]
    perm = _x_perm()
    assert np.abs(xs[:, perm >= 0] - x_ref[:, perm[perm >= 0]]).max() < 3e-6
    assert np.all(xs[:, 150] == 1.0) and np.all(xs[:, 151:] == 0)
    h1 = _unslab(ws[off[4]:].view(np.uint8), tiles, 128)[:M]
    h2 = _unslab(ws[off[5]:].view(np.uint8), tiles, 128)[:M]
    assert np.abs(h1 - h1_ref).max() < 5e-6 * max(1.0, np.abs(h1_ref).max())
    assert np.abs(h2 - h2_ref).max() < 5e-6 * max(1.0, np.abs(h2_ref).max())
    for o, h in ((off[8], h1), (off[9], h2)):
        words = ws[o: o + tiles * 128 * 4].view(np.uint32).reshape(tiles * 128, 4)[:M]
        bits = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(M, 128).astype(bool)
        assert np.array_equal(bits, h > 0)


@pytest.mark.parametrize("ca,rays,rpr", [(48, 300, 33), (16, 7, 19), (48, 128, 1), (32, 2000, 38)])
@pytest.mark.parametrize("scale", [1.0, 3e-6])
def test_fused_mlp_reverse(cuda, ca, rays, rpr, scale):
    """tensorf_mlp_fwd + tensorf_mlp_bwd with TENSORF_MLP_FUSED (k_mlp_fused_bwd, k_mlp_fused_wgrad) vs fp64 autograd of the
    oracle: d_features and every MLP leaf within 1e-4 (inf and L2, relative to the leaf).  `scale` shrinks the incoming
    cotangent to the magnitudes of a real training step (loss_scale ~ 1e-4 x small residuals): the kernels carry the
    gradients multiplied by a power of two so that fp16 terms do not underflow."""
    import tensorf_oracle as O
    from helpers import T, assert_close_grad, kink_rows
    from tensorf_b200 import ops
    M, p_np, feat, vd, mc, P64, f64, vd64 = _mlp_case(ca, rays, rpr)
    rng = np.random.default_rng(5)
    d_rgb = (scale * rng.normal(size=(M, 3)) * rng.uniform(0, 1, size=(M, 1)) ** 4).astype(np.float32)
    aux = {}
    ref = O.feature_mlp(mc, P64, f64, vd64, None, aux=aux)
    amb = kink_rows(aux)
    d_rgb[amb] = 0.0
    (ref * T(d_rgb, torch.float64)).sum().backward()
    desc = ops.make_desc(R=rays, N=rpr, K=rpr, G=4, cd=1, ca=ca, feat_freqs=2, view_freqs=2, mlp_impl=ops.MLP_FUSED)
    call = ops.MlpCall(desc, M, cuda)
    params = {k: T(v, device=cuda) for k, v in p_np.items()}
    call.forward(params, T(feat, device=cuda), T(vd, device=cuda), None, rpr)
    d_feat, grads = call.backward(T(d_rgb, device=cuda))
    torch.cuda.synchronize()
    assert_close_grad(d_feat.cpu().numpy(), f64.grad.numpy(), what="fused d_features")
    for k in P64:
        assert_close_grad(grads[k].cpu().numpy(), P64[k].grad.numpy(), what=f"fused d {k}")

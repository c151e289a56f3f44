"""tcgen05 GEMM building blocks of the MLP vs fp64 matmul (split-bf16 operands, fp32 accumulate)."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(x, ref):
    return float(np.abs(x - ref).max() / np.abs(ref).max())


@pytest.mark.parametrize("nsplit,tol", [(2, 2e-5), (3, 3e-6)])
@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (300, 144, 27), (1000, 150, 128), (4099, 128, 128), (777, 390, 128),
                                   (513, 27, 144), (256, 128, 150), (20000, 160, 128), (700, 128, 390)])
def test_rowgemm(cuda, M, K, N, nsplit, tol):
    from tensorf_b200 import _lib, ops
    rng = np.random.default_rng(M + K + N)
    A = torch.from_numpy(rng.normal(size=(M, K)).astype(np.float32)).to(cuda)
    W = torch.from_numpy((rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)).to(cuda)
    bias = torch.from_numpy(rng.normal(size=(N,)).astype(np.float32)).to(cuda)
    mask_np = rng.normal(size=(M, N)).astype(np.float32) > 0
    words = (((N + 15) // 16 * 16) + 31) // 32
    bits_np = np.zeros((M, words * 32), dtype=bool)
    bits_np[:, :N] = mask_np
    bits = torch.from_numpy(np.packbits(bits_np.reshape(M, words, 32), axis=-1, bitorder="little").view(np.uint32).reshape(M, words).view(np.int32)).to(cuda)
    bits_out = torch.zeros((M, words), dtype=torch.int32, device=cuda)
    out = torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
    scratch = torch.empty(128 * 1024 * ((K + 63) // 64 + 1) * ((N + 255) // 256), dtype=torch.uint8, device=cuda)
    lib = _lib.load()
    for relu, use_mask in ((0, False), (1, True)):
        _lib.check(lib.tensorf_tc_rowgemm_test(ops._stream(), A.data_ptr(), M, K, W.data_ptr(), N, bias.data_ptr(), relu,
                                               bits.data_ptr() if use_mask and N <= 256 else None,
                                               bits_out.data_ptr() if N <= 256 else None, out.data_ptr(),
                                               scratch.data_ptr(), scratch.numel(), nsplit))
        ref = A.double().cpu() @ W.double().cpu() + bias.double().cpu()
        if relu:
            ref = torch.relu(ref)
        if use_mask and N <= 256:
            ref = ref * torch.from_numpy(mask_np)
        err = _rel(out.cpu().numpy(), ref.numpy())
        assert err < tol, f"rowgemm M={M} K={K} N={N} relu={relu} nsplit={nsplit}: rel err {err:.3e}"
        if N <= 256:  # bits_out = (C > 0), up to sign flips of values within rounding of zero
            got = np.unpackbits(bits_out.cpu().numpy().view(np.uint32).reshape(M, words, 1).view(np.uint8), axis=-1, bitorder="little").reshape(M, words * 32)[:, :N].astype(bool)
            exp = ref.numpy() > 0
            diff = got != exp
            assert (np.abs(ref.numpy()[diff]) < 1e-4).all() and diff.mean() < 1e-3


@pytest.mark.parametrize("rows,Mg,Nx", [(256, 128, 128), (5000, 128, 150), (3001, 27, 144), (40000, 128, 128), (999, 128, 390),
                                          (20000, 128, 400), (7001, 32, 400)])
def test_redgemm(cuda, rows, Mg, Nx):
    from tensorf_b200 import _lib, ops
    rng = np.random.default_rng(rows + Mg + Nx)
    G = torch.from_numpy(rng.normal(size=(rows, Mg)).astype(np.float32)).to(cuda)
    X = torch.from_numpy(rng.normal(size=(rows, Nx)).astype(np.float32)).to(cuda)
    out = torch.zeros((Nx, Mg), dtype=torch.float32, device=cuda)
    _lib.check(_lib.load().tensorf_tc_redgemm_test(ops._stream(), G.data_ptr(), Mg, X.data_ptr(), Nx, rows, out.data_ptr()))
    ref = X.double().cpu().T @ G.double().cpu()
    err = _rel(out.cpu().numpy(), ref.numpy())
    assert err < 2e-5, f"redgemm rows={rows} Mg={Mg} Nx={Nx}: rel err {err:.3e}"


@pytest.mark.parametrize("K,N,nsplit", [(128, 128, 2), (160, 128, 3), (128, 128, 3), (144, 32, 3)])
def test_rowgemm_bitwise_reproducible(cuda, K, N, nsplit):
    """Race detector for the warp-specialised pipeline (TMA raw ring, operand stages, TMEM double buffer): the
    same launch on the same data must give bit-identical results, and every run must match fp64.  148 CTAs x
    8 tiles each keeps all rings wrapping many times."""
    from tensorf_b200 import _lib, ops
    M = 148 * 8 * 128
    rng = np.random.default_rng(7)
    A = torch.from_numpy(rng.normal(size=(M, K)).astype(np.float32)).to(cuda)
    W = torch.from_numpy((rng.normal(size=(K, N)) / np.sqrt(K)).astype(np.float32)).to(cuda)
    bias = torch.zeros(N, device=cuda)
    scratch = torch.empty(8 << 20, dtype=torch.uint8, device=cuda)
    ref = (A.double() @ W.double()).clamp_min(0)
    lib = _lib.load()
    first = None
    for rep in range(4):
        out = torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
        _lib.check(lib.tensorf_tc_rowgemm_test(ops._stream(), A.data_ptr(), M, K, W.data_ptr(), N, bias.data_ptr(), 1, None, None,
                                               out.data_ptr(), scratch.data_ptr(), scratch.numel(), nsplit))
        err = float((out.double() - ref).abs().max())
        assert err < (2e-4 if nsplit == 2 else 3e-5), (rep, err)
        if first is None:
            first = out
        else:
            assert torch.equal(out, first), f"run {rep} differs from run 0 in {int((out != first).sum())} elements"

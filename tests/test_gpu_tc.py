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
    mask = torch.from_numpy(rng.normal(size=(M, N)).astype(np.float32)).to(cuda)
    out = torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
    scratch = torch.empty(128 * 1024 * ((K + 63) // 64 + 1) * ((N + 255) // 256), dtype=torch.uint8, device=cuda)
    lib = _lib.load()
    for relu, use_mask in ((0, False), (1, True)):
        _lib.check(lib.tensorf_tc_rowgemm_test(ops._stream(), A.data_ptr(), M, K, W.data_ptr(), N, bias.data_ptr(), relu,
                                               mask.data_ptr() if use_mask else None, out.data_ptr(), scratch.data_ptr(),
                                               scratch.numel(), nsplit))
        ref = A.double().cpu() @ W.double().cpu() + bias.double().cpu()
        if relu:
            ref = torch.relu(ref)
        if use_mask:
            ref = ref * (mask.cpu() > 0)
        err = _rel(out.cpu().numpy(), ref.numpy())
        assert err < tol, f"rowgemm M={M} K={K} N={N} relu={relu} nsplit={nsplit}: rel err {err:.3e}"


@pytest.mark.parametrize("rows,Mg,Nx", [(256, 128, 128), (5000, 128, 150), (3001, 27, 144), (40000, 128, 128), (999, 128, 390)])
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
